cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for v in 1 2 3 4; do
CFX_CLIST_GRID_MB=$v timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_gmb$v.json 2> gpurun_out/bench_gmb$v.err
echo GRID_MB=$v; python tools/show_bench.py gpurun_out/bench_gmb$v.json | grep "ms/step\|clist"
done
