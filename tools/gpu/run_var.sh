cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for mb in 4 5 6; do
CFX_CLIST_MB=$mb timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_mb$mb.json 2> gpurun_out/bench_mb$mb.err
echo MB=$mb; python tools/show_bench.py gpurun_out/bench_mb$mb.json | grep "ms/step\|clist"
done
