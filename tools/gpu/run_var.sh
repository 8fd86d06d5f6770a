cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for v in 32 16 8; do
CFX_CLIST_CH=$v timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ch$v.json 2> gpurun_out/bench_ch$v.err
echo CLIST_CH=$v; python tools/show_bench.py gpurun_out/bench_ch$v.json | grep "ms/step\|clist"
done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
