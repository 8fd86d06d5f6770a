set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py -m gpu -x -q 2>&1 | tail -15
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r3a.json 2> gpurun_out/bench_r3a.err
tail -c 3000 gpurun_out/bench_r3a.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gather_matrix_clist_kernel -c 1 -o gpurun_out/prof_r3a -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_r3a.log 2>&1
tail -3 gpurun_out/ncu_r3a.log
