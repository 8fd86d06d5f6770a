# usage: bash tools/gpu/run_quick3.sh TAG -- parity subset (with NaN-poisoned values) + one C3 bench line
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
TAG=$1
CFX_POISON_VALUES=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_deferred.py tests/test_parity_at_size.py tests/test_parallel_gpu.py tests/test_gpu_edge_cases.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "rc=$?"; tail -1 gpurun_out/bench_$TAG.err
python tools/show_bench.py gpurun_out/bench_$TAG.json
