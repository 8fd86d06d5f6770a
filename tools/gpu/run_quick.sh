# usage: bash tools/gpu/run_quick.sh TAG [pytest args]  -- GPU tests, then one bench line into gpurun_out/bench_TAG.json
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
TAG=${1:-x}; shift
TESTS=${@:-tests/test_gpu_parity.py tests/test_gpu_edge_cases.py}
timeout 1200 python -m pytest $TESTS -m gpu -x -q 2>&1 | tail -30
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
python tools/show_bench.py gpurun_out/bench_$TAG.json
