# usage: bash tools/gpu/run_r9g.sh TAG -- GPU suite + default bench
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
TAG=$1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_suite.log 2>&1; tail -4 gpurun_out/${TAG}_suite.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "rc=$?"; tail -2 gpurun_out/bench_${TAG}.err
python tools/show_bench.py gpurun_out/bench_${TAG}.json
