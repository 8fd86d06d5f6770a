# usage: bash tools/gpu/run_final.sh TAG -- what the driver runs at round end: GPU suite, smoke(), default bench, reference arm
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
TAG=$1
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_suite.log 2>&1; tail -3 gpurun_out/${TAG}_suite.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_${TAG}.err
python tools/show_bench.py gpurun_out/bench_${TAG}.json
timeout 900 python bench.py --impl reference > gpurun_out/bench_${TAG}_ref.json 2> gpurun_out/bench_${TAG}_ref.err; echo "ref rc=$?"; cut -c1-700 gpurun_out/bench_${TAG}_ref.json
