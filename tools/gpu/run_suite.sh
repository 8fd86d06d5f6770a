# usage: bash tools/gpu/run_suite.sh TAG [WORKLOADS...] -- whole GPU suite, then bench lines (default C3 C5)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
TAG=$1; shift
WL=${@:-C3 C5}
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_suite.log 2>&1; grep -E "^E  |passed|failed|\.py:[0-9]+: in|Error" gpurun_out/${TAG}_suite.log | head -30
for W in $WL; do
timeout 900 python bench.py --workload $W --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_$W.json 2> gpurun_out/bench_${TAG}_$W.err
echo "bench $W rc=$?"; tail -3 gpurun_out/bench_${TAG}_$W.err
python tools/show_bench.py gpurun_out/bench_${TAG}_$W.json
done
