# usage: bash tools/gpu/run_quick2.sh TAG [pytest selection...] -- selected GPU tests, then a C3 bench line (+C5 if TAG ends in 5)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
TAG=$1; shift
TESTS=${@:-tests/test_gpu_parity.py tests/test_gpu_edge_cases.py tests/test_gpu_deferred.py}
timeout 1500 python -m pytest $TESTS -m gpu -q -x > gpurun_out/${TAG}_tests.log 2>&1; grep -E "^E  |passed|failed|\.py:[0-9]+: in|Error" gpurun_out/${TAG}_tests.log | head -30
for W in C3 C5; do
timeout 600 python bench.py --workload $W --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_$W.json 2> gpurun_out/bench_${TAG}_$W.err
echo "bench $W rc=$?"; tail -3 gpurun_out/bench_${TAG}_$W.err
python tools/show_bench.py gpurun_out/bench_${TAG}_$W.json
done
