# usage: bash tools/gpu/run_r5a.sh TAG -- round-2 first GPU session: FP64 peaks, deferred/graph tests (+ memcheck),
# whole GPU suite (incl. oracle parity at 128^3 / 256^3), bench lines for C3 / C5 / C2
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
TAG=${1:-r5a}
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader; nproc; free -g | head -2
./tools/gpu/fp64_peak > gpurun_out/${TAG}_fp64_peak.json 2> gpurun_out/${TAG}_fp64_peak.err; cat gpurun_out/${TAG}_fp64_peak.json
echo "== deferred tests"
timeout 900 python -m pytest tests/test_gpu_deferred.py -m gpu -q -x > gpurun_out/${TAG}_deferred.log 2>&1; grep -E "^E  |passed|failed|test_gpu_deferred.py:[0-9]+: in" gpurun_out/${TAG}_deferred.log | head -30
echo "== gpu suite"
if [ "$2" != "nosuite" ]; then timeout 2400 python -m pytest tests -m gpu -q --deselect tests/test_gpu_deferred.py 2>&1 | tail -12 | tee gpurun_out/${TAG}_suite.log; fi
echo "== bench"
for W in C3 C5 C2; do
timeout 900 python bench.py --workload $W --steps 10 --warmup 3 > gpurun_out/bench_${TAG}_$W.json 2> gpurun_out/bench_${TAG}_$W.err
echo "bench $W rc=$?"; tail -3 gpurun_out/bench_${TAG}_$W.err
python tools/show_bench.py gpurun_out/bench_${TAG}_$W.json
done
