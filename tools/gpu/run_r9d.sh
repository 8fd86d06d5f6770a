# usage: bash tools/gpu/run_r9d.sh TAG -- GPU suite (fast subset), then C3 / C5 bench A/B: lanes on/off, look-back scan on/off
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
TAG=$1
timeout 1500 python -m pytest tests/test_gpu_deferred.py tests/test_gpu_parity.py tests/test_gpu_edge_cases.py tests/test_parallel_gpu.py tests/test_gpu_elasticity.py -m gpu -x -q > gpurun_out/${TAG}_suite.log 2>&1; tail -5 gpurun_out/${TAG}_suite.log
for W in C3 C5; do
for V in default nolanes scan3; do
  unset CFX_NO_LANES CFX_SCAN_3PASS
  [ $V = nolanes ] && export CFX_NO_LANES=1
  [ $V = scan3 ] && export CFX_SCAN_3PASS=1
  timeout 600 python bench.py --workload $W --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_${W}_$V.json 2> gpurun_out/bench_${TAG}_${W}_$V.err; echo "$W $V rc=$?"; tail -2 gpurun_out/bench_${TAG}_${W}_$V.err
  python tools/show_bench.py gpurun_out/bench_${TAG}_${W}_$V.json 2>/dev/null | head -1
done
done
