# usage: bash tools/gpu/run_r9d.sh TAG -- GPU suite, then bench A/B with and without lanes (C3, C5, C2, C4 at 96^3)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
TAG=$1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_suite.log 2>&1; tail -5 gpurun_out/${TAG}_suite.log
for W in C3 C5 C2 C4; do
for V in default nolanes; do
  unset CFX_NO_LANES
  [ $V = nolanes ] && export CFX_NO_LANES=1
  X=""; [ $W = C4 ] && X="--n 96"
  timeout 600 python bench.py --workload $W $X --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_${W}_$V.json 2> gpurun_out/bench_${TAG}_${W}_$V.err; echo "$W $V rc=$?"; tail -2 gpurun_out/bench_${TAG}_${W}_$V.err
  python tools/show_bench.py gpurun_out/bench_${TAG}_${W}_$V.json 2>/dev/null | head -1
done
done
