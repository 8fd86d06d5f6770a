# usage: bash tools/gpu/run_r9l.sh TAG -- GPU suite with poisoned (NaN) matrix values where cfx_create_sparsity wrote nothing,
# then A/B of the lazy zero fill (CFX_EAGER_ZERO=1: the old full memset) on C3, C2, C4 at 96^3
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
TAG=$1
CFX_POISON_VALUES=1 timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_suite_poison.log 2>&1; tail -4 gpurun_out/${TAG}_suite_poison.log
for W in C3 C2 C4; do
for V in default eagerzero; do
  unset CFX_EAGER_ZERO
  [ $V = eagerzero ] && export CFX_EAGER_ZERO=1
  X=""; [ $W = C4 ] && X="--n 96"
  timeout 600 python bench.py --workload $W $X --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_${W}_$V.json 2> gpurun_out/bench_${TAG}_${W}_$V.err; echo "$W $V rc=$?"; tail -2 gpurun_out/bench_${TAG}_${W}_$V.err
  python tools/show_bench.py gpurun_out/bench_${TAG}_${W}_$V.json 2>/dev/null | grep "ms/step\|create_sparsity"
done
done
