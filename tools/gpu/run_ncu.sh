# usage: bash tools/gpu/run_ncu.sh TAG KERNEL_REGEX [KERNEL_REGEX..] -- one `ncu --set full` capture per kernel (first launch after warm-up)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
TAG=$1; shift
for K in "$@"; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 2 -c 1 -o gpurun_out/prof_${TAG}_$K -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_${TAG}_$K.log 2>&1
tail -2 gpurun_out/ncu_${TAG}_$K.log
done
