# usage: bash tools/gpu/run_ncu_c4.sh TAG N KERNEL_REGEX... -- ncu --set full of kernels of the elasticity step (try_c4.py)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
TAG=$1; N=$2; shift; shift
python tools/gpu/try_c4.py $N 2>&1 | tail -18
for K in "$@"; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 2 -c 1 -o gpurun_out/prof_${TAG}_$K -f python tools/gpu/try_c4.py $N > gpurun_out/ncu_${TAG}_$K.log 2>&1
tail -2 gpurun_out/ncu_${TAG}_$K.log
done
