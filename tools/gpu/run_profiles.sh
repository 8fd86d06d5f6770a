# usage: bash tools/gpu/run_profiles.sh TAG -- bench (plain), ncu launch list of the same command, one --set full capture
# of each of the three largest kernels of the step (each ncu pass only after the plain run exited 0)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
TAG=$1
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err || exit 1
python tools/show_bench.py gpurun_out/bench_$TAG.json
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_$TAG.log 2>&1
for K in ${KERNELS:-gather_matrix_p1_kernel gather_matrix_band_p1_kernel pattern_rows_kernel pattern_static_kernel classify_kernel}; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 2 -c 1 -o gpurun_out/prof_${TAG}_$K -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_${TAG}_$K.log 2>&1
tail -1 gpurun_out/ncu_${TAG}_$K.log
done
