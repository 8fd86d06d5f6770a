# usage: bash tools/gpu/run_ncu2.sh TAG -- ncu --set full of kernels selected by their demangled names (template arguments)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
TAG=$1
i=0
for K in "rule_fill_kernel<3, 1>" "facet_p1_kernel" "cell_kernel<3, 1, 3, 1>" "compact_write_kernel<cfx::DnfPred>" "pattern_inactive_fill_kernel" "facet_rows_kernel"; do
i=$((i+1))
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$K" -s 2 -c 1 -o gpurun_out/prof_${TAG}_k$i -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_${TAG}_k$i.log 2>&1
tail -1 gpurun_out/ncu_${TAG}_k$i.log
done
