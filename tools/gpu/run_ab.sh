# usage: bash tools/gpu/run_ab.sh TAG ENVVAR [WORKLOADS] -- bench lines with and without an A/B environment switch
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
TAG=$1; VAR=$2; shift; shift
for W in ${@:-C3 C2}; do
for MODE in new old; do
if [ $MODE = old ]; then export $VAR=1; else unset $VAR; fi
timeout 600 python bench.py --workload $W --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_${W}_$MODE.json 2>/dev/null
echo "$W $MODE"; python tools/show_bench.py gpurun_out/bench_${TAG}_${W}_$MODE.json | grep "ms/step\|create_sparsity\|_kernel"
done; done
