# usage: bash tools/gpu/run_workloads.sh TAG [WORKLOADS] -- one bench line per workload (default C5 C2 C4), with the CPU arm beside it
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
TAG=$1; shift
for W in ${@:-C5 C2 C4}; do
timeout 1500 python bench.py --workload $W --steps 5 --warmup 3 > gpurun_out/bench_${TAG}_$W.json 2> gpurun_out/bench_${TAG}_$W.err
echo "bench $W rc=$?"; tail -3 gpurun_out/bench_${TAG}_$W.err
python tools/show_bench.py gpurun_out/bench_${TAG}_$W.json
done
