# usage: bash tools/gpu/run_r9k.sh TAG -- what the driver runs at round end (run_final.sh) + the other workloads at N = 1
cd $GRAFT_REPO_ROOT
bash tools/gpu/run_final.sh $1
for W in C5 C2 C4; do
  timeout 900 python bench.py --workload $W --steps 10 --warmup 3 > gpurun_out/bench_$1_$W.json 2> gpurun_out/bench_$1_$W.err; echo "$W rc=$?"; tail -1 gpurun_out/bench_$1_$W.err
  python tools/show_bench.py gpurun_out/bench_$1_$W.json | head -1
done
