# usage: bash tools/gpu/run_c2.sh TAG -- parity tests, then the C2 (P2, 4096^2 triangles) and C5 (moving sphere) bench lines
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
TAG=${1:-x}
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py -m gpu -x -q 2>&1 | tail -8
for W in C2 C5; do
timeout 900 python bench.py --workload $W --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_$W.json 2> gpurun_out/bench_${TAG}_$W.err
python tools/show_bench.py gpurun_out/bench_${TAG}_$W.json
done
