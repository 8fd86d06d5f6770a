# usage: bash tools/gpu/run_band.sh TAG -- band-row gather with G = 2,4,8,16 threads per row: C3 and C5 bench lines
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
TAG=$1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py tests/test_parity_at_size.py -m gpu -q -x 2>&1 | tail -3
for G in 2 4 8 16; do for W in C3 C5; do
CFX_BAND_G=$G timeout 600 python bench.py --workload $W --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_${W}_g$G.json 2>/dev/null
echo "G=$G $W"; python tools/show_bench.py gpurun_out/bench_${TAG}_${W}_g$G.json | grep "ms/step\|mask_kernel"
done; done
