cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
TAG=$1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_deferred.py tests/test_parity_at_size.py -m gpu -x -q 2>&1 | tail -3
for B in 5 6 7 8; do
  export CFX_P1_MINB=$B
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_b$B.json 2> gpurun_out/bench_${TAG}_b$B.err; echo "minb $B rc=$?"
  python tools/show_bench.py gpurun_out/bench_${TAG}_b$B.json 2>/dev/null | grep "ms/step\|gather_matrix_p1"
done
