cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
CUDA_LAUNCH_BLOCKING=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py -m gpu -x -q 2>&1 | tail -40
