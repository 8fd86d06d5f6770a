# usage: bash tools/gpu/run_r5c.sh TAG NGPU -- multi-GPU session: static-plan tests (emulated + NCCL graph), bench at N
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
TAG=${1:-r5c}; N=${2:-2}
nvidia-smi --query-gpu=name --format=csv,noheader | head -8; nvidia-smi topo -m 2>/dev/null | head -12
echo "== parallel tests"
timeout 1200 python -m pytest tests/test_parallel_gpu.py -m gpu -q -x > gpurun_out/${TAG}_parallel.log 2>&1; grep -E "^E  |passed|failed|skipped|test_parallel_gpu.py:[0-9]+: in|Error" gpurun_out/${TAG}_parallel.log | head -30
echo "== bench N=$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${TAG}_n$N.json 2> gpurun_out/bench_${TAG}_n$N.err
echo "bench rc=$?"; grep -v "^\[W\|^W1\|^\*\*\*\|OMP_NUM" gpurun_out/bench_${TAG}_n$N.err | tail -12
python tools/show_bench.py gpurun_out/bench_${TAG}_n$N.json
python - <<'PY'
import json,sys,glob
for f in sorted(glob.glob('gpurun_out/bench_*_n*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['n_gpus'], round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), d.get('checks'))
    except Exception as e: print(f, 'unreadable', e)
PY
