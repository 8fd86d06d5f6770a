# usage: bash tools/gpu/run_bench_n.sh TAG N [extra bench args] -- one bench line at N GPUs
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
TAG=$1; N=$2; shift; shift
if [ "$N" = "1" ]; then
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 "$@" > gpurun_out/bench_${TAG}_n$N.json 2> gpurun_out/bench_${TAG}_n$N.err
else
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 "$@" > gpurun_out/bench_${TAG}_n$N.json 2> gpurun_out/bench_${TAG}_n$N.err
fi
echo "bench rc=$?"; grep -v "^\[W\|^W1\|^\*\*\*\|OMP_NUM" gpurun_out/bench_${TAG}_n$N.err | tail -12
python tools/show_bench.py gpurun_out/bench_${TAG}_n$N.json
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_${TAG}_n$N.json').read().strip().splitlines()[-1])
print('counters', d.get('space_counters')); print('eager', d.get('stages_one_eager_step_ms')); print('checks', d.get('checks')); print('kernels/step', d.get('kernels_per_step'), 'e2e', d['e2e'])
PY
