# usage: bash tools/gpu/run_launches.sh TAG -- ncu launch list (gpu__time_duration.sum) of one bench command
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
TAG=$1
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_$TAG.log 2>&1 && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$TAG.log 2>&1
tail -c 600 gpurun_out/plain_$TAG.log | head -c 300; wc -l gpurun_out/launches_$TAG.csv
