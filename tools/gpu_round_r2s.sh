# 32 hardware queues by default: pipeline lanes 1..4 at 512 / 1024 / 2048 / 4096 proofs
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for P in 512 1024 2048 4096; do for L in 1 2 3 4; do
  if [ $P = 4096 ] && [ $L -gt 2 ]; then continue; fi
  timeout 300 python bench.py --steps 12 --warmup 4 --proofs $P --lanes $L --no-secondary --no-cpu-baseline > gpurun_out/bench_s.json 2> gpurun_out/bench_s.err || tail -3 gpurun_out/bench_s.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_s.json'))
print('lanes $L proofs $P','value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],2))
PY
done; done
