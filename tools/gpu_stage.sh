# stage-timed bench at 4096 and 512: gpu_stage.sh <tag>
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for P in 4096 512; do
  timeout 300 python bench.py --steps 6 --warmup 3 --proofs $P --no-secondary --no-cpu-baseline > gpurun_out/bench_$1_$P.json 2> gpurun_out/bench_$1_$P.err || tail -3 gpurun_out/bench_$1_$P.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_$1_$P.json'))
print('$1 $P','value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['roofline']['stage_ms'].items() if v>0.05})
PY
done
