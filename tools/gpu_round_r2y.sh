# same box: input record on / off (the check then re-executes), stage times at 4096
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for R in 1 0 1 0; do
  STWO_B200_RECORD_INPUTS=$R timeout 300 python bench.py --steps 6 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/bench_y.json 2> gpurun_out/bench_y.err || tail -3 gpurun_out/bench_y.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_y.json'))
print('record_inputs $R','value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['roofline']['stage_ms'].items() if v>0.05})
PY
done
