# whole GPU suite + default bench (4096) + 512 / 1024 with the pipeline's own lane choice
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/bench_r2t.json 2> gpurun_out/bench_r2t.err || tail -5 gpurun_out/bench_r2t.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_r2t.json'))
print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['roofline']['stage_ms'].items() if v>0.05})
print('multi', json.dumps(d['secondary']['multi_proofs'])[-300:])
print('shape_R', json.dumps(d['secondary']['shape_R'])[:300])
print(d['config']['pipeline'])
PY
for P in 512 1024; do
  timeout 300 python bench.py --steps 12 --warmup 4 --proofs $P --no-secondary --no-cpu-baseline > gpurun_out/bench_r2t_$P.json 2> gpurun_out/bench_r2t_$P.err || tail -3 gpurun_out/bench_r2t_$P.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_r2t_$P.json'))
print('proofs $P','value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],2), d['config']['pipeline'][:40])
PY
done
