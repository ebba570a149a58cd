# one full round on a 1-GPU box: every gpu test, smoke, both bench arms, small batch, launch list of one bench step
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 300 python bench.py --steps 10 --warmup 3 --proofs 512 --no-secondary --no-cpu-baseline > gpurun_out/bench_512.json 2> gpurun_out/bench_512.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/ncu_launches.log 2>&1
python - <<PY
import json
for f in ('bench','bench_512'):
    d=json.load(open('gpurun_out/%s.json'%f))
    print(f,'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],2), 'launches', d['gpu_launches'], {k:round(v,2) for k,v in d['roofline']['stage_ms'].items() if v>0.05})
    print(' roofline', {k:d['roofline'][k] for k in ('kernel','bound','achieved','peak','frac','traffic') if k in d['roofline']})
    if 'cpu_baseline' in d: print(' cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
    for k,v in d.get('secondary',{}).items():
        if isinstance(v,dict): print(' ',k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items() if not isinstance(b,(dict,list,str))})
d=json.load(open('gpurun_out/bench_ref.json')); print('ref', d['value'], d['cpu_baseline']['cores'])
PY
