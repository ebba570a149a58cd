# folding-stage circuit over synthetic instances: GPU tests, then the bench's synthetic leg
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_synth.py -x -q 2>&1 | tail -15
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2j.json 2> gpurun_out/bench_r2j.err || tail -5 gpurun_out/bench_r2j.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_r2j.json'))
print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],2))
print(json.dumps(d['secondary'].get('synthetic_4096'), indent=1))
PY
