cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v.json 2> gpurun_out/bench_v.err; tail -3 gpurun_out/bench_v.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_v.json'))
print('value', round(d['value']), 'e2e', round(d['e2e']['value']))
print(json.dumps(d['secondary']['multi_proofs']))
PY
