cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_circuit.py -m gpu -x -q 2>&1 | tail -1
timeout 300 python bench.py --steps 6 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/bench_st.json 2> gpurun_out/bench_st.err; tail -2 gpurun_out/bench_st.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_st.json'))
print('value', round(d['value']), 'ms', round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['roofline']['stage_ms'].items()})
PY
