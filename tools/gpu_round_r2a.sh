# round 2, first GPU call: every gpu test, smoke, the stage-timed bench at 4096 and 512
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 300 python bench.py --steps 6 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/bench_st.json 2> gpurun_out/bench_st.err; tail -2 gpurun_out/bench_st.err
timeout 300 python bench.py --steps 6 --warmup 3 --proofs 512 --no-secondary --no-cpu-baseline > gpurun_out/bench_st512.json 2> gpurun_out/bench_st512.err; tail -2 gpurun_out/bench_st512.err
python - <<PY
import json
for f in ('bench_st','bench_st512'):
    d=json.load(open('gpurun_out/%s.json'%f))
    print(f,'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['roofline']['stage_ms'].items()})
PY
