cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python tools/level_clock.py > gpurun_out/level_clock_4096.json 2>/dev/null
python tools/level_clock.py --proofs 512 > gpurun_out/level_clock_512.json 2>/dev/null
for n in 4096 512; do
timeout 300 python bench.py --steps 6 --warmup 3 --proofs $n --no-secondary --no-cpu-baseline > gpurun_out/bench_s_$n.json 2> gpurun_out/bench_s_$n.err; tail -3 gpurun_out/bench_s_$n.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_s_$n.json'))
print('n=$n value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],2), 'eval', round(d['roofline']['stage_ms']['trace_eval'],2))
PY
done
