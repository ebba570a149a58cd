cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for c in 32 16; do
for n in 4096 512; do
STWO_B200_EXPORT_ITEMS=$c timeout 300 python tools/trace_bench.py --proofs $n --reps 5 > gpurun_out/trace_t_$n.json 2> gpurun_out/trace_t_$n.err; tail -3 gpurun_out/trace_t_$n.err
python - <<PY
import json
d=json.load(open('gpurun_out/trace_t_$n.json'))
print('items=$c n=$n', {k: round(v,2) for k,v in d['trace_stage_ms'].items()}, 'untimed', round(d['trace_untimed_ms'],2))
PY
done
done
timeout 600 python -m pytest tests/test_gpu_circuit.py -m gpu -x -q 2>&1 | tail -2
python tools/level_clock.py > gpurun_out/level_clock_4096.json 2>/dev/null
python tools/level_clock.py --proofs 512 > gpurun_out/level_clock_512.json 2>/dev/null
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
