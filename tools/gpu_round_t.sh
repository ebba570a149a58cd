cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for c in 0 2 3 4; do
for n in 4096 512; do
STWO_B200_BESIDE_CTAS=$c timeout 300 python tools/trace_bench.py --proofs $n --reps 5 > gpurun_out/trace_t_$n.json 2> gpurun_out/trace_t_$n.err; tail -3 gpurun_out/trace_t_$n.err
python - <<PY
import json
d=json.load(open('gpurun_out/trace_t_$n.json'))
print('ctas=$c n=$n', {k: round(v,2) for k,v in d['trace_stage_ms'].items()}, 'sum', round(sum(d['trace_stage_ms'].values()),2), 'untimed', round(d['trace_untimed_ms'],2), 'v+t', round(d['verify_plus_trace_ms'],2))
PY
done
done
