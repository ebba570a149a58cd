# cluster evaluation: plain loop at 64 registers against the forwarding loop at 120 (one CTA per SM), 512 / 1024 / 4096 proofs
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for F in 0 1; do for P in 512 1024 4096; do
  STWO_B200_EVAL_FORWARD=$F timeout 300 python bench.py --steps 6 --warmup 3 --proofs $P --no-secondary --no-cpu-baseline > gpurun_out/bench_o.json 2> gpurun_out/bench_o.err || tail -3 gpurun_out/bench_o.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_o.json'))
print('fwd $F proofs $P','value', round(d['value']), 'ms', round(d['ms_per_step'],2), 'eval', round(d['roofline']['stage_ms']['trace_eval'],3))
PY
done; done
STWO_B200_EVAL_FORWARD=1 python tools/level_clock.py --proofs 512 > gpurun_out/level_clock_fwd120_512.json
timeout 900 python -m pytest tests/test_gpu_circuit.py -m gpu -x -q 2>&1 | tail -2
