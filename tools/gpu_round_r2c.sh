cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_circuit.py -m gpu -x -q 2>&1 | tail -3
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 6 --warmup 3 --proofs ${PROOFS:-4096} --no-secondary --no-cpu-baseline > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err || tail -3 gpurun_out/bench_$name.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_$name.json'))
print('$name','value', round(d['value']), 'ms', round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['roofline']['stage_ms'].items() if k in ('trace_eval',)})
PY
}
run default X=1
run cl3 STWO_B200_EVAL_CLUSTER=3
run grid STWO_B200_EVAL_MODE=grid
PROOFS=512 run default_512 X=1
PROOFS=512 run cl4_512 STWO_B200_EVAL_CLUSTER=4
PROOFS=1024 run default_1024 X=1
PROOFS=2048 run default_2048 X=1
python tools/mixed_breakdown.py 2>&1 | tail -10 | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print(d['shape'], d['n'], d['ms'], d['trace'].get('eval'))"
python tools/multi_proofs_probe2.py 2>&1 | tail -1 | cut -c 200-500
