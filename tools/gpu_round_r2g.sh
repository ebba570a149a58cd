cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
STWO_B200_BUNDLE=16 timeout 900 python -m pytest tests/test_gpu_circuit.py tests/test_gpu_chain.py -m gpu -x -q 2>&1 | tail -2
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 6 --warmup 3 --proofs ${PROOFS:-4096} --no-secondary --no-cpu-baseline > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err || tail -3 gpurun_out/bench_$name.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_$name.json'))
print('$name','value', round(d['value']), 'ms', round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['roofline']['stage_ms'].items() if k in ('trace_eval',)})
PY
}
for B in 1 8 16 32; do run b$B STWO_B200_BUNDLE=$B; PROOFS=512 run b${B}_512 STWO_B200_BUNDLE=$B; done
for B in 8 16 32; do STWO_B200_BUNDLE=$B python tools/multi_proofs_probe2.py 2>&1 | tail -1 | cut -c 230-330; done
