# A/B of the trace-export kernel: parity first, then the stage-timed bench per variant
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_circuit.py -m gpu -x -q 2>&1 | tail -3
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 6 --warmup 3 --proofs ${PROOFS:-4096} --no-secondary --no-cpu-baseline > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err || tail -3 gpurun_out/bench_$name.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_$name.json'))
print('$name','value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['roofline']['stage_ms'].items() if v>0.05})
PY
}
run stream2 X=1
run tiled32 STWO_B200_EXPORT_ITEMS=32
PROOFS=512 run stream2_512 X=1
bash tools/gpu_ncu_trace.sh k_cs_export_vals_stream 4096 1
