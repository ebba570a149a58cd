cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_verify.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 6 --warmup 3 --proofs ${PROOFS:-4096} --no-secondary --no-cpu-baseline > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err || tail -3 gpurun_out/bench_$name.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_$name.json'))
print('$name','value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['roofline']['stage_ms'].items() if v>0.05})
PY
}
run default X=1
PROOFS=512 run default_512 X=1


PROOFS=1024 run default_1024 X=1
