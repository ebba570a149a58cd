# per-caller-stream pools in the verify entry: regression tests, then pipeline lanes 1 / 2 / 3 at 512, 1024 and 4096 proofs
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_verify.py tests/test_gpu_boundary.py -m gpu -x -q 2>&1 | tail -3
for L in 1 2 3; do for P in 512 1024 4096; do
  if [ $L = 3 ] && [ $P = 4096 ]; then continue; fi
  timeout 300 python bench.py --steps 12 --warmup 4 --proofs $P --lanes $L --no-secondary --no-cpu-baseline > gpurun_out/bench_l${L}_$P.json 2> gpurun_out/bench_l${L}_$P.err || tail -3 gpurun_out/bench_l${L}_$P.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_l${L}_$P.json'))
print('lanes $L proofs $P','value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],2))
PY
done; done
