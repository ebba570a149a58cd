# hardware work queues: CUDA_DEVICE_MAX_CONNECTIONS 8 (default) against 32, for the mixed batch (11 shape groups x worker streams) and the pipeline
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for C in 8 32; do for O in size heavy; do
  CUDA_DEVICE_MAX_CONNECTIONS=$C STWO_B200_MIXED_ORDER=$O python tools/multi_proofs_probe2.py 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('connections $C order $O', round(d['ms_per_batch'],2))"
done; done
for C in 8 32; do for L in 1 2; do for P in 512 4096; do
  CUDA_DEVICE_MAX_CONNECTIONS=$C timeout 300 python bench.py --steps 12 --warmup 4 --proofs $P --lanes $L --no-secondary --no-cpu-baseline > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err || tail -3 gpurun_out/bench_c.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_c.json'))
print('connections $C lanes $L proofs $P','value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],2), 'fs', round(d['roofline']['stage_ms']['fiat_shamir'],3))
PY
done; done; done
timeout 600 python -m pytest tests/test_gpu_verify.py -m gpu -x -q 2>&1 | tail -2
