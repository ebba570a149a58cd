# permutation check beside the export (STWO_B200_BESIDE_CTAS) with 32 work queues, 4096 proofs
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for K in 0 2 3 4 6; do
  STWO_B200_BESIDE_CTAS=$K timeout 300 python bench.py --steps 10 --warmup 4 --no-secondary --no-cpu-baseline > gpurun_out/bench_w.json 2> gpurun_out/bench_w.err || tail -3 gpurun_out/bench_w.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_w.json'))
print('beside $K','value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],2))
PY
done
