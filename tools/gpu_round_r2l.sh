# 512-proof batches with two pipeline lanes: throughput-oriented kernel shapes (tree group width, evaluation cluster size) against the latency-oriented defaults
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_verify.py -m gpu -x -q -k "oods or fixture or tamper" 2>&1 | tail -2
for L in 1 2; do for G in 8 16; do for C in 2 4 8; do
  STWO_B200_TREE_G=$G STWO_B200_EVAL_CLUSTER=$C timeout 300 python bench.py --steps 12 --warmup 4 --proofs 512 --lanes $L --no-secondary --no-cpu-baseline > gpurun_out/bench_m.json 2> gpurun_out/bench_m.err || tail -3 gpurun_out/bench_m.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_m.json'))
print('lanes $L G $G cluster $C','value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['roofline']['stage_ms'].items() if v>0.05})
PY
done; done; done
