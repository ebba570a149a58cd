# tree group width at 4096 proofs with the 128-register kernels
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for G in 8 4 16; do for B in 8 10; do
  STWO_B200_TREE_G=$G STWO_B200_TREE_BLOCKS=$B timeout 300 python bench.py --steps 6 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/bench_z.json 2> gpurun_out/bench_z.err || tail -3 gpurun_out/bench_z.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_z.json'))
print('G $G blocks $B','value', round(d['value']), 'ms', round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['roofline']['stage_ms'].items() if k in ('single_tree','pair_tree','folds','fiat_shamir')})
PY
done; done
