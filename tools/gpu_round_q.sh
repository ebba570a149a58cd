set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
STWO_B200_TREE_G=8 timeout 400 python -m pytest tests/test_gpu_verify.py -x -q > gpurun_out/pytest_g8.log 2>&1; tail -2 gpurun_out/pytest_g8.log
for n in 4096 512; do
timeout 200 python bench.py --steps 5 --warmup 3 --proofs $n --no-secondary --no-cpu-baseline > gpurun_out/bench_q_$n.json 2> gpurun_out/bench_q_$n.err; tail -2 gpurun_out/bench_q_$n.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_q_$n.json'))
s=d['roofline']['stage_ms']
print('n=$n value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],2), {k: round(v,2) for k,v in s.items()})
PY
done
