set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
STWO_B200_TREE_G=16 timeout 300 python -m pytest tests/test_gpu_verify.py tests/test_gpu_circuit.py -x -q > gpurun_out/pytest_g16.log 2>&1; tail -2 gpurun_out/pytest_g16.log
STWO_B200_TREE_G=4 timeout 300 python -m pytest tests/test_gpu_verify.py tests/test_gpu_circuit.py -x -q > gpurun_out/pytest_g4.log 2>&1; tail -2 gpurun_out/pytest_g4.log
for n in 4096 512; do for g in 0 4 8 16; do
STWO_B200_TREE_G=$g timeout 200 python bench.py --steps 5 --warmup 3 --proofs $n --no-secondary --no-cpu-baseline > gpurun_out/bench_g${g}_$n.json 2> gpurun_out/bench_g${g}_$n.err; tail -2 gpurun_out/bench_g${g}_$n.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_g${g}_$n.json'))
s=d['roofline']['stage_ms']
print('G=$g n=$n value', round(d['value']), 'ms', round(d['ms_per_step'],2), 'single_tree', round(s['single_tree'],3), 'pair_tree', round(s['pair_tree'],3))
PY
done; done
