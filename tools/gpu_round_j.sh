set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for u in 0 1; do for n in 4096 4736; do
STWO_B200_EVAL_UNROLLED=$u timeout 200 python tools/trace_bench.py --proofs $n > gpurun_out/trace_u${u}_$n.json 2> gpurun_out/trace_u${u}_$n.err; tail -2 gpurun_out/trace_u${u}_$n.err; python -c "
import json; d=json.load(open('gpurun_out/trace_u${u}_$n.json')); print('unrolled=$u n=$n', d['trace_stage_ms'], d['proofs_per_sec_verify_plus_trace'])"
done; done
timeout 300 python -m pytest tests/test_gpu_circuit.py -x -q 2>&1 | tail -2
