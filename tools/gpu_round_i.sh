set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_circuit.py -x -q > gpurun_out/pytest_circuit.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_circuit.log; tail -5 gpurun_out/pytest_circuit.log
for mode in cta grid; do for n in 512 4096; do
STWO_B200_EVAL_MODE=$mode timeout 200 python tools/trace_bench.py --proofs $n > gpurun_out/trace_${mode}_$n.json 2> gpurun_out/trace_${mode}_$n.err; tail -2 gpurun_out/trace_${mode}_$n.err; cat gpurun_out/trace_${mode}_$n.json
done; done
timeout 200 python tools/trace_bench.py --proofs 4736 > gpurun_out/trace_auto_4736.json 2>&1; cat gpurun_out/trace_auto_4736.json
