# record-based check_poseidon_invocations: parity tests, then stage-timed bench
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_circuit.py tests/test_gpu_verify.py tests/test_gpu_synth.py tests/test_gpu_chain.py -m gpu -x -q 2>&1 | tail -8
bash tools/gpu_stage.sh r2x
