# bundle loop with operand forwarding: circuit parity tests, then stage-timed bench at 4096 / 512
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_circuit.py tests/test_gpu_synth.py -m gpu -x -q 2>&1 | tail -3
bash tools/gpu_stage.sh r2m
