# vectorised record-slot / permutation-record loads in the tape evaluation: parity tests + stage times
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_circuit.py tests/test_gpu_synth.py -m gpu -x -q 2>&1 | tail -3
bash tools/gpu_stage.sh r3b
