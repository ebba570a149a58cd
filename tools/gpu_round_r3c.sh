# flow record as 16-byte elements: parity tests (flow fetch, export_flow, check), stage times
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_circuit.py tests/test_gpu_synth.py tests/test_gpu_boundary.py tests/test_gpu_chain.py -m gpu -x -q 2>&1 | tail -3
bash tools/gpu_stage.sh r3c
