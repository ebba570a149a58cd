# second tape order (recorded permutations split): parity tests, stage times, the heavy group, A/B against STWO_B200_RECORDED_ORDER=0
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_circuit.py tests/test_gpu_synth.py tests/test_gpu_chain.py tests/test_gpu_pipeline.py -m gpu -x -q 2>&1 | tail -4
for R in 1 0; do
  STWO_B200_RECORDED_ORDER=$R bash tools/gpu_stage.sh r3a_o$R
  STWO_B200_RECORDED_ORDER=$R python tools/heavy_group_probe.py 2>&1 | tail -1 | cut -c150-600
done
