# tree rebuild kernels under register targets: 152 registers / 6 blocks (1), 128 / 8 blocks (8), 96 / 10 blocks (10)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for B in 1 8 10; do
  STWO_B200_TREE_BLOCKS=$B bash tools/gpu_stage.sh r2u_b$B
done
STWO_B200_TREE_BLOCKS=8 timeout 600 python -m pytest tests/test_gpu_verify.py tests/test_gpu_synth.py -m gpu -x -q 2>&1 | tail -2
STWO_B200_TREE_BLOCKS=10 timeout 600 python -m pytest tests/test_gpu_verify.py tests/test_gpu_synth.py -m gpu -x -q 2>&1 | tail -2
