set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for g in 4 8; do STWO_B200_TREE_G=$g timeout 400 python -m pytest tests/test_gpu_verify.py -x -q > gpurun_out/pytest_g$g.log 2>&1; tail -2 gpurun_out/pytest_g$g.log; done
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
