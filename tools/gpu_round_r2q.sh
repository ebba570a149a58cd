# mixed 256-proof batch: submission order (largest group first / heaviest shape first) x worker-stream pools (1 / 6), three runs each
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for R in 1 2; do for O in size heavy; do for K in 1 6; do
  STWO_B200_MIXED_ORDER=$O STWO_B200_VERIFY_POOLS=$K python tools/multi_proofs_probe2.py 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('order $O pools $K', round(d['ms_per_batch'],2))"
done; done; done
timeout 600 python -m pytest tests/test_gpu_verify.py tests/test_gpu_chain.py -m gpu -x -q 2>&1 | tail -2
bash tools/gpu_stage.sh r2q
