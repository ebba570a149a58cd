# multi-GPU round: usage gpu_round_multi.sh N   (under gpurun --gpus N)
cd $GRAFT_REPO_ROOT
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py > gpurun_out/multi_check_$N.json 2> gpurun_out/multi_check_$N.err; tail -1 gpurun_out/multi_check_$N.json; tail -2 gpurun_out/multi_check_$N.err
