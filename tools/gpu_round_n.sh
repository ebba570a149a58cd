set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; tail -3 gpurun_out/bench_2gpu.err; cut -c1-400 gpurun_out/bench_2gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --proofs 512 --no-secondary > gpurun_out/bench_2gpu_512.json 2> gpurun_out/bench_2gpu_512.err; cut -c1-300 gpurun_out/bench_2gpu_512.json
python bench.py --steps 5 --warmup 3 --proofs 512 --no-secondary --no-cpu-baseline > gpurun_out/bench_1gpu_512.json 2> gpurun_out/bench_1gpu_512.err; cut -c1-300 gpurun_out/bench_1gpu_512.json
