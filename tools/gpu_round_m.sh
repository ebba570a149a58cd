set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
for n in 1 256 2048; do timeout 200 python tools/trace_bench.py --proofs $n --fixture level13-1.bin --last-layer > gpurun_out/trace_last_$n.json 2> gpurun_out/trace_last_$n.err; tail -2 gpurun_out/trace_last_$n.err; done
timeout 200 python tools/trace_bench.py --proofs 4096 > gpurun_out/trace_4096.json 2> gpurun_out/trace_4096.err
