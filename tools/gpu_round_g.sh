set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python tools/trace_bench.py --proofs 4096 > gpurun_out/trace_4096.json 2> gpurun_out/trace_4096.err; tail -3 gpurun_out/trace_4096.err; cat gpurun_out/trace_4096.json
timeout 300 python tools/trace_bench.py --proofs 512 --fixture level1-5.bin > gpurun_out/trace_l15.json 2> gpurun_out/trace_l15.err; tail -3 gpurun_out/trace_l15.err; cat gpurun_out/trace_l15.json
