# ncu --set full of one kernel of the trace pass (tools/trace_bench.py).  usage: gpu_ncu_trace.sh <kernel-regex> [proofs] [skip]
cd $GRAFT_REPO_ROOT
K=$1; N=${2:-4096}; S=${3:-1}
mkdir -p gpurun_out
python tools/trace_bench.py --proofs $N --reps 2 > gpurun_out/plain_$K.log 2>&1 || { tail -5 gpurun_out/plain_$K.log; exit 1; }
tail -1 gpurun_out/plain_$K.log
ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c 1 -f -o gpurun_out/prof_$K python tools/trace_bench.py --proofs $N --reps 1 > gpurun_out/ncu_$K.log 2>&1
tail -2 gpurun_out/ncu_$K.log
