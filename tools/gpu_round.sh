#!/bin/bash
# One GPU visit: tests, smoke, bench (both arms), then the two ncu passes (launch list, full set on the dominant kernel).
# Everything lands in gpurun_out/.   usage: tools/gpu_round.sh [ncu <kernel-regex>]
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
python tools/verify_bench.py recursive_proof_16_15.bin 2048 > gpurun_out/verify_rec.json 2> gpurun_out/verify_rec.err; cat gpurun_out/verify_rec.json; tail -3 gpurun_out/verify_rec.err
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
if [ "$1" = "ncu" ]; then
  K=${2:-k_single_path}
  python bench.py --steps 1 --warmup 3 --proofs 1024 --no-cpu-baseline --no-secondary > gpurun_out/plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 1 --warmup 3 --proofs 1024 --no-cpu-baseline --no-secondary > gpurun_out/ncu_launches.log 2>&1
  python bench.py --steps 1 --warmup 3 --proofs 1024 --no-cpu-baseline --no-secondary > gpurun_out/plain2.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$K -s 2 -c 1 -o gpurun_out/prof_$K \
      python bench.py --steps 1 --warmup 3 --proofs 1024 --no-cpu-baseline --no-secondary > gpurun_out/ncu_full.log 2>&1
  ls -la gpurun_out
fi
