#!/bin/bash
# One GPU visit: tests, smoke, microbench, K1 shapes, bench, then the two ncu passes (launch list, full set on the
# dominant kernel).  Everything lands in gpurun_out/.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
./build/intpipe gpurun_out/intpipe.json > gpurun_out/intpipe.log 2>&1; tail -20 gpurun_out/intpipe.log
python tools/k1_variants.py > gpurun_out/k1_variants.json 2> gpurun_out/k1_variants.err; cat gpurun_out/k1_variants.json
python tools/verify_bench.py small_proof.bin 4096 > gpurun_out/verify_small.json 2> gpurun_out/verify_small.err; cat gpurun_out/verify_small.json; tail -3 gpurun_out/verify_small.err
python tools/verify_bench.py recursive_proof_16_15.bin 2048 > gpurun_out/verify_rec.json 2> gpurun_out/verify_rec.err; cat gpurun_out/verify_rec.json; tail -3 gpurun_out/verify_rec.err
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
if [ "$1" = "ncu" ]; then
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --trees 1 > gpurun_out/plain2.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:k_commit_leaves -s 1 -c 1 -o gpurun_out/prof_leaves \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --trees 1 > gpurun_out/ncu_full.log 2>&1
  ls -la gpurun_out
fi
