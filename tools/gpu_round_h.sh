set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -5 gpurun_out/bench.err; cat gpurun_out/bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary --proofs 2048 > gpurun_out/ncu_launches.log 2>&1; tail -2 gpurun_out/ncu_launches.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tape_eval -c 1 -o gpurun_out/prof_k_tape_eval -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary --proofs 2048 > gpurun_out/ncu_k_tape_eval.log 2>&1; tail -2 gpurun_out/ncu_k_tape_eval.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_cs_export_vals_tiled -c 1 -o gpurun_out/prof_k_export -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary --proofs 2048 > gpurun_out/ncu_k_export.log 2>&1; tail -2 gpurun_out/ncu_k_export.log
