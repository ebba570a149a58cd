set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 200 python tools/trace_bench.py --proofs 256 --fixture level13-1.bin --last-layer > gpurun_out/trace_last_256.json 2> gpurun_out/trace_last_256.err; tail -2 gpurun_out/trace_last_256.err
timeout 200 python tools/trace_bench.py --proofs 1 --fixture level13-1.bin --last-layer > gpurun_out/trace_last_1.json 2> gpurun_out/trace_last_1.err
timeout 200 python tools/trace_bench.py --proofs 2048 --fixture level13-1.bin --last-layer > gpurun_out/trace_last_2048.json 2> gpurun_out/trace_last_2048.err
timeout 200 python tools/trace_bench.py --proofs 256 --fixture recursive_proof_16_15.bin > gpurun_out/trace_rec_256.json 2> gpurun_out/trace_rec_256.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/ncu_launches.log 2>&1; tail -2 gpurun_out/ncu_launches.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tape_eval_grid -c 1 -o gpurun_out/prof_k_tape_eval_grid -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/ncu_k_tape_eval_grid.log 2>&1; tail -2 gpurun_out/ncu_k_tape_eval_grid.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_cs_export_vals_tiled -c 1 -o gpurun_out/prof_k_export_fused -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/ncu_k_export_fused.log 2>&1; tail -2 gpurun_out/ncu_k_export_fused.log
