set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python tools/level_clock.py > gpurun_out/level_clock_4096.json 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/ncu_launches.log 2>&1; tail -2 gpurun_out/ncu_launches.log
for K in k_tape_eval_grid k_cs_export_vals_tiled k_single_path k_pair_path; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 2 -c 1 -o gpurun_out/prof_$K -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/ncu_$K.log 2>&1; tail -2 gpurun_out/ncu_$K.log
done
