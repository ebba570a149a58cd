# cluster evaluation: per-level clocks and stage times, plain bundle loop against the forwarding one
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for F in 0 1; do
  for P in 512 4096; do
    STWO_B200_EVAL_FORWARD=$F python tools/level_clock.py --proofs $P > gpurun_out/level_clock_f${F}_$P.json 2> gpurun_out/level_clock.err || tail -3 gpurun_out/level_clock.err
  done
  STWO_B200_EVAL_FORWARD=$F bash tools/gpu_stage.sh r2n_f$F
done
