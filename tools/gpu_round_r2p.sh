# whole GPU suite + default bench + per-level clocks of the cluster evaluation on the current tree
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/bench_r2p.json 2> gpurun_out/bench_r2p.err || tail -5 gpurun_out/bench_r2p.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_r2p.json'))
print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['roofline']['stage_ms'].items() if v>0.05})
print('multi', json.dumps(d['secondary']['multi_proofs'])[-420:])
print('shape_R', json.dumps(d['secondary']['shape_R'])[:300])
PY
for P in 512 4096; do python tools/level_clock.py --proofs $P > gpurun_out/level_clock_cluster_$P.json 2>> gpurun_out/level_clock.err; done
