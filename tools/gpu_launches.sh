# per-kernel durations (ncu launch list) of verify+trace at a small batch: gpu_launches.sh <fixture> <proofs> <tag>
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_$3.csv python tools/trace_bench.py --fixture $1 --proofs $2 --reps 1 > gpurun_out/launches_$3.log 2>&1
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/launches_$3.csv')) if len(r)>10]
hdr=rows[0]; ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value')
d=collections.OrderedDict()
for r in rows[1:]:
    k=r[ik].split('(')[0][-40:]; d.setdefault(k,[]).append(float(r[iv].replace(',','')))
for k,v in d.items(): print('%-42s n=%3d last=%9.1f us  min=%9.1f' % (k, len(v), v[-1]/1e3, min(v)/1e3))
PY
