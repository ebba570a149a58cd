# multi-GPU lines: gpu_round_scale2.sh N   (under gpurun --gpus N): weak (4096 per GPU), strong (4096 in total), the mixed 256-proof batch, the C-entry gathers
cd $GRAFT_REPO_ROOT
N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29521 bench.py --gpus $N --steps 8 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/scale_weak_$N.json 2> gpurun_out/scale_weak_$N.err
$TR --master-port 29522 bench.py --gpus $N --steps 8 --warmup 3 --scaling strong --proofs 4096 --no-secondary --no-cpu-baseline > gpurun_out/scale_strong_$N.json 2> gpurun_out/scale_strong_$N.err
$TR --master-port 29523 tools/multi_proofs_probe2.py > gpurun_out/scale_mixed_$N.json 2> gpurun_out/scale_mixed_$N.err
$TR --master-port 29524 tools/multi_gpu_check.py > gpurun_out/multi_check_$N.json 2> gpurun_out/multi_check_$N.err
python - <<PY
import json
for f in ('scale_weak_$N','scale_strong_$N'):
    try:
        d=json.load(open('gpurun_out/%s.json'%f)); print(f,'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],2), d['config']['proofs_per_gpu'])
    except Exception as e: print(f,'FAILED',e)
for f in ('scale_mixed_$N','multi_check_$N'):
    try: print(f, open('gpurun_out/%s.json'%f).read().strip()[-260:])
    except Exception as e: print(f,'FAILED',e)
PY
