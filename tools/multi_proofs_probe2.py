"""configs[3] leg alone (1 GPU or under torchrun): python tools/multi_proofs_probe2.py"""
import importlib, json, os, sys
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
pkg = importlib.import_module("recursive-stwo_b200"); sharding = importlib.import_module("recursive-stwo_b200.sharding")
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1: dist.init_process_group("nccl", device_id=dev)
pkg.init(local)
out = bench.multi_proofs_leg(pkg, sharding, rank, world, dev)
if rank == 0: print(json.dumps(out))
if world > 1: dist.destroy_process_group()
