"""N > 1 check of the C-ABI multi-GPU entry points (run under torchrun, one rank per GPU): the library's own NCCL communicator
(stwo_b200_comm_*), stwo_b200_gather_verdicts and stwo_b200_gather_trace_columns against torch.distributed's collectives, on a
sharded batch of real proofs with tampered ones mixed in.  Prints one JSON line on rank 0."""
import importlib
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("recursive-stwo_b200")
sharding = importlib.import_module("recursive-stwo_b200.sharding")

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
pkg.init(local)
comm = sharding.Comm(rank, world, dev)

blob = bytearray(open(os.path.join(ROOT, "tests", "golden", "proofs", "small_proof.bin"), "rb").read())
n_total = 64 * world + 3                                 # ragged blocks
bad = {p for p in range(n_total) if p % 11 == 4}
lo, hi = sharding.shard_range(n_total, rank, world)
blobs = []
for p in range(lo, hi):
    b = bytearray(blob)
    if p in bad:
        b[30000 + p % 64] ^= 1
    blobs.append(bytes(b))
vb = pkg.VerifyBatch(blobs, inputs=pkg.INPUTS_SINGLE)
v, s = vb.run(full=True)
circ = pkg.VerifierCircuit(vb.shape, inputs=pkg.INPUTS_SINGLE)
r = circ.trace(vb, check=True, export=True, preprocessed=False)
gv, gs = comm.gather_verdicts(v, s, n_total)
tv, ts = sharding.gather_verdicts(v, s, n_total)
ok = bool(torch.equal(gv, tv)) and bool(torch.equal(gs, ts))
want = np.array([1 if p in bad else 0 for p in range(n_total)], dtype=np.uint8)
ok = ok and np.array_equal(gv.cpu().numpy() != 0, want != 0)
vals = r["values"][:, :, :4096].contiguous()            # a slice of the columns keeps the check quick
gc = comm.gather_trace_columns(vals, n_total, dst=0)
gt = sharding.gather_trace_columns(vals, n_total, dst=0)
if rank == 0:
    ok = ok and gc.shape == gt.shape and bool(torch.equal(gc, gt))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(json.dumps({"world": world, "n_total": n_total, "c_entry_gathers_match_torch_distributed": bool(flag.item()), "rejected": int(want.sum())}))
comm.close()
dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)
