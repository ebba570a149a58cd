"""How fast can this B200 WRITE?  The trace export moves 3.6-4 GB of reads and 13.9 GB of writes per launch; MEASURED_PEAKS' hbm_gbs is a
copy (half reads, half writes).  Times torch fill_ (write only), copy_ (1:1) and a 1:3.5 read:write mix on 8 GiB buffers."""
import json, torch
dev = torch.device("cuda:0")
n = 1 << 31                                   # 2 Gi int32 = 8 GiB
a = torch.empty(n, dtype=torch.int32, device=dev)
b = torch.empty(n, dtype=torch.int32, device=dev)
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
out = {}
ms = timed(lambda: a.fill_(7)); out["fill_write_only_gbs"] = n * 4 / ms / 1e6
ms = timed(lambda: torch.cuda.memset if False else a.zero_()); out["zero_write_only_gbs"] = n * 4 / ms / 1e6
ms = timed(lambda: b.copy_(a)); out["copy_1r_1w_gbs"] = 2 * n * 4 / ms / 1e6
# 1 read : 3.5 writes -- read a quarter-size source, write it out 3.5 times (expand + copy into a [7, n/8] view of b... use 2 reads : 7 writes)
src = a[: n // 8]
dst = b[: 7 * (n // 16)].view(7, n // 16)
s2 = a[: n // 16]
ms = timed(lambda: dst.copy_(s2.unsqueeze(0).expand(7, -1))); out["expand_1r_7w_gbs_dram"] = (n // 16 * 4 + 7 * (n // 16) * 4) / ms / 1e6
ms = timed(lambda: a.sum()); out["read_only_gbs"] = n * 4 / ms / 1e6
print(json.dumps(out))
