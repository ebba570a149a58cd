#!/usr/bin/env python3
"""bench.py — headline benchmark of the B200-native recursive-stwo verifier hot path.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
    python bench.py --impl reference ...                      (the CPU oracle port on the host cores)

Workload (BASELINE.json configs[2]): synthetic Poseidon31 Merkle decommitment sweep — per GPU,
T trees of 2^20 leaves x C M31 columns (splitmix64(seed = tree id)), committed with Poseidon2-M31,
Q queries per tree decommitted and every authentication path recomputed and compared with the
root.  One "step" = commit + decommit + verify of all T trees.  Metric: Poseidon2 permutations/s
(whole job, all GPUs).  `value` has inputs resident in HBM; `e2e` goes through the host-pointer
C ABI (pinned host columns / paths -> device -> roots / verdicts back) every step.
Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "poseidon31_perms_per_sec"
UNIT = "perms/s"
# executed integer lane-instructions per permutation of the rolled kernel shape (SASS dynamic count:
# 5 ext-MDS + 4*... see DESIGN.md §roofline) and the measured issue peak (profiles/intpipe_r01.json)
LANE_OPS_PER_PERM = 4264
INT_PEAK_TLOPS_FALLBACK = 30.9


def peaks():
    p = {"hbm_gbs": 6650.0, "src": "fallback"}
    try:
        m = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        p = {"hbm_gbs": float(m["hbm_gbs"]), "src": "measured"}
    except Exception:
        pass
    try:
        ip = json.load(open(os.path.join(ROOT, "profiles", "intpipe_r01.json")))
        p["int_tlops"] = float(ip["peak_tera_lane_ops_per_s"])
        p["int_src"] = "measured (profiles/intpipe_r01.json)"
    except Exception:
        p["int_tlops"] = INT_PEAK_TLOPS_FALLBACK
        p["int_src"] = "fallback"
    return p


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [int(r[0]) for r in self.rows if r and r[0].isdigit()]
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": int(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def perms_per_tree(log_n, n_cols, n_q):
    n = 1 << log_n
    leaf = (n_cols + 7) // 8 + 1
    return n * leaf + (n - 1) + n_q * (leaf + log_n)


# ----------------------------------------------------------------------------------------------------
def run_reference(args):
    """The reference's CPU implementation of the path = the oracle port (the Rust workspace cannot be built here:
    no cargo, stwo git dependency absent).  Bounded sample of the same workload on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle_py as O
    lib = O.load_oracle()
    lib.orc_merkle_build_mt.restype = ctypes.c_uint64
    cores = os.cpu_count() or 1
    log_n, C, Q = args.ref_log_n, args.cols, args.queries
    n = 1 << log_n
    leaves = O.synth_m31(0, n * C).reshape(n, C)
    nodes = np.zeros((2 * n - 1, 8), dtype=np.uint32)
    idx = (O.splitmix64(0xABCDEF, Q) & np.uint64(n - 1)).astype(np.uint32)
    ncols = np.zeros(log_n + 1, dtype=np.uint32)
    ncols[log_n] = C
    verdict = np.zeros(Q, dtype=np.uint8)

    def step():
        lib.orc_merkle_build_mt(O.vp(leaves), log_n, C, O.vp(nodes), cores)
        pc = np.ascontiguousarray(leaves[idx])
        sib = np.stack([np.stack([nodes[(1 << (log_n - l)) - 1 + ((int(i) >> l) ^ 1)] for l in range(log_n)]) for i in idx])
        lib.orc_merkle_paths_verify_mt(log_n, O.vp(ncols), ctypes.c_size_t(Q), O.vp(idx), O.vp(pc), O.vp(np.ascontiguousarray(sib)),
                                       O.vp(nodes[0].copy()), O.vp(verdict), cores)
        assert verdict.all()

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    perms = perms_per_tree(log_n, C, Q) * args.steps
    v = perms / dt
    sample = "1 tree of 2^%d leaves x %d cols, %d queries per step (the GPU arm's step is %d trees of 2^%d leaves per GPU)" % (
        log_n, C, Q, args.trees, args.log_n)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32 (M31)", "data": "synthetic",
        "config": {"workload": "merkle-decommit-sweep", "log_leaves": args.log_n, "cols": C, "queries": Q, "trees_per_gpu": args.trees},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ----------------------------------------------------------------------------------------------------
def cpu_baseline(args):
    import oracle_py as O
    lib = O.load_oracle()
    lib.orc_merkle_build_mt.restype = ctypes.c_uint64
    cores = os.cpu_count() or 1
    log_n, C = args.ref_log_n, args.cols
    n = 1 << log_n
    leaves = O.synth_m31(0, n * C).reshape(n, C)
    nodes = np.zeros((2 * n - 1, 8), dtype=np.uint32)
    lib.orc_merkle_build_mt(O.vp(leaves), 12, C, O.vp(nodes), cores)     # warm-up
    reps, perms, t0 = 0, 0, time.perf_counter()
    while time.perf_counter() - t0 < args.cpu_seconds:
        perms += lib.orc_merkle_build_mt(O.vp(leaves), log_n, C, O.vp(nodes), cores)
        reps += 1
    dt = time.perf_counter() - t0
    return {"value": perms / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d x commit of one 2^%d-leaf x %d-col tree with the oracle port on %d pthreads (%.1f s)" % (reps, log_n, C, cores, dt)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log-n", type=int, default=20)
    ap.add_argument("--cols", type=int, default=8)
    ap.add_argument("--queries", type=int, default=128)
    ap.add_argument("--trees", type=int, default=4, help="trees per GPU per step")
    ap.add_argument("--ref-log-n", type=int, default=18)
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    pkg = importlib.import_module("recursive-stwo_b200")
    import oracle_py as O     # only for the synthetic-input generator and the cpu_baseline leg

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a B200; there is no CPU fallback"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    pkg.init(local)

    log_n, C, Q, T = args.log_n, args.cols, args.queries, args.trees
    n = 1 << log_n
    # synthetic inputs: tree id = rank*T + t (weak scaling: T trees per GPU)
    h_cols_t = torch.empty((T, C, n), dtype=torch.int32).pin_memory()
    h_cols = h_cols_t.numpy().view(np.uint32)
    h_idx = np.empty((T, Q), dtype=np.uint32)
    for t in range(T):
        tid = rank * T + t
        h_cols[t] = O.synth_m31(tid, C * n).reshape(C, n)
        h_idx[t] = (O.splitmix64(tid ^ 0xABCDEF, Q) & np.uint64(n - 1)).astype(np.uint32)
    d_cols = h_cols_t.to(dev)
    d_idx = torch.from_numpy(h_idx.view(np.int32)).to(dev)
    nodes = torch.empty((T, 2 * n - 1, 8), dtype=torch.int32, device=dev)
    root_id = torch.arange(T, dtype=torch.int32, device=dev).repeat_interleave(Q).contiguous()
    shape = pkg.PathShape.make(log_n, {log_n: C})
    perms_step = perms_per_tree(log_n, C, Q) * T
    l2_note = "inputs larger than L2: %d MB of columns + %d MB of tree nodes per step" % (h_cols.nbytes >> 20, nodes.numel() * 4 >> 20)

    def step():
        pkg.merkle_commit(d_cols, nodes)
        pcols, sib = pkg.merkle_decommit(d_cols, nodes, d_idx)
        roots = nodes[:, 0, :].contiguous()
        return pkg.merkle_path_verify(shape, d_idx.reshape(-1), pcols, sib, roots, root_id)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        v = step()
    assert bool(v.all().item()), "honest paths must verify"
    # ---- timed region: device-resident ----------------------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = pkg.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        v = step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = pkg.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms = float(t_ms.item())
    value = perms_step * world * args.steps / (ms * 1e-3)

    # ---- e2e: host buffers through the host-pointer C ABI -------------------------------------------------
    pcols, sib = pkg.merkle_decommit(d_cols, nodes, d_idx)
    h_pcols = pcols.cpu().numpy().view(np.uint32).copy()
    h_sib = sib.cpu().numpy().view(np.uint32).copy()
    h_rid = root_id.cpu().numpy().view(np.uint32).copy()
    h_flat_idx = h_idx.reshape(-1).copy()

    def e2e_step():
        roots = pkg.merkle_commit_host(h_cols)                                   # H2D columns, D2H roots
        return pkg.merkle_path_verify_host(shape, h_flat_idx, h_pcols, h_sib, roots, h_rid)   # H2D paths, D2H verdicts

    for _ in range(2):
        hv = e2e_step()
    assert hv.all()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(2, args.steps // 2)
    for _ in range(e2e_steps):
        hv = e2e_step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t_e = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_perms = (perms_per_tree(log_n, C, Q) - (0)) * T      # the e2e step verifies the same paths; decommit gather is host-side input
    e2e_value = e2e_perms * world * e2e_steps / float(t_e.item())
    h2d = h_cols.nbytes + h_flat_idx.nbytes + h_pcols.nbytes + h_sib.nbytes + h_rid.nbytes + T * 32
    d2h = T * 32 + T * Q

    # ---- roofline of the dominant kernel (leaf layer of the commit = 2/3 of all permutations) ----------------
    pk = peaks()
    leaf_out = torch.empty((T * n, 8), dtype=torch.int32, device=dev)
    flat_cols = d_cols            # [T, C, n]; hash the leaf layer of tree 0..T-1 with the same kernel code path
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    reps = 5
    pkg.hash_node_batch(None, flat_cols[0], n)
    torch.cuda.synchronize()
    ev[0].record()
    for r in range(reps):
        for t in range(T):
            pkg.hash_node_batch(None, flat_cols[t], n)
    ev[1].record()
    torch.cuda.synchronize()
    leaf_ms = ev[0].elapsed_time(ev[1]) / (reps * T)
    leaf_perms = n * ((C + 7) // 8 + 1)
    leaf_rate = leaf_perms / (leaf_ms * 1e-3)
    achieved_tlops = leaf_rate * LANE_OPS_PER_PERM / 1e12
    alg_bytes = n * (C * 4 + 32)
    roofline = {
        "bound": "int32-issue", "kernel": "k_hash_node_layer (leaf sponge)", "achieved": achieved_tlops, "peak": pk["int_tlops"],
        "unit": "T lane-ops/s", "frac": achieved_tlops / pk["int_tlops"], "peak_src": pk["int_src"],
        "lane_ops_per_perm": LANE_OPS_PER_PERM, "perms_per_s": leaf_rate, "launch_ms": leaf_ms, "traffic": None,
        "hbm": {"bound": "hbm", "achieved": alg_bytes / (leaf_ms * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": alg_bytes / (leaf_ms * 1e-3) / 1e9 / pk["hbm_gbs"], "algorithmic_bytes_per_launch": alg_bytes,
                "peak_src": pk["src"]},
    }

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32 (M31)", "data": "synthetic",
            "config": {"workload": "merkle-decommit-sweep", "log_leaves": log_n, "cols": C, "queries": Q, "trees_per_gpu": T,
                       "perms_per_step_per_gpu": perms_step, "l2": l2_note, "parallelism": "trees sharded by rank, no data-path collective"},
            "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": float(t_e.item()) / e2e_steps * 1e3},
            "roofline": roofline,
        }
        if not args.no_cpu_baseline and world == 1:
            out["cpu_baseline"] = cpu_baseline(args)
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
