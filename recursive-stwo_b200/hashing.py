"""poseidon31 / merkle primitives over the C ABI.

Mirrors, for batches, the value side of
  primitives/poseidon31/src/lib.rs (Poseidon2HalfVar::permute),
  primitives/merkle/src/lib.rs (Poseidon31MerkleHasherVar),
  components/recursive/data_structures/src/lib.rs:315-354 (SinglePathMerkleProofVar::verify).
Device functions take/return torch CUDA tensors of dtype int32/uint32 (only the data pointer
crosses the ABI); `*_host` functions take numpy uint32 arrays and go through the host-pointer
entry points (copies inside the call).
"""
import ctypes

import numpy as np

from . import _lib

_initialised = False


def init(device=0):
    global _initialised
    _lib.call("stwo_b200_init", device)
    _initialised = True


def _need_init():
    if not _initialised:
        init(_current_device())


def _current_device():
    import torch
    if not torch.cuda.is_available():
        raise _lib.StwoB200Error("stwo_b200_init", _lib.E_NO_DEVICE)
    return torch.cuda.current_device()


def _stream():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _check_dev_u32(t, name):
    import torch
    if not t.is_cuda or not t.is_contiguous() or t.dtype not in (torch.int32, torch.uint32):
        raise TypeError("%s must be a contiguous CUDA int32/uint32 tensor" % name)


def _hptr(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def _check_host_u32(a, name):
    if not isinstance(a, np.ndarray) or a.dtype != np.uint32 or not a.flags.c_contiguous:
        raise TypeError("%s must be a C-contiguous numpy uint32 array" % name)


def launch_count():
    return int(_lib.load().stwo_b200_launch_count())


def path_perms(shape):
    return int(_lib.load().stwo_b200_path_perms(ctypes.byref(shape)))


# ---- K1 -------------------------------------------------------------------------------------------
def poseidon2_permute(states, variant=0):
    """In-place Poseidon2 permutation of an [n, 16] CUDA tensor."""
    _need_init()
    _check_dev_u32(states, "states")
    if states.numel() % 16:
        raise ValueError("states must hold n x 16 words")
    _lib.call("stwo_b200_poseidon2_permute_dev_variant", _dptr(states), states.numel() // 16, variant, _stream())
    return states


def poseidon2_permute_host(states):
    """In-place permutation of an [n, 16] numpy uint32 array through the host-pointer entry point."""
    _need_init()
    _check_host_u32(states, "states")
    if states.size % 16:
        raise ValueError("states must hold n x 16 words")
    _lib.call("stwo_b200_poseidon2_permute", _hptr(states), states.size // 16)
    return states


# ---- hash_node / commit ---------------------------------------------------------------------------
def hash_node_batch(children, cols, n):
    """hash_node for n nodes. children: [n,16] CUDA tensor or None; cols: [n_cols, n] CUDA tensor or None."""
    import torch
    _need_init()
    if children is not None:
        _check_dev_u32(children, "children")
    n_cols = 0
    if cols is not None:
        _check_dev_u32(cols, "cols")
        n_cols = cols.shape[0]
    ref = children if children is not None else cols
    out = torch.empty((n, 8), dtype=ref.dtype, device=ref.device)
    _lib.call("stwo_b200_hash_node_batch_dev", _dptr(children), _dptr(cols), n_cols, n, n, _dptr(out), _stream())
    return out


def merkle_commit(cols, nodes=None):
    """cols: [n_trees, n_cols, 2^log_n] CUDA tensor -> nodes [n_trees, 2^(log_n+1)-1, 8] (root first)."""
    import torch
    _need_init()
    _check_dev_u32(cols, "cols")
    n_trees, n_cols, n = cols.shape
    log_n = n.bit_length() - 1
    if 1 << log_n != n:
        raise ValueError("leaf count must be a power of two")
    if nodes is None:
        nodes = torch.empty((n_trees, 2 * n - 1, 8), dtype=cols.dtype, device=cols.device)
    _lib.call("stwo_b200_merkle_commit_dev", _dptr(cols), n_cols, log_n, n_trees, _dptr(nodes), _stream())
    return nodes


def merkle_commit_host(cols):
    """cols: numpy [n_trees, n_cols, 2^log_n] -> roots numpy [n_trees, 8] (host-pointer entry point)."""
    _need_init()
    _check_host_u32(cols, "cols")
    n_trees, n_cols, n = cols.shape
    log_n = n.bit_length() - 1
    roots = np.empty((n_trees, 8), dtype=np.uint32)
    _lib.call("stwo_b200_merkle_commit", _hptr(cols), n_cols, log_n, n_trees, _hptr(roots))
    return roots


def merkle_decommit(cols, nodes, index):
    """index: [n_trees, n_queries] -> (path_cols [n_trees*n_queries, n_cols], siblings [.., log_n, 8])."""
    import torch
    _need_init()
    n_trees, n_cols, n = cols.shape
    log_n = n.bit_length() - 1
    n_q = index.shape[1]
    path_cols = torch.empty((n_trees * n_q, n_cols), dtype=cols.dtype, device=cols.device)
    sib = torch.empty((n_trees * n_q, log_n, 8), dtype=cols.dtype, device=cols.device)
    _lib.call("stwo_b200_merkle_decommit_dev", _dptr(cols), n_cols, log_n, n_trees, _dptr(nodes), _dptr(index), n_q,
              _dptr(path_cols), _dptr(sib), _stream())
    return path_cols, sib


# ---- K2 -------------------------------------------------------------------------------------------
def merkle_path_verify(shape, index, cols, siblings, roots, root_id=None, want_roots=False):
    """Batched authentication-path verification on device tensors; returns verdict [n] uint8 (and roots)."""
    import torch
    _need_init()
    n_paths = index.numel()
    verdict = torch.empty((n_paths,), dtype=torch.uint8, device=index.device)
    computed = torch.empty((n_paths, 8), dtype=index.dtype, device=index.device) if want_roots else None
    _lib.call("stwo_b200_merkle_path_verify_dev", ctypes.byref(shape), n_paths, _dptr(index), _dptr(cols), _dptr(siblings),
              _dptr(roots), _dptr(root_id), _dptr(verdict), _dptr(computed), _stream())
    return (verdict, computed) if want_roots else verdict


def merkle_path_verify_host(shape, index, cols, siblings, roots, root_id=None, want_roots=False):
    """Same through the host-pointer entry point (numpy in, numpy out; copies inside the call)."""
    _need_init()
    for a, nm in ((index, "index"), (cols, "cols"), (siblings, "siblings"), (roots, "roots")):
        _check_host_u32(a, nm)
    n_paths = index.size
    verdict = np.empty((n_paths,), dtype=np.uint8)
    computed = np.empty((n_paths, 8), dtype=np.uint32) if want_roots else None
    _lib.call("stwo_b200_merkle_path_verify", ctypes.byref(shape), n_paths, _hptr(index), _hptr(cols), _hptr(siblings),
              _hptr(roots), roots.size // 8, _hptr(root_id), _hptr(verdict), _hptr(computed))
    return (verdict, computed) if want_roots else verdict
