"""Synthetic FRI + Merkle instances (BASELINE configs[4] part i): generated on the device, all different,
verified by the same channel / fold / tree-rebuild kernels a real proof's query phase goes through.  Input synthesis + its verifier
entry; the generator is never inside a timed region."""
import ctypes

import numpy as np

from . import _lib
from ._lib import VerifyDetail, VERIFY_FULL, VERIFY_PATH_KERNELS, FETCH
from .hashing import _need_init, _stream, _dptr


class SynthBatch:
    """n instances of `shape` (seed of instance p = seed0 + p), resident on the device."""

    def __init__(self, shape, n, seed0=0, device=None, chunk=512, distinct=True):
        """distinct=False: n replicas of instance seed0 (the comparison leg: what a replica batch hides)"""
        import torch
        _need_init()
        self.shape, self.n = shape, n
        self.words = int(_lib.load().stwo_b200_synth_blob_words(ctypes.byref(shape)))
        if self.words == 0:
            raise ValueError("no synthetic instances for this shape (proof-shape rules, log_last <= 6, pow_bits <= 10)")
        dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.d_words = torch.empty(n * self.words, dtype=torch.int32, device=dev)
        self.d_off = torch.empty(n + 1, dtype=torch.int64, device=dev)
        self.status = torch.empty(n, dtype=torch.int32, device=dev)
        n_gen = n if distinct else 1
        # the generator keeps whole trees of every instance: built in chunks, the scratch is released afterwards
        for lo in range(0, n_gen, chunk):
            m = min(chunk, n_gen - lo)
            nbytes = int(_lib.load().stwo_b200_synth_scratch_bytes(ctypes.byref(shape), m))
            scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            off = torch.empty(m + 1, dtype=torch.int64, device=dev)
            _lib.call("stwo_b200_synth_generate_dev", ctypes.byref(shape), m, seed0 + lo, _dptr(self.d_words[lo * self.words:]), _dptr(off),
                      _dptr(scratch), nbytes, _dptr(self.status[lo:]), _stream())
            self.d_off[lo: lo + m + 1] = off + lo * self.words
            torch.cuda.synchronize()
            del scratch
        if not distinct:
            self.d_words.view(n, self.words)[1:] = self.d_words.view(n, self.words)[0]
            self.d_off.copy_(torch.arange(n + 1, dtype=torch.int64, device=dev) * self.words)
            self.status[1:] = self.status[0]
        if int((self.status != 0).sum().item()):
            raise RuntimeError("the generator's own checks failed for %d instances" % int((self.status != 0).sum().item()))
        self.ws_bytes = int(_lib.load().stwo_b200_verify_workspace_bytes(ctypes.byref(shape), n))
        self.d_ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=dev)
        self.d_verdict = torch.empty(n, dtype=torch.uint8, device=dev)
        self.d_stage = torch.empty(n, dtype=torch.uint8, device=dev)

    def run(self, full=True, path_kernels=False):
        flags = (VERIFY_FULL if full else 0) | (VERIFY_PATH_KERNELS if path_kernels else 0)
        self.last_full = bool(full)                 # VerifierCircuit(folding=True).trace(self) takes the permutation record then
        _lib.call("stwo_b200_synth_verify_batch_dev", _dptr(self.d_words), _dptr(self.d_off), self.n, ctypes.byref(self.shape), flags,
                  _dptr(self.d_ws), self.ws_bytes, _dptr(self.d_verdict), _dptr(self.d_stage), _stream())
        return self.d_verdict, self.d_stage

    def blob(self, p):
        """host copy of instance p (uint32 words)"""
        return self.d_words[p * self.words: (p + 1) * self.words].cpu().numpy().view(np.uint32).copy()

    def fetch(self, p, what):
        nq, nf = self.shape.n_queries, 1 + self.shape.n_inner
        if what == "detail":
            out = VerifyDetail()
            _lib.call("stwo_b200_verify_fetch", _dptr(self.d_ws), ctypes.byref(self.shape), self.n, p, FETCH[what], ctypes.byref(out),
                      ctypes.sizeof(out), _stream())
            return out
        shapes = {"circle_folds": (3, nq, 4), "line_folds": (32, nq, 4), "last_evals": (nq, 4), "path_roots": (4 + nf, nq, 8), "answers": (3, nq, 4),
                  "record_trees": (1,), "perm_record": (int(_lib.load().stwo_b200_proof_record_slots(ctypes.byref(self.shape))), 16),
                  "perm_record_inputs": (int(_lib.load().stwo_b200_proof_record_slots(ctypes.byref(self.shape))), 16)}
        out = np.zeros(shapes[what], dtype=np.uint32)
        _lib.call("stwo_b200_verify_fetch", _dptr(self.d_ws), ctypes.byref(self.shape), self.n, p, FETCH[what],
                  out.ctypes.data_as(ctypes.c_void_p), out.nbytes, _stream())
        return out
