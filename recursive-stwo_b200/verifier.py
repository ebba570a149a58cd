"""Batched native verification of PlonkWithPoseidon proofs over the C ABI.

Host-side mirror of the reference's verifier driver for the accelerated path:
  examples/single-proof/src/main.rs:33-83   (hints -> fiat_shamir -> composition -> answer -> folding),
  examples/multi-proofs/src/main.rs:49-139  (the same verifier applied to several proofs).
`verify_proofs` is the call a user of the reference switches to (host blobs in, verdicts out, any mix of shapes);
`VerifyBatch` keeps a same-shape batch resident on the device (the bench's device-resident leg and the parity
tests, which read every intermediate value back).
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import ProofShape, PcsConfig, VerifyDetail, VERIFY_FULL, VERIFY_TIMED, VERIFY_PATH_KERNELS, STAGE_KERNELS, FETCH, STAGES
from .hashing import _need_init, _stream, _dptr

INPUTS_SINGLE = ([1], [[1, 0, 0, 0]])                                            # examples/single-proof/src/main.rs:28-33
INPUTS_RECURSIVE = ([1, 2, 3], [[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0]])      # (1,1), (2,i), (3,j)

# The PcsConfigs the reference's own drivers verify under (pow_bits, log_blowup, log_last, n_queries).  A verifier never takes
# these from the proof: FiatShamirHints::new(&proof, config, ..) gets `config` from its caller
# (components/hints/src/fiat_shamir.rs:69-73).  This allow-list is the default of verify_proofs / VerifyBatch; pass your own.
CONFIG_SINGLE = PcsConfig(20, 5, 2, 16)                   # examples/single-proof/src/main.rs:28-31
CONFIG_STANDARD = PcsConfig(20, 5, 8, 16)                 # examples/multi-proofs/src/main.rs:173-196
CONFIG_FAST_PROVER = PcsConfig(20, 1, 8, 80)
CONFIG_FAST_PROVER2 = PcsConfig(20, 3, 8, 27)
CONFIG_FAST_VERIFIER = PcsConfig(23, 7, 8, 11)
CONFIG_FAST_VERIFIER2 = PcsConfig(20, 8, 8, 10)
CONFIG_FAST_VERIFIER3 = PcsConfig(28, 9, 7, 8)            # also examples/last-layer/src/main.rs:36-39
REFERENCE_CONFIGS = (CONFIG_SINGLE, CONFIG_STANDARD, CONFIG_FAST_PROVER, CONFIG_FAST_PROVER2, CONFIG_FAST_VERIFIER,
                     CONFIG_FAST_VERIFIER2, CONFIG_FAST_VERIFIER3)


def _config_list(config):
    return [config] if isinstance(config, PcsConfig) else list(config)


def _as_aligned(blob):
    """bytes / uint8 array -> (uint8 array whose buffer is 4-byte aligned, length)"""
    a = np.frombuffer(blob, dtype=np.uint8) if not isinstance(blob, np.ndarray) else blob
    n = a.size
    if n % 4 == 0 and a.ctypes.data % 4 == 0 and a.flags["C_CONTIGUOUS"]:
        return a, n                                      # already word-aligned: no copy
    buf = np.zeros((n + 3) // 4, dtype=np.uint32).view(np.uint8)
    buf[:n] = a
    return buf, n


def shape_from_config(config, log_size_plonk, log_size_poseidon):
    """The batch shape a caller's PcsConfig implies for proofs of the given component log sizes (stwo_b200_shape_from_config)."""
    s = ProofShape()
    rc = _lib.load().stwo_b200_shape_from_config(ctypes.byref(config), log_size_plonk, log_size_poseidon, ctypes.byref(s))
    if rc != _lib.OK:
        raise ValueError("no proof shape for %r with log sizes (%d, %d)" % (config, log_size_plonk, log_size_poseidon))
    return s


def shape_for(blob, config=REFERENCE_CONFIGS):
    """Shape to verify `blob`'s batch under: the statement's log sizes from the blob, the PcsConfig from the CALLER's allow-list
    (the one the header names must be in it; ValueError otherwise)."""
    claimed = proof_shape(blob)
    for c in _config_list(config):
        if c.key() == (claimed.pow_bits, claimed.log_blowup, claimed.log_last, claimed.n_queries):
            return shape_from_config(c, claimed.log_size_plonk, claimed.log_size_poseidon)
    raise ValueError("the proof claims a PcsConfig outside the caller's allow-list: pow_bits=%d log_blowup=%d log_last=%d n_queries=%d"
                     % (claimed.pow_bits, claimed.log_blowup, claimed.log_last, claimed.n_queries))


def proof_shape(blob):
    """The shape one proof blob CLAIMS (host-side header walk, no device needed, untrusted); ValueError if it does not parse."""
    buf, n = _as_aligned(blob)
    s = ProofShape()
    rc = _lib.load().stwo_b200_proof_shape_of(buf.ctypes.data_as(ctypes.c_void_p), n, ctypes.byref(s))
    if rc != _lib.OK:
        raise ValueError("not a PlonkWithPoseidon/Poseidon31 proof blob (status %d)" % rc)
    return s


def proof_perms(shape):
    return int(_lib.load().stwo_b200_proof_perms(ctypes.byref(shape)))


def verify_proofs(blobs, inputs=INPUTS_RECURSIVE, full=True, config=REFERENCE_CONFIGS):
    """Verify host blobs (any mix of shapes) -> (verdict uint8[n], stage uint8[n]).  Copies are inside the call.
    config: the PcsConfig (or allow-list of them) to verify under; a blob claiming another one is rejected at stage parse."""
    _need_init()
    n = len(blobs)
    keep = [_as_aligned(b) for b in blobs]
    ptrs = (ctypes.c_void_p * n)(*[k[0].ctypes.data for k in keep])
    lens = (ctypes.c_size_t * n)(*[k[1] for k in keep])
    idx = np.ascontiguousarray(inputs[0], dtype=np.uint32)
    vals = np.ascontiguousarray(inputs[1], dtype=np.uint32)
    verdict = np.full(n, 255, dtype=np.uint8)
    stage = np.full(n, 255, dtype=np.uint8)
    cfgs = _config_list(config)
    c_cfgs = (PcsConfig * len(cfgs))(*[PcsConfig(*c.key()) for c in cfgs])
    _lib.call("stwo_b200_verify_proofs_batch", ptrs, lens, n, c_cfgs, len(cfgs), idx.ctypes.data_as(ctypes.c_void_p), vals.ctypes.data_as(ctypes.c_void_p),
              idx.size, VERIFY_FULL if full else 0, verdict.ctypes.data_as(ctypes.c_void_p), stage.ctypes.data_as(ctypes.c_void_p))
    return verdict, stage


class VerifyBatch:
    """A same-shape batch resident in HBM: blobs, offsets, workspace, verdicts."""

    def __init__(self, blobs, inputs=INPUTS_RECURSIVE, shape=None, device=None, config=REFERENCE_CONFIGS):
        """shape: the shape to verify under (shape_from_config); None: the statement's log sizes of blobs[0] + the PcsConfig of
        the caller's `config` allow-list that its header names (ValueError when it names none of them)."""
        import torch
        _need_init()
        self.n = len(blobs)
        keep = [_as_aligned(b) for b in blobs]
        self.shape = shape if shape is not None else shape_for(blobs[0], config)
        words = [np.frombuffer(k[0][: (k[1] + 3) // 4 * 4].tobytes(), dtype=np.uint32) for k in keep]
        off = np.zeros(self.n + 1, dtype=np.uint64)
        off[1:] = np.cumsum([w.size for w in words])
        dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.h_words = torch.from_numpy(np.concatenate(words).view(np.int32)).pin_memory()
        self.h_off = torch.from_numpy(off.view(np.int64)).pin_memory()
        self.d_words = self.h_words.to(dev)
        self.d_off = self.h_off.to(dev)
        self.d_idx = torch.from_numpy(np.ascontiguousarray(inputs[0], dtype=np.uint32).view(np.int32)).to(dev)
        self.d_vals = torch.from_numpy(np.ascontiguousarray(inputs[1], dtype=np.uint32).view(np.int32)).to(dev)
        self.n_inputs = len(inputs[0])
        self.ws_bytes = int(_lib.load().stwo_b200_verify_workspace_bytes(ctypes.byref(self.shape), self.n))
        if self.ws_bytes == 0:
            raise ValueError("unsupported proof shape")
        self.d_ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=dev)
        self.d_verdict = torch.empty(self.n, dtype=torch.uint8, device=dev)
        self.d_stage = torch.empty(self.n, dtype=torch.uint8, device=dev)

    def upload(self):
        """host (pinned) -> device copy of the blobs, on the current stream"""
        self.d_words.copy_(self.h_words, non_blocking=True)
        self.d_off.copy_(self.h_off, non_blocking=True)

    def run(self, full=True, timed=False, path_kernels=False):
        """path_kernels: produce the permutation record with the thread-per-path kernels (the checker of the default, which takes
        it from the tree rebuilds)"""
        flags = (VERIFY_FULL if full else 0) | (VERIFY_TIMED if timed else 0) | (VERIFY_PATH_KERNELS if path_kernels else 0)
        self.last_full = bool(full)          # full mode records the path permutations the circuit reuses (TRACE_NATIVE_HINTS)
        _lib.call("stwo_b200_verify_proofs_batch_dev", _dptr(self.d_words), _dptr(self.d_off), self.n, ctypes.byref(self.shape),
                  _dptr(self.d_idx), _dptr(self.d_vals), self.n_inputs, flags, _dptr(self.d_ws), self.ws_bytes,
                  _dptr(self.d_verdict), _dptr(self.d_stage), _stream())
        return self.d_verdict, self.d_stage

    def run_from_host(self, full=True):
        """verify with the blobs taken from the pinned host copy: the library uploads each slice on the stream that verifies it,
        so the transfer overlaps the kernels of the other slices (replaces upload() + run())"""
        flags = VERIFY_FULL if full else 0
        self.last_full = bool(full)
        _lib.call("stwo_b200_verify_proofs_batch_pinned_dev", ctypes.c_void_p(self.h_words.data_ptr()), ctypes.c_void_p(self.h_off.data_ptr()),
                  _dptr(self.d_words), _dptr(self.d_off), self.n, ctypes.byref(self.shape), _dptr(self.d_idx), _dptr(self.d_vals),
                  self.n_inputs, flags, _dptr(self.d_ws), self.ws_bytes, _dptr(self.d_verdict), _dptr(self.d_stage), _stream())
        return self.d_verdict, self.d_stage

    def stage_ms(self):
        """device time of each stage kernel of the last run(timed=True) -> {kernel: ms}"""
        ms = (ctypes.c_float * len(STAGE_KERNELS))()
        _lib.call("stwo_b200_verify_stage_ms", ms)
        return dict(zip(STAGE_KERNELS, [float(x) for x in ms]))

    def verdicts_to_host(self):
        return self.d_verdict.cpu().numpy(), self.d_stage.cpu().numpy()

    def fetch(self, p, what):
        """Read one proof's intermediate values back (see STWO_B200_FETCH_* in include/stwo_b200.h)."""
        nq, nf = self.shape.n_queries, 1 + self.shape.n_inner
        if what == "detail":
            out = VerifyDetail()
            _lib.call("stwo_b200_verify_fetch", _dptr(self.d_ws), ctypes.byref(self.shape), self.n, p, FETCH[what], ctypes.byref(out),
                      ctypes.sizeof(out), _stream())
            return out
        shapes = {"domain_points": (3, nq, 2), "answers": (3, nq, 4), "circle_folds": (3, nq, 4), "line_folds": (32, nq, 4),
                  "last_evals": (nq, 4), "path_roots": (4 + nf, nq, 8), "path_cols": (4, nq, 64), "path_siblings": (4, nq, 30, 8),
                  "pair_hints": (nf, nq * 256), "record_trees": (1,),
                  "perm_record": (int(_lib.load().stwo_b200_proof_record_slots(ctypes.byref(self.shape))), 16),
                  "perm_record_inputs": (int(_lib.load().stwo_b200_proof_record_slots(ctypes.byref(self.shape))), 16)}
        out = np.zeros(shapes[what], dtype=np.uint32)
        _lib.call("stwo_b200_verify_fetch", _dptr(self.d_ws), ctypes.byref(self.shape), self.n, p, FETCH[what],
                  out.ctypes.data_as(ctypes.c_void_p), out.nbytes, _stream())
        return out


class VerifyStream:
    """A stream of same-shape batches: two device slots, the next batch uploads (pinned host -> device, its own stream) while
    the current one is verified and traced, so a steady flow of batches runs at the device-resident rate.

        vs = VerifyStream(first_blobs, inputs); circ = VerifierCircuit(vs.shape, inputs)
        vs.feed(blobs_0)                               # upload of batch 0 starts
        for k in ...:
            vs.feed(blobs_k+1)                         # upload of the next batch, beside the work below (None: reuse the host copy)
            batch = vs.take()                          # the uploaded batch (a VerifyBatch), ordered after its upload
            verdict, stage = batch.run(); trace = circ.trace(batch, ...)
            vs.release(batch)                          # its slot may be overwritten once the work queued so far is done
    """

    def __init__(self, blobs, inputs=INPUTS_RECURSIVE, config=REFERENCE_CONFIGS):
        import torch
        self.slots = [VerifyBatch(blobs, inputs=inputs, config=config), VerifyBatch(blobs, inputs=inputs, config=config)]
        self.shape, self.n = self.slots[0].shape, self.slots[0].n
        self.copy_stream = torch.cuda.Stream()
        self.uploaded = [torch.cuda.Event(), torch.cuda.Event()]
        self.free = [torch.cuda.Event(), torch.cuda.Event()]
        for e in self.free:
            e.record()
        self._fed, self._taken = 0, 0

    def feed(self, blobs=None):
        """queue the upload of one batch into the next slot; blobs=None re-sends the slot's pinned host copy"""
        import torch
        i = self._fed % 2
        assert self._fed - self._taken < 2, "both slots hold batches that were not taken yet"
        slot = self.slots[i]
        if blobs is not None:
            # the slot's pinned buffers may still be the source of its previous (asynchronous) upload: that copy must have left the
            # host before they are overwritten (take() only orders the device side)
            if self._fed >= 2:
                self.uploaded[i].synchronize()
            words = np.concatenate([np.frombuffer(_as_aligned(b)[0].tobytes(), dtype=np.uint32) for b in blobs])
            off = np.zeros(len(blobs) + 1, dtype=np.uint64)
            off[1:] = np.cumsum([(len(b) + 3) // 4 for b in blobs])
            if len(blobs) != slot.n or words.size > slot.h_words.numel():
                raise ValueError("a VerifyStream takes batches of the size and (at most) the byte length it was built with")
            slot.h_words[: words.size].copy_(torch.from_numpy(words.view(np.int32)))
            slot.h_off.copy_(torch.from_numpy(off.view(np.int64)))
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.free[i])
            slot.upload()
            self.uploaded[i].record(self.copy_stream)
        self._fed += 1

    def take(self):
        import torch
        assert self._taken < self._fed, "nothing was fed"
        i = self._taken % 2
        torch.cuda.current_stream().wait_event(self.uploaded[i])
        self._taken += 1
        return self.slots[i]

    def release(self, batch):
        self.free[self.slots.index(batch)].record()
