"""Stage-level entry points of the boundary: replace ONE hint stage of the reference for a batch of host blobs.

Host-side mirror of
  FiatShamirHints::new                 components/hints/src/fiat_shamir.rs:69-307        -> channel_replay
  AnswerHints::compute                 components/hints/src/answer.rs:40-48              -> fri_answers
  FirstLayerHints / InnerLayersHints   components/hints/src/folding.rs:326-363,481-595   -> fri_folds
  hash_column_get_capacity             components/hints/src/folding.rs:77                -> hash_column_capacity
Every call takes the PcsConfig from the CALLER (like the reference's `config` argument); blobs in, numpy arrays out.
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import PcsConfig, VerifyDetail
from .hashing import _need_init
from .verifier import INPUTS_RECURSIVE, _as_aligned


def _args(blobs, config, inputs):
    _need_init()
    n = len(blobs)
    keep = [_as_aligned(b) for b in blobs]
    ptrs = (ctypes.c_void_p * n)(*[k[0].ctypes.data for k in keep])
    lens = (ctypes.c_size_t * n)(*[k[1] for k in keep])
    idx = np.ascontiguousarray(inputs[0], dtype=np.uint32)
    vals = np.ascontiguousarray(inputs[1], dtype=np.uint32)
    cfg = PcsConfig(*config.key())
    verdict, stage = np.full(n, 255, dtype=np.uint8), np.full(n, 255, dtype=np.uint8)
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    return keep, (ptrs, lens, n, ctypes.byref(cfg), vp(idx), vp(vals), idx.size), (cfg, idx, vals), verdict, stage, vp


def channel_replay(blobs, config, inputs=INPUTS_RECURSIVE):
    """-> (details: ctypes array of VerifyDetail (every draw, the OODS values, PoW), verdict, stage) up to the OODS check"""
    keep, head, hold, verdict, stage, vp = _args(blobs, config, inputs)
    out = (VerifyDetail * len(blobs))()
    _lib.call("stwo_b200_channel_replay_batch", *head, out, vp(verdict), vp(stage))
    return out, verdict, stage


def fri_answers(blobs, config, inputs=INPUTS_RECURSIVE):
    """-> (answers [n, 3, n_queries, 4], domain_points [n, 3, n_queries, 2], verdict, stage)"""
    keep, head, hold, verdict, stage, vp = _args(blobs, config, inputs)
    n, nq = len(blobs), config.n_queries
    answers = np.zeros((n, 3, nq, 4), dtype=np.uint32)
    points = np.zeros((n, 3, nq, 2), dtype=np.uint32)
    _lib.call("stwo_b200_fri_answers_batch", *head, vp(answers), vp(points), vp(verdict), vp(stage))
    return answers, points, verdict, stage


def fri_folds(blobs, config, inputs=INPUTS_RECURSIVE):
    """-> (circle_folds [n, 3, nq, 4], line_folds [n, 32, nq, 4], last_evals [n, nq, 4], verdict, stage)"""
    keep, head, hold, verdict, stage, vp = _args(blobs, config, inputs)
    n, nq = len(blobs), config.n_queries
    circle = np.zeros((n, 3, nq, 4), dtype=np.uint32)
    line = np.zeros((n, 32, nq, 4), dtype=np.uint32)
    last = np.zeros((n, nq, 4), dtype=np.uint32)
    _lib.call("stwo_b200_fri_fold_batch", *head, vp(circle), vp(line), vp(last), vp(verdict), vp(stage))
    return circle, line, last, verdict, stage


def hash_column_capacity(cols):
    """cols: [n, n_cols] uint32 (host) -> [n, 8] capacities of the column sponge (primitives/merkle/src/lib.rs:141-181)"""
    _need_init()
    cols = np.ascontiguousarray(cols, dtype=np.uint32)
    n, n_cols = cols.shape
    out = np.zeros((n, 8), dtype=np.uint32)
    _lib.call("stwo_b200_hash_column_capacity_batch", cols.ctypes.data_as(ctypes.c_void_p), n_cols, n, out.ctypes.data_as(ctypes.c_void_p))
    return out
