"""Seeded synthetic inputs of the BASELINE workloads: splitmix64 streams reduced to M31 words.
Input synthesis only -- no arithmetic of the path lives here."""
import numpy as np

P = (1 << 31) - 1


def splitmix64(seed, n):
    """n u64 outputs of splitmix64 seeded with `seed`"""
    x = np.uint64(seed)
    with np.errstate(over="ignore"):
        idx = np.arange(1, n + 1, dtype=np.uint64)
        z = x + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def synth_m31(seed, n):
    """n canonical M31 words from splitmix64(seed)"""
    return (splitmix64(seed, n) % np.uint64(P)).astype(np.uint32)
