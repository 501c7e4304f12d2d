"""Deterministic synthetic CSDR input streams for the BASELINE configs.

A CSDR is an int32 array with one active-cell index per column, column (x, y) at index y + x*size_y
(reference: Helpers.h:240-245). All generators are pure functions of (t, shape, seed), so the product, the CPU
checkers and a multi-GPU run can regenerate the same inputs independently. Noise comes from a counter-based
integer hash (no shared RNG state).
"""
import math

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _hash_u32(seed: int, t: int, n: int) -> np.ndarray:
    """splitmix64-style hash of (seed, t, column index) -> uint32 per column."""
    with np.errstate(over="ignore"):
        z = (np.arange(n, dtype=np.uint64) + np.uint64((seed * 0x9E3779B97F4A7C15 + t * 0xD1B54A32D192ED03) & 0xFFFFFFFFFFFFFFFF))
        z = (z + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(32)).astype(np.uint32)


def sine(t: int, size=(1, 1, 32)) -> np.ndarray:
    """C1: c_t = int((sin(t*0.02*2pi)*0.5+0.5)*(Z-1) + 0.5), broadcast to every column."""
    z = size[2]
    v = int((math.sin(t * 0.02 * 2.0 * math.pi) * 0.5 + 0.5) * (z - 1) + 0.5)
    return np.full(size[0] * size[1], v, dtype=np.int32)


def coherent(t: int, size, shift: int = 0) -> np.ndarray:
    """c_t(x,y) = ((3(x+t') + 5(y+2t')) // 4) mod Z with t' = t + shift: a drifting diagonal pattern."""
    sx, sy, z = size
    tt = t + shift
    x = np.arange(sx, dtype=np.int64)[:, None]
    y = np.arange(sy, dtype=np.int64)[None, :]
    return ((((3 * (x + tt) + 5 * (y + 2 * tt)) // 4) % z).astype(np.int32)).reshape(-1)


def noisy(t: int, size, seed: int = 7, p: float = 0.02, shift: int = 0) -> np.ndarray:
    """coherent() with each column independently replaced w.p. p by a uniform cell index."""
    sx, sy, z = size
    base = coherent(t, size, shift)
    h = _hash_u32(seed, t, sx * sy)
    hit = (h >> np.uint32(8)).astype(np.float64) * (1.0 / (1 << 24)) < p
    repl = (_hash_u32(seed ^ 0x5BD1E995, t, sx * sy) % np.uint32(z)).astype(np.int32)
    return np.where(hit, repl, base).astype(np.int32)


def iid(t: int, size, seed: int = 7) -> np.ndarray:
    """every column uniform in {0..Z-1}: worst case for SparseCoder weight writes."""
    sx, sy, z = size
    return (_hash_u32(seed, t, sx * sy) % np.uint32(z)).astype(np.int32)


def make(kind: str, t: int, size, seed: int = 7, shift: int = 0) -> np.ndarray:
    if kind == "sine":
        return sine(t, size)
    if kind == "coherent":
        return coherent(t, size, shift)
    if kind == "noisy":
        return noisy(t, size, seed, shift=shift)
    if kind == "iid":
        return iid(t, size, seed)
    raise ValueError(kind)
