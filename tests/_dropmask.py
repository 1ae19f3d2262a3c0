"""Vectorised numpy restatement of the library's counter-based dropout mask (drakegpt_b200/csrc/common.cuh:
one SplitMix64 per group of 32 consecutive elements, one 32-bit multiply-add per element, keep iff word >= p * 2^32).
Test infrastructure only; tests/test_cpu_host.py pins it against the C-ABI's dgpt_dropout_keep_host."""
import numpy as np
import torch

_M32 = np.uint64(0xFFFFFFFF)


def _mix64(z):
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def _drop_mul(e):
    x = ((e + 1) * 0x9E3779B1) & 0xFFFFFFFF
    x ^= x >> 15
    x = (x * 0x85EBCA77) & 0xFFFFFFFF
    x ^= x >> 13
    return x | 1


def _drop_add(e):
    x = ((e + 33) * 0xC2B2AE3D) & 0xFFFFFFFF
    x ^= x >> 16
    x = (x * 0x27D4EB2F) & 0xFFFFFFFF
    x ^= x >> 15
    return x


_MUL = np.array([_drop_mul(e) for e in range(32)], dtype=np.uint64)
_ADD = np.array([_drop_add(e) for e in range(32)], dtype=np.uint64)


def threshold(p):
    return min(max(int(p * 4294967296.0 + 0.5), 0), 4294967295)


def words(seed, site, n, start=0):
    """The 32-bit mask words of elements [start, start + n) (row-major element index)."""
    with np.errstate(over="ignore"):
        idx = np.arange(start, start + n, dtype=np.uint64)
        grp, e = idx >> np.uint64(5), (idx & np.uint64(31)).astype(np.int64)
        h = _mix64(np.uint64(seed % (1 << 64)) + grp * np.uint64(0x9E3779B97F4A7C15)
                   + np.uint64(site + 1) * np.uint64(0xD1B54A32D192ED03))
        s = np.where(e & 1, h >> np.uint64(32), h & _M32)
        return (s * _MUL[e] + _ADD[e]) & _M32


def keep_mask(shape, seed, site, p):
    """float32 torch tensor of `shape`: 1 where the element is kept, 0 where dropped."""
    n = int(np.prod(shape))
    return torch.from_numpy((words(seed, site, n) >= np.uint64(threshold(p))).astype(np.float32)).view(*shape)
