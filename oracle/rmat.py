"""RMAT edge-list generator — TEST/BENCH INFRASTRUCTURE (oracle side), numpy only.

Restates the sampling rule of the reference generator
(cpp/src/generators/generate_rmat_edgelist.cuh:66-108): for every edge and for every
bit from scale-1 down to 0 draw two uniforms r0, r1;
    src_bit = r0 > (a+b);   dst_bit = r1 > (src_bit ? c/(1-(a+b)) : a/(a+b))
and optionally apply the Graph500 id scramble (cpp/src/generators/scramble.cuh:44-67).
The reference draws its uniforms from raft's device RNG, which is not vendored, so the
*stream* differs; the distribution is the same (parity is never checked on RMAT ids, only on
algorithm outputs computed from the same generated edge list).
"""
from __future__ import annotations

import numpy as np


def _bitreverse32(v: np.ndarray) -> np.ndarray:
    v = v.astype(np.uint32)
    v = ((v >> np.uint32(1)) & np.uint32(0x55555555)) | ((v & np.uint32(0x55555555)) << np.uint32(1))
    v = ((v >> np.uint32(2)) & np.uint32(0x33333333)) | ((v & np.uint32(0x33333333)) << np.uint32(2))
    v = ((v >> np.uint32(4)) & np.uint32(0x0F0F0F0F)) | ((v & np.uint32(0x0F0F0F0F)) << np.uint32(4))
    v = ((v >> np.uint32(8)) & np.uint32(0x00FF00FF)) | ((v & np.uint32(0x00FF00FF)) << np.uint32(8))
    v = (v >> np.uint32(16)) | (v << np.uint32(16))
    return v


def scramble32(value: np.ndarray, lgn: int) -> np.ndarray:
    """32-bit variant of the reference's `detail::scramble` (scramble.cuh:44-67)."""
    s0 = np.uint32(282475248)
    s1 = np.uint32(2617694917)
    with np.errstate(over="ignore"):
        v = value.astype(np.uint32)
        v = v + s0 + s1
        v = v * (s0 | np.uint32(0x4519840211493211 & 0xFFFFFFFF))
        v = _bitreverse32(v) >> np.uint32(32 - lgn)
        v = v * (s1 | np.uint32(0x3050852102C843A5 & 0xFFFFFFFF))
        v = _bitreverse32(v) >> np.uint32(32 - lgn)
    return v.astype(np.int32)


def rmat_edgelist(scale: int, num_edges: int, a: float = 0.57, b: float = 0.19, c: float = 0.19,
                  seed: int = 0, clip_and_flip: bool = False, scramble: bool = True,
                  chunk: int = 1 << 22):
    """Return (src, dst) int32 arrays of `num_edges` RMAT edges over 2**scale vertices."""
    assert scale < 31
    rng = np.random.Generator(np.random.Philox(seed))
    a_plus_b = np.float32(a + b)
    a_norm = np.float32(a / (a + b) if (a + b) > 0 else 0.0)
    c_norm = np.float32(c / (1.0 - (a + b)) if (1.0 - (a + b)) > 0 else 0.0)
    src = np.empty(num_edges, dtype=np.int32)
    dst = np.empty(num_edges, dtype=np.int32)
    done = 0
    while done < num_edges:
        n = min(chunk, num_edges - done)
        s = np.zeros(n, dtype=np.int32)
        d = np.zeros(n, dtype=np.int32)
        for bit in range(scale - 1, -1, -1):
            r0 = rng.random(n, dtype=np.float32)
            r1 = rng.random(n, dtype=np.float32)
            sb = r0 > a_plus_b
            db = r1 > np.where(sb, c_norm, a_norm)
            if clip_and_flip:
                flip = (s == d) & (~sb) & db
                sb = sb ^ flip
                db = db ^ flip
            s += sb.astype(np.int32) << bit
            d += db.astype(np.int32) << bit
        if scramble:
            s = scramble32(s, scale)
            d = scramble32(d, scale)
        src[done:done + n] = s
        dst[done:done + n] = d
        done += n
    return src, dst


def _mix64(z: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def rmat_edgelist_counter(scale: int, num_edges: int, a: float = 0.57, b: float = 0.19, c: float = 0.19, seed: int = 0,
                          clip_and_flip: bool = False, scramble: bool = True):
    """numpy twin of the DEVICE generator (cugraph_b200/csrc/generators.cu: cugraph_b200_generate_rmat_edgelist) — the
    same sampling rule as rmat_edgelist above over the device's counter-based uniform stream: for edge e and bit k,
    r = mix64(seed ^ (64 e + k)), r0 = (r >> 40) / 2^24, r1 = ((r >> 8) & 0xffffff) / 2^24.  Bit-exact with the device."""
    assert 1 <= scale <= 31
    ab = np.float32(a + b)
    a_norm = np.float32(a / (a + b) if (a + b) > 0 else 0.0)
    c_norm = np.float32(c / (1.0 - (a + b)) if (1.0 - (a + b)) > 0 else 0.0)
    e = np.arange(num_edges, dtype=np.uint64)
    s = np.zeros(num_edges, dtype=np.uint32)
    d = np.zeros(num_edges, dtype=np.uint32)
    inv = np.float32(1.0 / 16777216.0)
    for bit in range(scale - 1, -1, -1):
        with np.errstate(over="ignore"):
            r = _mix64(np.uint64(seed) ^ (e * np.uint64(64) + np.uint64(bit)))
        r0 = (r >> np.uint64(40)).astype(np.float32) * inv
        r1 = ((r >> np.uint64(8)) & np.uint64(0xFFFFFF)).astype(np.float32) * inv
        sb = r0 > ab
        db = r1 > np.where(sb, c_norm, a_norm)
        if clip_and_flip:
            flip = (s == d) & (~sb) & db
            sb = sb | flip
            db = db & ~flip
        s |= sb.astype(np.uint32) << np.uint32(bit)
        d |= db.astype(np.uint32) << np.uint32(bit)
    if scramble:
        s = scramble32(s, scale).astype(np.uint32)
        d = scramble32(d, scale).astype(np.uint32)
    return s.astype(np.int32), d.astype(np.int32)


def uniform_counter(n: int, seed: int, lo: float, hi: float, dtype=np.float32):
    """numpy twin of cugraph_b200_generate_uniform (csrc/generators.cu): value i = lo + u_i (hi - lo) with u_i the top 24
    (float32) / 53 (float64) bits of mix64(seed ^ i) as a fraction; int32: lo + mix64(seed ^ i) % (hi - lo)."""
    with np.errstate(over="ignore"):
        r = _mix64(np.uint64(seed) ^ np.arange(n, dtype=np.uint64))
    if np.dtype(dtype) == np.int32:
        return (np.int64(lo) + (r % np.uint64(int(hi) - int(lo))).astype(np.int64)).astype(np.int32)
    if np.dtype(dtype) == np.float32:
        u = (r >> np.uint64(40)).astype(np.float64) * (1.0 / 16777216.0)
    else:
        u = (r >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    return (lo + u * (hi - lo)).astype(dtype)
