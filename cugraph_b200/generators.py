"""Synthetic inputs generated on the device with torch (plumbing for tests / bench, not the product).

RMAT follows the sampling rule of the reference generator
(cpp/src/generators/generate_rmat_edgelist.cuh:66-108) and its id scramble (scramble.cuh:44-67),
see oracle/rmat.py for the numpy twin used by the CPU tests."""
from __future__ import annotations


def _bitreverse32(v):
    import torch
    v = ((v >> 1) & 0x55555555) | ((v & 0x55555555) << 1)
    v = ((v >> 2) & 0x33333333) | ((v & 0x33333333) << 2)
    v = ((v >> 4) & 0x0F0F0F0F) | ((v & 0x0F0F0F0F) << 4)
    v = ((v >> 8) & 0x00FF00FF) | ((v & 0x00FF00FF) << 8)
    v = ((v >> 16) | (v << 16)) & 0xFFFFFFFF
    return v


def scramble(v, lgn: int):
    """v: int64 tensor holding values < 2**lgn (lgn <= 31). 32-bit arithmetic emulated in int64."""
    M = 0xFFFFFFFF
    s0, s1 = 282475248, 2617694917
    m0 = (s0 | (0x4519840211493211 & M)) & M
    m1 = (s1 | (0x3050852102C843A5 & M)) & M
    v = (v + s0 + s1) & M
    v = (v * m0) & M
    v = _bitreverse32(v) >> (32 - lgn)
    v = (v * m1) & M
    v = _bitreverse32(v) >> (32 - lgn)
    return v


def rmat_edgelist(scale: int, num_edges: int, a=0.57, b=0.19, c=0.19, seed=0, scramble_ids=True,
                  device="cuda", chunk=1 << 25):
    """(src, dst) int32 CUDA tensors of `num_edges` RMAT edges over 2**scale vertices."""
    import torch
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    a_plus_b = a + b
    a_norm = a / (a + b)
    c_norm = c / (1.0 - (a + b))
    src = torch.empty(num_edges, dtype=torch.int32, device=device)
    dst = torch.empty(num_edges, dtype=torch.int32, device=device)
    done = 0
    while done < num_edges:
        n = min(chunk, num_edges - done)
        s = torch.zeros(n, dtype=torch.int64, device=device)
        d = torch.zeros(n, dtype=torch.int64, device=device)
        for bit in range(scale - 1, -1, -1):
            r0 = torch.rand(n, device=device, generator=gen)
            r1 = torch.rand(n, device=device, generator=gen)
            sb = r0 > a_plus_b
            thr = torch.where(sb, torch.full_like(r1, c_norm), torch.full_like(r1, a_norm))
            db = r1 > thr
            s += sb.to(torch.int64) << bit
            d += db.to(torch.int64) << bit
        if scramble_ids:
            s = scramble(s, scale)
            d = scramble(d, scale)
        src[done:done + n] = s.to(torch.int32)
        dst[done:done + n] = d.to(torch.int32)
        done += n
    return src, dst
