"""Device RMAT generator (csrc/generators.cu behind cugraph_b200_generate_rmat_edgelist) vs its numpy restatement
(oracle/rmat.py:rmat_edgelist_counter), bit for bit, through the emulated library; and the sampling rule itself against the
reference-shaped generator of oracle/rmat.py on distribution level (quadrant frequencies of the top bit)."""
import ctypes as C

import numpy as np
import pytest

from oracle.rmat import rmat_edgelist_counter, uniform_counter
from tests.test_emu_staging_cpu import INT32, emu  # noqa: F401


def _generate(L, scale, n, a, b, c, seed, clip, scramble):
    L.cugraph_b200_generate_rmat_edgelist.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_double, C.c_double, C.c_double,
                                                      C.c_uint64, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    src = np.full(n, -7, dtype=np.int32)
    dst = np.full(n, -7, dtype=np.int32)
    vs = C.c_void_p(L.cugraph_type_erased_device_array_view_create(src.ctypes.data, n, INT32))
    vd = C.c_void_p(L.cugraph_type_erased_device_array_view_create(dst.ctypes.data, n, INT32))
    err = C.c_void_p()
    code = L.cugraph_b200_generate_rmat_edgelist(C.c_void_p(L.handle), scale, n, a, b, c, seed, int(clip), int(scramble), vs, vd,
                                                 C.byref(err))
    return code, src, dst, err


@pytest.mark.parametrize("scale,n,seed,clip,scramble", [(10, 20000, 0, False, True), (17, 50000, 12345, False, True),
                                                         (12, 30000, 7, True, False), (31, 4000, 99, True, True),
                                                         (1, 100, 3, False, False)])
def test_device_rmat_matches_numpy_twin(emu, scale, n, seed, clip, scramble):  # noqa: F811
    code, src, dst, err = _generate(emu, scale, n, 0.57, 0.19, 0.19, seed, clip, scramble)
    assert code == 0, emu.cugraph_error_message(err)
    rs, rd = rmat_edgelist_counter(scale, n, 0.57, 0.19, 0.19, seed, clip, scramble)
    assert np.array_equal(src, rs) and np.array_equal(dst, rd)
    assert src.min() >= 0 and dst.min() >= 0
    if scale < 31:
        assert src.max() < (1 << scale) and dst.max() < (1 << scale)
    if clip and not scramble:
        assert (src >= dst).all()          # clip-and-flip keeps every edge on or below the diagonal


def test_device_rmat_quadrants_and_errors(emu):  # noqa: F811
    n = 200000
    code, src, dst, err = _generate(emu, 8, n, 0.5, 0.2, 0.2, 5, False, False)
    assert code == 0
    top_s, top_d = src >> 7, dst >> 7
    freq = np.array([((top_s == i) & (top_d == j)).mean() for i in (0, 1) for j in (0, 1)])   # a, b, c, d
    assert np.allclose(freq, [0.5, 0.2, 0.2, 0.1], atol=0.01), freq
    code, *_ = _generate(emu, 8, 10, 0.6, 0.3, 0.3, 5, False, False)     # a + b + c > 1
    assert code != 0
    code, *_ = _generate(emu, 32, 10, 0.57, 0.19, 0.19, 5, False, False)  # ids would not fit 32 bits
    assert code != 0


@pytest.mark.parametrize("dtype,tid,lo,hi", [(np.float32, 8, 0.0, 1.0), (np.float64, 9, -2.5, 7.25), (np.int32, 2, 2, 6)])
def test_device_uniform_matches_numpy_twin(emu, dtype, tid, lo, hi):  # noqa: F811
    L = emu
    L.cugraph_b200_generate_uniform.argtypes = [C.c_void_p, C.c_uint64, C.c_double, C.c_double, C.c_void_p, C.c_void_p]
    n = 10000
    out = np.zeros(n, dtype=dtype)
    v = C.c_void_p(L.cugraph_type_erased_device_array_view_create(out.ctypes.data, n, tid))   # 8 FLOAT32, 9 FLOAT64, 2 INT32
    err = C.c_void_p()
    code = L.cugraph_b200_generate_uniform(C.c_void_p(L.handle), 4242, lo, hi, v, C.byref(err))
    assert code == 0, L.cugraph_error_message(err)
    ref = uniform_counter(n, 4242, lo, hi, dtype)
    assert np.array_equal(out, ref)
    assert out.min() >= lo and out.max() < hi
    code = L.cugraph_b200_generate_uniform(C.c_void_p(L.handle), 1, 3.0, 3.0, v, C.byref(err))   # empty range
    assert code != 0
