"""The per-lane code of the blocked pull sweep on the CPU: hot_run_groups / hot_run_groups_c1 / hot_run_groups_narrow /
hot_slot_sum / hot_emit (spmv_hot.cuh, spmv_hot_x.cuh) compiled against the host emulation shim and driven lane by lane
over the real piece layout, units in order, groups dealt to the 32 warps as the kernels deal them (emu/emu_debug.cpp:
model_blocked).  What a CUDA kernel adds on top — TMA fills, the unit cursor, barriers — is not modelled.
The fp64 row sums must match numpy on the CSR."""
import ctypes as C

import numpy as np
import pytest

from tests.test_emu_staging_cpu import create_graph, emu, make_edges, primary  # noqa: F401  (emu is a fixture)


def run_model(L, g, P, x, mode):
    L.emu_blocked_sweep.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    L.emu_padded_x_elems.restype = C.c_size_t
    L.emu_padded_x_elems.argtypes = [C.c_int32, C.c_size_t]
    xp = np.zeros(L.emu_padded_x_elems(P["nv"], 4), dtype=np.float32)     # zeros behind n_vertices, as the callers keep it
    xp[:P["nv"]] = x
    acc = np.zeros(max(P["seg"][0], 1), dtype=np.float64)
    rc = L.emu_blocked_sweep(C.c_void_p(L.handle), g, xp.ctypes.data, acc.ctypes.data, mode)
    assert rc == 0, f"emu_blocked_sweep returned {rc}"
    return acc


def expected(P, x):
    n_hi, nnz_hi = P["seg"][0], P["nnz_hi"]
    rows = np.repeat(np.arange(n_hi), np.diff(P["off"][:n_hi + 1]))
    vals = x[P["idx"][:nnz_hi]].astype(np.float64)
    if P["w"] is not None:   # the kernels multiply in the storage type and accumulate in fp64
        vals = (x[P["idx"][:nnz_hi]] * P["w"][:nnz_hi]).astype(np.float64)
    return np.bincount(rows, weights=vals, minlength=n_hi)


CASES = [
    ("default", {}, False, 0),
    ("default weighted", {}, True, 0),
    ("four groups in flight", {}, False, 1),
    ("four groups in flight weighted", {}, True, 1),
    ("cold block", {"CUGRAPH_B200_HOT_BLOCKS": "1"}, False, 1),
    ("cold block weighted", {"CUGRAPH_B200_HOT_BLOCKS": "1"}, True, 0),
    ("narrow classes", {"CUGRAPH_B200_HOT_NARROW": "1"}, False, 1),
    ("narrow classes, small units", {"CUGRAPH_B200_HOT_NARROW": "1", "CUGRAPH_B200_HOT_UNIT_SLOTS": "1024"}, False, 0),
]


@pytest.mark.parametrize("name,env,weighted,mode", CASES, ids=[c[0] for c in CASES])
def test_blocked_sweep_model(emu, monkeypatch, name, env, weighted, mode):  # noqa: F811
    monkeypatch.setenv("CUGRAPH_B200_SWEEP_MIN_EDGES", "0")
    for k in ("CUGRAPH_B200_HOT_BLOCKS", "CUGRAPH_B200_HOT_NARROW", "CUGRAPH_B200_HOT_UNIT_SLOTS"):
        monkeypatch.delenv(k, raising=False)
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    src, dst, w = make_edges(130_000, 800_000, seed=31 + len(name), weighted=weighted)
    g = create_graph(emu, src, dst, w)
    P = primary(emu, g)
    assert P["seg"][0] > 500
    x = np.random.default_rng(2).random(P["nv"]).astype(np.float32)
    acc = run_model(emu, g, P, x, mode)
    np.testing.assert_allclose(acc[:P["seg"][0]], expected(P, x), rtol=1e-12, atol=0)
    emu.cugraph_graph_free(g)


@pytest.mark.parametrize("weighted", [False, True])
def test_low_ell_hot_model(emu, monkeypatch, weighted):  # noqa: F811
    """k_spmv_low_ell_hot's loop structure and its shared-memory / global gather split"""
    monkeypatch.setenv("CUGRAPH_B200_LOW_ELL", "2")
    src, dst, w = make_edges(140_000, 500_000, seed=77 + weighted, weighted=weighted)
    g = create_graph(emu, src, dst, w)
    P = primary(emu, g)
    nv, n_hi = P["nv"], P["seg"][0]
    assert nv > 49088                      # sources on both sides of the shared-memory slice
    emu.emu_padded_x_elems.restype = C.c_size_t
    emu.emu_padded_x_elems.argtypes = [C.c_int32, C.c_size_t]
    emu.emu_low_ell_hot_sweep.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_int]
    xp = np.zeros(emu.emu_padded_x_elems(nv, 4), dtype=np.float32)
    xp[:nv] = np.random.default_rng(4).random(nv).astype(np.float32)
    y = np.full(nv, -3.0, dtype=np.float32)
    rc = emu.emu_low_ell_hot_sweep(C.c_void_p(emu.handle), g, xp.ctypes.data, y.ctypes.data, 0.85, 0.25, 7)
    assert rc == 0
    deg = np.diff(P["off"])
    rows = np.repeat(np.arange(P["n_rows"]), deg)
    vals = xp[P["idx"]]
    if weighted:
        vals = vals * P["w"]
    exp = np.bincount(rows, weights=vals.astype(np.float64), minlength=P["n_rows"]) * 0.85 + 0.25
    assert (y[:n_hi] == -3.0).all()
    np.testing.assert_allclose(y[n_hi:], exp[n_hi:].astype(np.float32), rtol=2e-6, atol=0)
    emu.cugraph_graph_free(g)
