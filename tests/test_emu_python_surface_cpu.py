"""The Python surface on the CPU: cugraph_b200.pylibcugraph wrappers, bench.py's single-GPU arm and scripts/bench_side.py
driven through the emulation build of the library (tests/emu_py.py).  Catches Python-level mistakes in the wrappers and
the measurement scripts before they reach the GPU box; says nothing about timing or stream ordering."""
import argparse
import importlib.util
import io
import json
import os
import sys

import numpy as np
import pytest

import oracle
from oracle.rmat import rmat_edgelist as rmat_np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def surface():
    torch = pytest.importorskip("torch")  # noqa: F841
    from tests.emu_py import emulated_python_surface
    try:
        cm = emulated_python_surface()
        L = cm.__enter__()
    except Exception as e:  # no host compiler
        pytest.skip(f"emulation build unavailable: {e}")
    yield L
    cm.__exit__(None, None, None)


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_wrappers_match_oracle(surface):
    import torch
    from cugraph_b200 import pylibcugraph as plc
    scale = 9
    V = 1 << scale
    s, d = rmat_np(scale, 16 << scale, seed=3)
    h = plc.ResourceHandle()
    verts_all = torch.arange(V, dtype=torch.int32)
    g = plc.SGGraph(h, plc.GraphProperties(is_multigraph=True), torch.as_tensor(s), torch.as_tensor(d),
                    store_transposed=True, renumber=True, vertices_array=verts_all)
    v, p, conv = plc.pagerank(h, g, None, None, None, None, 0.85, 0.0, 20, False, fail_on_nonconvergence=False)
    ref, _, _ = oracle.pagerank(s, d, V, None, alpha=0.85, epsilon=0.0, max_iterations=20)
    got = np.zeros(V)
    got[v.numpy()] = p.numpy()
    np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-12)
    assert conv is False or conv == 0
    s2, d2 = np.concatenate([s, d]), np.concatenate([d, s])
    w = np.random.default_rng(0).random(s.shape[0]).astype(np.float32)
    w2 = np.concatenate([w, w])
    g2 = plc.SGGraph(h, plc.GraphProperties(is_symmetric=True, is_multigraph=True), torch.as_tensor(s2), torch.as_tensor(d2),
                     weight_array=torch.as_tensor(w2), renumber=True, vertices_array=verts_all)
    src = int(s[0])
    dist, pred, bv = plc.bfs(h, g2, torch.tensor([src], dtype=torch.int32), True, 0, True, False)
    rd, _ = oracle.bfs(s2, d2, V, [src])
    gd = np.zeros(V, dtype=np.int32)
    gd[bv.numpy()] = dist.numpy()
    assert np.array_equal(gd, rd)
    sv, sd, sp = plc.sssp(h, g2, src, float("inf"), True, False)
    rs, _ = oracle.sssp(s2, d2, w2, V, src)
    gs = np.zeros(V)
    gs[sv.numpy()] = sd.numpy()
    assert np.array_equal(gs, rs)


def test_bench_single_gpu_arm(surface, monkeypatch, capsys):
    """bench.run_single end to end at a toy scale: one JSON line with every key of the contract; the side processes
    (which need a real GPU) fail here and must not take the main line with them."""
    monkeypatch.setenv("CUGRAPH_B200_SWEEP_MIN_EDGES", "0")
    monkeypatch.setenv("CUGRAPH_B200_BENCH_SIDE_BUDGET_S", "0")   # traversal process only; variants are skipped
    bench = _load(os.path.join(ROOT, "bench.py"), "bench_under_test")
    monkeypatch.setattr(bench, "_run_side", lambda argv, timeout_s: {"error": "no GPU in this test"})
    args = argparse.Namespace(gpus=1, steps=2, warmup=1, impl="b200", scale=10, cpu_sample_scale=10)
    bench.run_single(args)
    line = [ln for ln in capsys.readouterr().out.splitlines() if ln.startswith("{")][-1]
    out = json.loads(line)
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline",
                "timing", "side"):
        assert key in out, key
    assert out["value"] > 0 and out["gpu_launches"] > 0 and out["steps"] == 2
    assert out["e2e"]["value"] is not None and out["e2e"]["value"] > 0, out["e2e"]
    assert out["e2e"]["h2d_bytes_per_step"] == 2 * 4 * (16 << 10)
    assert out["roofline"]["bound"] == "hbm" and out["roofline"]["achieved"] > 0 and 0 < out["roofline"]["frac"]
    assert out["cpu_baseline"]["kind"] == "port" and out["cpu_baseline"]["value"] > 0
    assert out["cpu_baseline"]["networkx"]["value"] > 0 and out["cpu_baseline"]["networkx"]["cores"] == 1
    assert out["side"]["traversal"] == {"error": "no GPU in this test"}
    assert all("skipped" in v for v in out["side"]["variants"])


def test_bench_side_subprocess_failure_is_contained(monkeypatch):
    """_run_side with a script that cannot succeed here (no GPU): an error record, not an exception"""
    bench = _load(os.path.join(ROOT, "bench.py"), "bench_under_test2")
    res = bench._run_side(["no-such-mode"], 120)
    assert "error" in res


def _side_variants():
    return _load(os.path.join(ROOT, "bench.py"), "bench_for_variant_list").SIDE_VARIANTS


@pytest.mark.parametrize("cfg", _side_variants())
def test_bench_side_variant(surface, monkeypatch, capsys, cfg):
    """every switch set bench.py's side run will time on the GPU: parity of the configured sweep against the plain one"""
    monkeypatch.setenv("CUGRAPH_B200_SWEEP_MIN_EDGES", "0")
    side = _load(os.path.join(ROOT, "scripts", "bench_side.py"), "bench_side_under_test")
    for kv in cfg.split(","):
        if "=" in kv:
            monkeypatch.setenv(*kv.split("="))
    side.variant(10, cfg)
    out = json.loads(capsys.readouterr().out.splitlines()[-1])
    assert out["config"] == cfg and out["parity_ok"], out
    assert out["sweep_ms"] > 0 and abs(out["pagerank_mass"] - 1.0) < 1e-4


def test_bench_side_traversal(surface, capsys):
    side = _load(os.path.join(ROOT, "scripts", "bench_side.py"), "bench_side_under_test3")
    side.traversal(9, 2, 1)
    out = json.loads(capsys.readouterr().out.splitlines()[-1])
    assert len(out["schedule_ab"]) == 4 and all(r.get("harmonic_mean_mteps", 0) > 0 for r in out["schedule_ab"]), out["schedule_ab"]
    for name in ("bfs", "sssp"):
        assert out[name]["harmonic_mean_mteps"] > 0
        assert all(out[name]["check"][k] for k in ("tree_property", "source_distance_zero",
                                                   "every_reached_vertex_but_the_source_has_a_predecessor")), out[name]


def test_graft_entry_smoke(surface, capsys):
    """__graft_entry__.smoke() itself (scale-12 PageRank + BFS + SSSP against the oracle), kernels emulated"""
    entry = _load(os.path.join(ROOT, "__graft_entry__.py"), "graft_entry_under_test")
    entry.smoke()
    assert "smoke ok" in capsys.readouterr().out
