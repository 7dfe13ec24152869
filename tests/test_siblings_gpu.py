"""Katz centrality, HITS and weakly connected components on the GPU (pylibcugraph-compatible wrappers over the C ABI) against the
numpy restatements of the reference tests' CPU references (oracle.katz / .hits / .wcc)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _graph(scale=14, symmetric=False, store_transposed=True, weighted=False):
    import torch
    from cugraph_b200 import pylibcugraph as plc
    from oracle.rmat import rmat_edgelist
    s, d = rmat_edgelist(scale, 8 << scale, seed=3)
    if symmetric:
        s, d = np.concatenate([s, d]), np.concatenate([d, s])
    w = (np.random.default_rng(1).random(s.size).astype(np.float32) + 0.5) if weighted else None
    h = plc.ResourceHandle()
    g = plc.SGGraph(h, plc.GraphProperties(is_symmetric=symmetric, is_multigraph=True), torch.as_tensor(s).cuda(),
                    torch.as_tensor(d).cuda(), weight_array=None if w is None else torch.as_tensor(w).cuda(),
                    store_transposed=store_transposed, renumber=True)
    ids, inv = np.unique(np.concatenate([s, d]), return_inverse=True)
    return plc, h, g, ids, inv[:s.size], inv[s.size:], w


@pytest.mark.parametrize("weighted", [False, True])
def test_katz_gpu(weighted):
    import oracle
    plc, h, g, ids, s, d, w = _graph(weighted=weighted)
    alpha = 0.5 / (np.bincount(d).max() * (float(w.max()) if weighted else 1.0))
    verts, vals = plc.katz_centrality(h, g, None, alpha, 1.0, 1e-5, 500, False)
    ref, _ = oracle.katz(s, d, ids.size, w, alpha=alpha, beta=1.0, epsilon=1e-5, dtype=np.float32)
    got = np.zeros(ids.size)
    got[np.searchsorted(ids, verts.cpu().numpy())] = vals.cpu().numpy()
    np.testing.assert_allclose(got, ref, rtol=2e-5)


@pytest.mark.parametrize("store_transposed", [True, False])
def test_hits_gpu(store_transposed):
    import oracle
    plc, h, g, ids, s, d, _ = _graph(store_transposed=store_transposed)
    verts, hubs, auth = plc.hits(h, g, 1e-7, 500, None, None, True, False)
    rh, ra, _, _ = oracle.hits(s, d, ids.size, epsilon=1e-7)
    v = np.searchsorted(ids, verts.cpu().numpy())
    gh, ga = np.zeros(ids.size), np.zeros(ids.size)
    gh[v], ga[v] = hubs.cpu().numpy(), auth.cpu().numpy()
    np.testing.assert_allclose(gh, rh, rtol=2e-3, atol=1e-9)
    np.testing.assert_allclose(ga, ra, rtol=2e-3, atol=1e-9)


def test_wcc_gpu():
    import oracle
    plc, h, g, ids, s, d, _ = _graph(scale=16, symmetric=True, store_transposed=False)
    verts, labels = plc.weakly_connected_components(h, g, None, None, None, None, False)
    ref = oracle.wcc(s, d, ids.size)
    got = np.zeros(ids.size, dtype=np.int64)
    got[np.searchsorted(ids, verts.cpu().numpy())] = labels.cpu().numpy()
    pairs = set(zip(ref.tolist(), got.tolist()))
    assert len(pairs) == len(set(ref.tolist())) == len(set(got.tolist()))
