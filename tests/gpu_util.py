"""Helpers shared by the -m gpu tests: graphs are created and algorithms called through the
pylibcugraph-compatible surface, i.e. through the C-ABI of libcugraph_c.so."""
import numpy as np


def have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def make_graph(src, dst, weights=None, store_transposed=False, renumber=True, symmetric=False,
               vertex_dtype=np.int32, weight_dtype=np.float32, vertices=None, **kw):
    import torch
    from cugraph_b200 import pylibcugraph as plc
    h = plc.ResourceHandle()
    s = torch.as_tensor(np.asarray(src, dtype=vertex_dtype)).cuda()
    d = torch.as_tensor(np.asarray(dst, dtype=vertex_dtype)).cuda()
    w = None if weights is None else torch.as_tensor(np.asarray(weights, dtype=weight_dtype)).cuda()
    v = None if vertices is None else torch.as_tensor(np.asarray(vertices, dtype=vertex_dtype)).cuda()
    symmetric = symmetric or bool(kw.get("symmetrize"))   # the reference rejects symmetrize without the property (graph_sg.cpp:737-742)
    g = plc.SGGraph(h, plc.GraphProperties(is_symmetric=symmetric, is_multigraph=True), s, d, weight_array=w,
                    store_transposed=store_transposed, renumber=renumber, vertices_array=v, **kw)
    return h, g


def by_vertex(verts, vals, n=None):
    """Scatter a (vertices, values) result into an array indexed by external id."""
    v = verts.cpu().numpy()
    x = vals.cpu().numpy()
    n = int(v.max()) + 1 if n is None else n
    out = np.zeros(n, dtype=x.dtype)
    out[v] = x
    return out
