"""cugraph_b200 — B200-native replacement for cuGraph's PageRank / BFS / SSSP primitive hot path.

Layout:
  csrc/          CUDA kernels + the C-ABI (`libcugraph_c.so` drop-in, headers in /include/cugraph_c)
  pylibcugraph/  Python mirror of the reference's `pylibcugraph` surface for this path (ctypes)
  build.py       nvcc build for sm_100a

The product path has no CPU fallback: importing `cugraph_b200.pylibcugraph` loads the CUDA library
and raises if it is missing.
"""
__version__ = "0.1.0"
