"""Python surface mirroring `pylibcugraph` for the PageRank / BFS / SSSP path
(reference python/pylibcugraph/pylibcugraph/__init__.py).  Same names, argument order and error
behaviour; device arrays go in as anything exposing `__cuda_array_interface__` (torch CUDA tensors,
cupy, numba) and come back as torch CUDA tensors (cupy is not a dependency here)."""
from cugraph_b200.pylibcugraph.exceptions import FailedToConvergeError
from cugraph_b200.pylibcugraph.resource_handle import ResourceHandle
from cugraph_b200.pylibcugraph.graph_properties import GraphProperties
from cugraph_b200.pylibcugraph.graphs import SGGraph
from cugraph_b200.pylibcugraph.algorithms import (pagerank, personalized_pagerank, bfs, sssp, katz_centrality, hits,
                                                  weakly_connected_components, strongly_connected_components,
                                                  generate_rmat_edgelist)

__all__ = ["FailedToConvergeError", "ResourceHandle", "GraphProperties", "SGGraph",
           "pagerank", "personalized_pagerank", "bfs", "sssp", "katz_centrality", "hits", "weakly_connected_components", "strongly_connected_components", "generate_rmat_edgelist"]
