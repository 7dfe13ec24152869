"""GraphProperties (reference python/pylibcugraph/pylibcugraph/graph_properties.pyx)."""
from cugraph_b200 import _capi


class GraphProperties:
    def __init__(self, is_symmetric=False, is_multigraph=False):
        self.c = _capi.GraphPropertiesStruct(int(bool(is_symmetric)), int(bool(is_multigraph)))

    @property
    def is_symmetric(self):
        return bool(self.c.is_symmetric)

    @is_symmetric.setter
    def is_symmetric(self, v):
        self.c.is_symmetric = int(bool(v))

    @property
    def is_multigraph(self):
        return bool(self.c.is_multigraph)

    @is_multigraph.setter
    def is_multigraph(self, v):
        self.c.is_multigraph = int(bool(v))
