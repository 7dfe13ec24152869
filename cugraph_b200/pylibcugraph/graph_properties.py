"""GraphProperties (reference python/pylibcugraph/pylibcugraph/graph_properties.pyx)."""
import operator

from cugraph_b200 import _capi


def _flag(value, name):
    """What Cython's assignment to a bool_t struct field accepts: bool / int-like; anything else is a TypeError."""
    try:
        return int(bool(operator.index(value)))
    except TypeError:
        raise TypeError(f"{name} must be a bool or an integer, got {type(value).__name__}") from None


class GraphProperties:
    def __init__(self, is_symmetric=False, is_multigraph=False):
        self.c = _capi.GraphPropertiesStruct(_flag(is_symmetric, "is_symmetric"), _flag(is_multigraph, "is_multigraph"))

    @property
    def is_symmetric(self):
        return bool(self.c.is_symmetric)

    @is_symmetric.setter
    def is_symmetric(self, v):
        self.c.is_symmetric = _flag(v, "is_symmetric")

    @property
    def is_multigraph(self):
        return bool(self.c.is_multigraph)

    @is_multigraph.setter
    def is_multigraph(self, v):
        self.c.is_multigraph = _flag(v, "is_multigraph")

    # pickle support, as the reference (graph_properties.pyx: __getnewargs_ex__ / __getstate__ / __setstate__)
    def __reduce__(self):
        return (GraphProperties, (self.is_symmetric, self.is_multigraph))
