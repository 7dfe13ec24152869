"""Device RMAT generator on the GPU (cugraph_b200.generators.rmat_edgelist -> cugraph_b200_generate_rmat_edgelist) against
its numpy twin, bit for bit."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("scale,seed,clip", [(16, 0, False), (20, 31337, True)])
def test_rmat_device_vs_numpy(scale, seed, clip):
    from cugraph_b200.generators import rmat_edgelist
    from oracle.rmat import rmat_edgelist_counter
    n = 300_000
    s, d = rmat_edgelist(scale, n, seed=seed, clip_and_flip=clip)
    rs, rd = rmat_edgelist_counter(scale, n, seed=seed, clip_and_flip=clip)
    assert np.array_equal(s.cpu().numpy(), rs) and np.array_equal(d.cpu().numpy(), rd)
