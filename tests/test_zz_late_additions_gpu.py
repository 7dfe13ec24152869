"""GPU tests of the entry points that were added AFTER the round's last GPU call (validated under the emulation and by the
reference's own test programs on CPU only): eigenvector centrality and the degree functions.  The file name sorts last on
purpose: under `pytest -x` everything that has already run on hardware runs first."""
import os
import subprocess

import numpy as np
import pytest

from tests.test_reference_c_tests_cpu import ROOT, check_output
from tests.test_siblings_gpu import _graph

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["eigenvector_centrality", "degrees"])
def test_reference_c_test_program_on_gpu(name):
    exe = os.path.join(ROOT, "oracle", "_ref", f"ref_{name}_test_gpu")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/ref_*_test_gpu not built (needs the reference sources at build time)")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    check_output(name, r)


def test_eigenvector_centrality_gpu():
    import oracle
    plc, h, g, ids, s, d, _ = _graph()
    verts, vals = plc.eigenvector_centrality(h, g, 1e-7, 1000, False)
    ref, _ = oracle.eigenvector(s, d, ids.size, None, epsilon=1e-7, max_iterations=1000)
    got = np.zeros(ids.size)
    got[np.searchsorted(ids, verts.cpu().numpy())] = vals.cpu().numpy()
    np.testing.assert_allclose(got, ref, rtol=2e-3, atol=1e-8)


def test_degrees_gpu():
    plc, h, g, ids, s, d, _ = _graph()
    v, din, dout = plc.degrees(h, g, None, False)
    vi = np.searchsorted(ids, v.cpu().numpy())
    assert np.array_equal(din.cpu().numpy(), np.bincount(d, minlength=ids.size)[vi])
    assert np.array_equal(dout.cpu().numpy(), np.bincount(s, minlength=ids.size)[vi])
