"""The reference's OWN pylibcugraph tests for this path — python/pylibcugraph/pylibcugraph/tests/{test_pagerank,test_sssp,
test_graph_sg,test_katz_centrality,test_connected_components,test_rmat,test_structure,test_utils,test_version,test_eigenvector_centrality}.py with their conftest.py, unmodified, from where they lie under /root/reference — run against this
repository's pylibcugraph mirror (oracle/ref_pytests/run.py: `pylibcugraph` and `cupy` resolve to small stand-ins, the
library is the CPU emulation build): karate / dolphins / Simple_1 / Simple_2 PageRank and SSSP goldens with the reference's
tolerances, GraphProperties / ResourceHandle / SGGraph construction and the exception types for invalid input.  Skipped
where the reference sources are absent (the GPU box)."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("REF", "/root/reference")


def test_reference_pylibcugraph_tests():
    if not os.path.isdir(os.path.join(REF, "python", "pylibcugraph", "pylibcugraph", "tests")):
        pytest.skip("reference sources not present")
    pytest.importorskip("pandas")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "ref_pytests", "run.py")], capture_output=True, text=True,
                       timeout=900, cwd=ROOT)
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0, tail
    m = re.search(r"(\d+) passed", r.stdout)
    assert m and int(m.group(1)) == 322 and "failed" not in r.stdout, tail   # 4 + 4 + 8 + 1 + 11 + 288 + 1 + 3 + 1 + 1 (deselected: cudf, SCC)
