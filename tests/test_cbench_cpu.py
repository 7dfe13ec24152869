"""scripts/cbench.cu (the Python-free development probe for the GPU box) compiled against the CPU emulation shim and run at
a toy scale: its own logic — device RMAT generator, graph creation through the C ABI, the sweep parity hook, PageRank, BFS
and SSSP calls, JSON output — checked without a GPU."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cbench_under_emulation(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "emu"))
    import build_emu
    try:
        lib = build_emu.build()
    except Exception as e:
        pytest.skip(f"emulation build unavailable: {e}")
    exe = str(tmp_path / "cbench_emu")
    cmd = ["g++", "-std=c++17", "-O1", "-DB200_HOST_EMU", "-I", os.path.join(ROOT, "emu"), "-I", os.path.join(ROOT, "include"),
           "-Wno-attributes", "-x", "c++", os.path.join(ROOT, "scripts", "cbench.cu"), "-o", exe, "-L", os.path.dirname(lib),
           "-l:" + os.path.basename(lib), "-Wl,-rpath," + os.path.dirname(lib)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    env = dict(os.environ, CUGRAPH_B200_SWEEP_MIN_EDGES="0")
    r = subprocess.run([exe, "8", "all", "1"], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = {}
    for ln in r.stdout.splitlines():
        d = json.loads(ln)
        lines[d["mode"]] = d
    assert set(lines) == {"sweep", "pagerank", "bfs", "sssp"}
    assert lines["sweep"]["bad_rows_ge32"] == 0 and lines["sweep"]["bad_rows_lt32"] == 0 and lines["sweep"]["sweep_ms"] > 0
    assert lines["pagerank"]["pagerank100_ms"] > 0
    assert lines["bfs"]["reached_last"] == lines["sssp"]["reached_last"] > 50   # same component through both algorithms
