"""A short run of emu/fuzz.py: random small multigraphs (id width, weight type, storage order, renumbering, every sweep
variant switch) through the emulated C ABI against the oracle.  `bash emu/run_asan.sh` runs the same under AddressSanitizer."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fuzz_emulated_library():
    env = dict(os.environ, FUZZ_SEED="11")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "emu", "fuzz.py"), "12"], capture_output=True, text=True, env=env,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "no mismatch" in r.stdout
