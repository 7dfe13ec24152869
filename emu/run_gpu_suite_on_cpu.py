"""Run the `-m gpu` test suite against the CPU emulation build of the library (tests/emu_py.py): the tests' Python logic
and the kernels' LOGIC at the test sizes, without a GPU.  Multi-GPU (NCCL) tests are left out; the RMAT-24 certificates
run at CUGRAPH_B200_FULL_SCALE (default here: 10).  Not a substitute for the GPU run — timing, memory ordering and
scheduling only exist there — but it finds everything else first.
    python emu/run_gpu_suite_on_cpu.py [pytest args]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("CUGRAPH_B200_FULL_SCALE", "10")

import pytest  # noqa: E402
from tests.emu_py import emulated_python_surface  # noqa: E402

with emulated_python_surface():
    sys.exit(pytest.main([os.path.join(ROOT, "tests"), "-q", "-m", "gpu", "--deselect", "tests/test_mg_gpu.py",
                          "--deselect", "tests/test_reference_c_tests_gpu.py",
                          "--deselect", "tests/test_zz_late_additions_gpu.py::test_reference_c_test_program_on_gpu",
                          "-p", "no:cacheprovider"] + sys.argv[1:]))
