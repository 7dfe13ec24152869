"""TEST INFRASTRUCTURE: `pylibcugraph.testing.utils` is the reference's own pure-Python helper module, loaded from where
it lies in the reference tree (nothing is copied)."""
import importlib.util as _u
import os as _os
import sys as _sys

_REF = _os.environ.get("REF", "/root/reference")
_spec = _u.spec_from_file_location("pylibcugraph.testing.utils",
                                   _os.path.join(_REF, "python", "pylibcugraph", "pylibcugraph", "testing", "utils.py"))
utils = _u.module_from_spec(_spec)
_sys.modules["pylibcugraph.testing.utils"] = utils
_spec.loader.exec_module(utils)
