"""TEST INFRASTRUCTURE: makes `import pylibcugraph` resolve to this repository's mirror (cugraph_b200.pylibcugraph), so
that the reference's own pylibcugraph tests run unmodified against it (oracle/ref_pytests/run.py)."""
from cugraph_b200.pylibcugraph import *  # noqa: F401,F403
from cugraph_b200.pylibcugraph import exceptions  # noqa: F401
from cugraph_b200.pylibcugraph import __version__, __git_commit__  # noqa: F401,E402
