"""TEST INFRASTRUCTURE: `pylibcugraph.utilities.api_tools` resolves to the mirror's module."""
from cugraph_b200.pylibcugraph.utilities.api_tools import *  # noqa: F401,F403
from cugraph_b200.pylibcugraph.utilities.api_tools import experimental_warning_wrapper  # noqa: F401
