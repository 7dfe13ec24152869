"""Exception classes (reference python/pylibcugraph/pylibcugraph/exceptions.py)."""


class FailedToConvergeError(Exception):
    """An iterative algorithm did not converge within its iteration budget."""
