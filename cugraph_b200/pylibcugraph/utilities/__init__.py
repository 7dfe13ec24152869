"""Helper namespace of the mirror (reference python/pylibcugraph/pylibcugraph/utilities/)."""
