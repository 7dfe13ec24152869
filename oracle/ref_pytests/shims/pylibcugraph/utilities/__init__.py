"""TEST INFRASTRUCTURE: `pylibcugraph.utilities` resolves to the mirror's (oracle/ref_pytests/run.py)."""
