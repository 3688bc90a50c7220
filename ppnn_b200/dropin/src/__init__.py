"""PYTHONPATH shim for the `src.` spelling used by Bv/ and post_process/."""
