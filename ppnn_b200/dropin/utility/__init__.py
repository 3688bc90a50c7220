"""PYTHONPATH shim for the reference `utility` package."""
