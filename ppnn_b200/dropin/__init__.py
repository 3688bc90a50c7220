"""Directory to put FIRST on PYTHONPATH so the reference scripts import ppnn_b200."""
