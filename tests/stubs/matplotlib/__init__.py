"""Minimal stand-in for matplotlib (not installed in this image) so that the reference's
train2D.py -- which imports matplotlib.pyplot at line 5 and draws two pcolormesh panels per
tensorboard figure (train2D.py:101-108, 251-259) -- can run UNCHANGED in the tests.
It renders every figure as a blank canvas; nothing numerical depends on it."""
__version__ = "0.0-stub"
