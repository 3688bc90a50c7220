"""pyplot subset used by train2D.py:101-108 (subplots / pcolormesh / colorbar) and by
torch.utils.tensorboard's figure_to_image (close)."""
import numpy as np


class _Canvas:
    def __init__(self, fig):
        self.fig = fig

    def get_width_height(self):
        return self.fig.size


class _Axes:
    def pcolormesh(self, data, **kw):
        return np.asarray(data)


class Figure:
    def __init__(self, figsize=(4, 4)):
        self.size = (int(figsize[0] * 8), int(figsize[1] * 8))
        self.canvas = _Canvas(self)

    def colorbar(self, mappable, ax=None):
        return None


def subplots(nrows=1, ncols=1, figsize=(4, 4), **kw):
    fig = Figure(figsize)
    n = nrows * ncols
    return fig, (_Axes() if n == 1 else [_Axes() for _ in range(n)])


def close(fig=None):
    return None
