from ppnn_b200.operators import *  # noqa: F401,F403
