from ppnn_b200.utility.utils import mesh_convertor, model_count, numpy2string  # noqa: F401
