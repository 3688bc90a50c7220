from ppnn_b200.utility.utils import mesh_convertor, model_count  # noqa: F401
