from .tensor_store import TensorStore  # noqa: F401
