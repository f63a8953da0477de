from .tensor_store import AsyncTensorStore, TensorStore  # noqa: F401
