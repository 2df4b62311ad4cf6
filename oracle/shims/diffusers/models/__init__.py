from . import attention, embeddings  # noqa: F401
