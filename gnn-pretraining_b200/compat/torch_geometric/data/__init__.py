from gnnb200.data import Batch, Data  # noqa: F401
