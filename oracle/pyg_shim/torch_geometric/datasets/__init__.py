"""ORACLE shim: import-time stubs only (src/data/data_setup.py:10); datasets need network."""


class _Unavailable:
    def __init__(self, *a, **k):
        raise RuntimeError("dataset download is outside the hot path (no network)")


class Planetoid(_Unavailable):
    pass


class TUDataset(_Unavailable):
    pass
