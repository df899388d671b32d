"""ORACLE shim: import-time stub (src/data/data_setup.py:11)."""


class NormalizeFeatures:
    def __call__(self, data):
        x = data.x
        data.x = x / x.sum(dim=-1, keepdim=True).clamp(min=1.0)
        return data
