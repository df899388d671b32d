class _NeedsNetwork:
    def __init__(self, *args, **kwargs):
        raise RuntimeError('dataset download/preprocessing is outside the hot path')


class Planetoid(_NeedsNetwork):
    pass


class TUDataset(_NeedsNetwork):
    pass
