class NormalizeFeatures:
    def __call__(self, data):
        data.x = data.x / data.x.sum(dim=-1, keepdim=True).clamp(min=1.0)
        return data
