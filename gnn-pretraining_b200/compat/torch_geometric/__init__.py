"""gnnb200's stand-in for the ``torch_geometric`` names the reference imports (see gnnb200.compat)."""
__version__ = '0.0.0-gnnb200'
__gnnb200__ = True
from . import data, datasets, loader, nn, transforms, utils  # noqa: F401,E402
