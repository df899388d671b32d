from gnnb200.nn import GINConv, global_add_pool, global_max_pool, global_mean_pool  # noqa: F401
