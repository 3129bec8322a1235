"""Same exports as the reference's gnns/__init__.py:1-7."""
from .ggnn import sparse_ggnn_layer  # noqa: F401
from .gnn_edge_mlp import sparse_gnn_edge_mlp_layer  # noqa: F401
from .gnn_film import sparse_gnn_film_layer  # noqa: F401
from .rgat import sparse_rgat_layer  # noqa: F401
from .rgdcn import sparse_rgdcn_layer  # noqa: F401
from .rgcn import rgcn_layer_stack, sparse_rgcn_layer  # noqa: F401
from .rgin import sparse_rgin_layer  # noqa: F401
