"""graphneuralnetworks.jl_b200 — B200-native (sm_100a) message-passing engine behind GNNlib.jl's
`propagate` / `apply_edges` / `aggregate_neighbors` API.

Import name: ``gnnb200`` (the directory name contains a dot, so the repo-root shim ``gnnb200.py`` loads this
package under that name).  Product code = csrc/ (CUDA kernels + C ABI, built into lib/libgnnb200.so) and the
host-side mirror of the reference interface in this package.  Nothing here imports ``oracle/``.
"""
from . import _lib
from ._lib import GNNBError, device_count, launch_count, version
from .graph import (GNNGraph, add_self_loops, batch, colmajor, degree, edge_index, get_edge_weight,
                    graph_indicator, jl_randn, jl_zeros, rmat_graph, rows, set_edge_weight, unrows)
from .msgpass import (Fix1, aggregate_neighbors, apply_edges, check_num_edges, check_num_nodes, copy_xi, copy_xj,
                      e_mul_xj, expand_srcdst, mean, propagate, softmax_edge_neighbors, w_mul_xj, xi_dot_xj,
                      xi_sub_xj, xj_sub_xi)
from .layers import (AGNNConv, GATConv, GATv2Conv, GCNConv, GINConv, GatedGraphConv, GraphConv, SAGEConv, SGConv,
                     TAGConv, TransformerConv, agnn_conv, gat_conv, gat_message, gated_graph_conv, gatv2_conv,
                     gatv2_message, gcn_conv, gin_conv, graph_conv, identity, relu, sage_conv, sg_conv, sgc_conv,
                     tag_conv, transformer_conv)
from .layers_more import (CGConv, ChebConv, DConv, EdgeConv, EGNNConv, GMMConv, MEGNetConv, NNConv,
                          ResGatedGraphConv, cg_conv, cheb_conv, d_conv, edge_conv, egnn_conv, gmm_conv, megnet_conv,
                          nn_conv, res_gated_graph_conv)
from .readout import (broadcast_edges, broadcast_nodes, global_attention_pool, global_pool, reduce_edges, reduce_nodes,
                      softmax_edges, softmax_nodes)
from .transform import csr, remove_multi_edges, remove_self_loops, sort_edge_index, to_bidirected, unbatch
from .sampling import NeighborLoader, induced_subgraph, sample_edge_ids, sample_neighbors
from .query import (adjacency_list, adjacency_matrix, has_multi_edges, has_self_loops, inneighbors, is_bidirected,
                    outneighbors)

__all__ = [n for n in dir() if not n.startswith("_")]
