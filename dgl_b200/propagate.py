"""Message propagation along traversal frontiers (parity: /root/reference/python/dgl/propagate.py)."""
from __future__ import annotations

from . import traversal as trv

__all__ = ["prop_nodes", "prop_edges", "prop_nodes_bfs", "prop_nodes_topo", "prop_edges_dfs"]


def prop_nodes(graph, nodes_generator, message_func="default", reduce_func="default", apply_node_func="default"):
    graph.prop_nodes(nodes_generator, message_func, reduce_func, apply_node_func if apply_node_func != "default" else None)


def prop_edges(graph, edges_generator, message_func="default", reduce_func="default", apply_node_func="default"):
    graph.prop_edges(edges_generator, message_func, reduce_func, apply_node_func if apply_node_func != "default" else None)


def prop_nodes_bfs(graph, source, message_func, reduce_func, reverse=False, apply_node_func=None):
    gen = trv.bfs_nodes_generator(graph, source, reverse)
    graph.prop_nodes(gen[1:], message_func, reduce_func, apply_node_func)


def prop_nodes_topo(graph, message_func, reduce_func, reverse=False, apply_node_func=None):
    gen = trv.topological_nodes_generator(graph, reverse)
    graph.prop_nodes(gen[1:], message_func, reduce_func, apply_node_func)


def prop_edges_dfs(graph, source, message_func, reduce_func, reverse=False, has_reverse_edge=False,
                   has_nontree_edge=False, apply_node_func=None):
    gen = trv.dfs_labeled_edges_generator(graph, source, reverse, has_reverse_edge, has_nontree_edge, return_labels=False)
    graph.prop_edges(gen, message_func, reduce_func, apply_node_func)
