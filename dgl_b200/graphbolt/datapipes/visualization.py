"""Rendering a datapipe chain (parity: /root/reference/python/dgl/graphbolt/datapipes/visualization.py: Node,
to_nodes, to_graph). ``to_graph`` returns a ``graphviz.Digraph`` when graphviz is importable and DOT text otherwise
(the DOT source is what the reference's Digraph holds as well)."""
from __future__ import annotations

from .utils import traverse_dps

__all__ = ["Node", "to_nodes", "to_graph"]


class Node:
    """One datapipe and the nodes it feeds."""

    def __init__(self, dp, name=None):
        self.dp = dp
        self.name = name or type(dp).__name__.replace("IterDataPipe", "")
        self.childs = set()
        self.parents = set()

    def add_parent(self, parent):
        self.parents.add(parent)
        parent.childs.add(self)

    def add_child(self, child):
        child.add_parent(self)

    def remove_parent(self, parent):
        self.parents.discard(parent)
        parent.childs.discard(self)

    def remove_child(self, child):
        child.remove_parent(self)

    def __repr__(self):
        return f"<Node {self.name} parents={sorted(p.name for p in self.parents)}>"

    def __hash__(self):
        return id(self.dp)

    def __eq__(self, other):
        return isinstance(other, Node) and other.dp is self.dp


def to_nodes(dp, *, debug=False):
    """All nodes of the chain ending at ``dp`` (upstream first), parents/childs linked."""
    nodes = {}

    def walk(level, child):
        for key, (pipe, ups) in level.items():
            node = nodes.setdefault(key, Node(pipe))
            if child is not None:
                child.add_parent(node)
            walk(ups, node)

    walk(traverse_dps(dp), None)
    ordered, done = [], set()

    def emit(node):
        if node in done:
            return
        done.add(node)
        for p in sorted(node.parents, key=lambda n: n.name):
            emit(p)
        ordered.append(node)

    for n in nodes.values():
        emit(n)
    # disambiguate equal names
    counts = {}
    for n in ordered:
        counts[n.name] = counts.get(n.name, 0) + 1
        if counts[n.name] > 1:
            n.name = f"{n.name}_{counts[n.name]}"
    return ordered


def to_graph(dp, *, debug=False):
    nodes = to_nodes(dp, debug=debug)
    edges = [(p.name, n.name) for n in nodes for p in sorted(n.parents, key=lambda x: x.name)]
    try:
        import graphviz
        g = graphviz.Digraph(graph_attr={"labelloc": "t", "label": "datapipes"})
        for n in nodes:
            g.node(n.name, n.name, shape="box")
        for a, b in edges:
            g.edge(a, b)
        return g
    except ImportError:
        lines = ["digraph {", '\tlabelloc=t label=datapipes']
        lines += [f'\t"{n.name}" [label="{n.name}" shape=box]' for n in nodes]
        lines += [f'\t"{a}" -> "{b}"' for a, b in edges]
        return "\n".join(lines + ["}"])
