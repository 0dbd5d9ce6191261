"""Small helpers used across the package (parity: python/dgl/utils/internal.py, checks.py)."""
from __future__ import annotations

import torch

from ..base import DGLError

__all__ = ["expand_as_pair", "check_eq_shape", "prepare_tensor", "prepare_tensor_dict", "recursive_apply",
           "context_of", "dtype_of", "extract_node_subframes", "extract_edge_subframes", "set_new_frames"]


def expand_as_pair(input_, g=None):
    """Return (src_feat, dst_feat); for blocks dst nodes are the prefix of src nodes
    (reference utils/internal.py:expand_as_pair)."""
    if isinstance(input_, tuple):
        return input_
    if g is not None and g.is_block:
        if isinstance(input_, dict):
            input_dst = {k: v[: g.number_of_dst_nodes(k)] for k, v in input_.items() if k in g.dsttypes}
        else:
            input_dst = input_[: g.number_of_dst_nodes()]
        return input_, input_dst
    return input_, input_


def check_eq_shape(input_):
    srcdata, dstdata = expand_as_pair(input_)
    if tuple(srcdata.shape[1:]) != tuple(dstdata.shape[1:]):
        raise DGLError(f"The feature shape of source nodes: {tuple(srcdata.shape[1:])} should be equal to the "
                       f"feature shape of destination nodes: {tuple(dstdata.shape[1:])}.")


def prepare_tensor(g, data, name):
    if torch.is_tensor(data):
        if data.dtype != g.idtype:
            data = data.to(g.idtype)
        return data.to(g.device) if data.device != g.device and not (data.device.type == "cpu" and g.is_pinned()) else data
    return torch.as_tensor(data, dtype=g.idtype, device=g.device).reshape(-1)


def prepare_tensor_dict(g, data, name):
    return {k: prepare_tensor(g, v, f"{name}[{k}]") for k, v in data.items()}


def recursive_apply(data, fn, *args, **kwargs):
    if isinstance(data, dict):
        return {k: recursive_apply(v, fn, *args, **kwargs) for k, v in data.items()}
    if isinstance(data, (list, tuple)):
        return type(data)(recursive_apply(v, fn, *args, **kwargs) for v in data)
    return fn(data, *args, **kwargs)


def context_of(data):
    if isinstance(data, dict):
        return next(iter(data.values())).device
    if isinstance(data, (list, tuple)):
        return context_of(data[0])
    return data.device


def dtype_of(data):
    if isinstance(data, dict):
        return next(iter(data.values())).dtype
    return data.dtype


def extract_node_subframes(graph, nodes_or_device, store_ids=True):
    from ..base import NID
    from ..frame import Frame
    frames = []
    for i, fr in enumerate(graph._node_frames):
        if nodes_or_device is None or not isinstance(nodes_or_device, (list, tuple)):
            frames.append(fr.clone())
        else:
            ind = nodes_or_device[i]
            sub = fr.subframe(ind)
            if store_ids:
                sub[NID] = ind
            frames.append(sub)
    return frames


def extract_edge_subframes(graph, edges, store_ids=True):
    from ..base import EID
    frames = []
    for i, fr in enumerate(graph._edge_frames):
        if edges is None:
            frames.append(fr.clone())
        else:
            sub = fr.subframe(edges[i])
            if store_ids:
                sub[EID] = edges[i]
            frames.append(sub)
    return frames


def set_new_frames(graph, *, node_frames=None, edge_frames=None):
    if node_frames is not None:
        graph._node_frames = node_frames
    if edge_frames is not None:
        graph._edge_frames = edge_frames


from .containers import *  # noqa: E402,F401,F403  (Index, dict views, relabel family, thread-count / NUMA queries ...)
