"""Argument validation helpers (parity: /root/reference/python/dgl/utils/checks.py).

``prepare_tensor`` / ``prepare_tensor_dict`` live in :mod:`.internal`; the rest of the reference's
checks module is here."""
from __future__ import annotations

from collections.abc import Mapping

import torch

from ..base import DGLError
from .internal import prepare_tensor, prepare_tensor_dict

__all__ = ["prepare_tensor_or_dict", "parse_edges_arg_to_eid", "check_all_same_idtype", "check_device",
           "check_all_same_device", "check_all_same_schema", "check_all_same_schema_for_keys", "check_valid_idtype",
           "is_sorted_srcdst"]


def prepare_tensor_or_dict(g, data, name):
    """Tensor in, id tensor of ``g`` out; mapping in, mapping of id tensors out (checks.py:97)."""
    if isinstance(data, Mapping):
        return prepare_tensor_dict(g, data, name)
    return prepare_tensor(g, data, name)


def parse_edges_arg_to_eid(g, edges, etid, argname="edges"):
    """An ``edges`` argument is either edge ids or an (u, v) endpoint pair; resolve to edge ids (checks.py:120)."""
    if isinstance(edges, tuple):
        u = prepare_tensor(g, edges[0], f"{argname}[0]")
        v = prepare_tensor(g, edges[1], f"{argname}[1]")
        return g.edge_ids(u, v, etype=g.canonical_etypes[etid])
    return prepare_tensor(g, edges, argname)


def _all_equal(items, what, name, fmt):
    first = None
    for i, it in enumerate(items):
        val = what(it)
        if i == 0:
            first = val
        elif val != first:
            raise DGLError(fmt.format(name=name, i=i, want=first, got=val))


def check_all_same_idtype(glist, name):
    _all_equal(glist, lambda g: g.idtype, name, "Expect {name}[{i}] to have {want} type ID, but got {got}.")


def check_all_same_device(glist, name):
    _all_equal(glist, lambda g: g.device, name, "Expect {name}[{i}] to be on device {want}, but got {got}.")


def check_device(data, device):
    """True iff the tensor (or every tensor of the dict) lives on ``device``."""
    device = torch.device(device)
    vals = data.values() if isinstance(data, Mapping) else [data]
    return all(v.device == device for v in vals)


def check_all_same_schema(schemas, name):
    _all_equal(schemas, lambda s: s, name,
               "Expect all graphs to have the same schema on {name}, but graph {i} got\n\t{got}\nwhich is different "
               "from\n\t{want}.")


def check_all_same_schema_for_keys(schemas, keys, name):
    keys = set(keys)
    for i, schema in enumerate(schemas):
        if not keys.issubset(schema.keys()):
            raise DGLError(f"Expect all graphs to have keys {keys} on {name}, but graph {i} got keys {schema.keys()}.")
    _all_equal(schemas, lambda s: {k: s[k] for k in keys}, name,
               "Expect all graphs to have the same schema for keys " + str(keys).replace("{", "{{").replace("}", "}}") +
               " on {name}, but graph {i} got \n\t{got}\n which is different from\n\t{want}.")


def check_valid_idtype(idtype):
    if idtype not in (None, torch.int32, torch.int64):
        raise DGLError(f"Expect idtype to be a framework object of int32/int64, got {idtype}")


def is_sorted_srcdst(src, dst, num_src=None, num_dst=None):
    """(src_sorted, dst_sorted_within_src) of an edge list: is it in ascending src-major order, and if so are the
    destinations of every source ascending too (checks.py:257)."""
    src, dst = torch.as_tensor(src), torch.as_tensor(dst)
    if src.numel() <= 1:
        return True, True
    ds = src[1:] - src[:-1]
    row_sorted = bool((ds >= 0).all())
    if not row_sorted:
        return False, False
    same = ds == 0
    col_sorted = bool(((dst[1:] >= dst[:-1]) | ~same).all())
    return row_sorted, col_sorted
