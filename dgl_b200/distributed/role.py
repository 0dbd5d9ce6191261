"""Process roles inside a distributed job (parity: /root/reference/python/dgl/distributed/role.py).
Roles come from the launcher's environment: servers (DGL_ROLE=server), trainers ("default"), samplers."""
from __future__ import annotations

import os

from .service_ids import GET_ROLE, GET_ROLE_MSG, REG_ROLE_MSG, REGISTER_ROLE  # noqa: F401

_ROLE = {"name": None, "rank": 0, "global_rank": 0, "num": 1}
# registry views with the reference's names (role.py:114-123): role -> {client id -> rank inside the role} / global rank
PER_ROLE_RANK = {}
GLOBAL_RANK = {}
CUR_ROLE = None
IS_STANDALONE = False


def init_role(role):
    global CUR_ROLE, IS_STANDALONE
    _ROLE["name"] = CUR_ROLE = role
    IS_STANDALONE = os.environ.get("DGL_DIST_MODE", "standalone") == "standalone"
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        _ROLE["rank"] = _ROLE["global_rank"] = dist.get_rank()
        _ROLE["num"] = dist.get_world_size()
    else:
        _ROLE["rank"] = _ROLE["global_rank"] = int(os.environ.get("RANK", "0"))
        _ROLE["num"] = int(os.environ.get("WORLD_SIZE", "1"))
    PER_ROLE_RANK.setdefault(role, {})[_ROLE["global_rank"]] = _ROLE["rank"]
    GLOBAL_RANK[_ROLE["global_rank"]] = _ROLE["global_rank"]


def get_role():
    return _ROLE["name"] or os.environ.get("DGL_ROLE", "default")


def get_trainer_rank():
    return _ROLE["rank"]


def get_global_rank():
    return _ROLE["global_rank"]


def get_rank():
    """Rank of this process inside its role (== ``get_trainer_rank`` for trainers; reference role.py:244)."""
    return _ROLE["rank"]


def get_num_trainers():
    return _ROLE["num"]


def __getattr__(name):
    # class-based request / response types of this service live in messages.py (lazy: avoids an import cycle)
    if name.startswith("__"):
        raise AttributeError(name)
    from . import messages
    if name in messages.__all__:
        return getattr(messages, name)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
