#!/usr/bin/env python
"""Upgrade an old partition config whose edge types are bare names to canonical ``src:etype:dst`` keys (reference
tools/change_etype_to_canonical_etype.py)."""
import argparse
import json

DELIM = ":"


def is_old_version(config):
    return any(DELIM not in k for k in config["etypes"])


def convert_conf(part_config, canonical=None):
    """``canonical``: optional {etype: (src, etype, dst)}; homogeneous configs are upgraded without it."""
    with open(part_config) as f:
        config = json.load(f)
    if not is_old_version(config):
        return False
    ntypes = list(config["ntypes"])
    mapping = {}
    for et in config["etypes"]:
        if len(ntypes) == 1:
            mapping[et] = DELIM.join((ntypes[0], et, ntypes[0]))
        elif canonical and et in canonical:
            mapping[et] = DELIM.join(canonical[et])
        else:
            raise ValueError(f"cannot infer the canonical form of edge type {et!r}; pass it explicitly")
    config["etypes"] = {mapping[k]: v for k, v in config["etypes"].items()}
    config["edge_map"] = {mapping.get(k, k): v for k, v in config["edge_map"].items()}
    with open(part_config, "w") as f:
        json.dump(config, f, indent=4)
    return True


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--part-config", "--part_config", dest="cfg", required=True)
    a = ap.parse_args()
    print("converted" if convert_conf(a.cfg) else "already canonical")
