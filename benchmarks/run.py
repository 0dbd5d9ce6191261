#!/usr/bin/env python
"""Run the benchmark suite:  python benchmarks/run.py [--device cuda] [--filter kernel] [--out results.json]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cpu")
    ap.add_argument("--filter", default="")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    import suite
    results = []
    for group, name, fn, pg in suite.REGISTRY:
        if a.filter and a.filter not in f"{group}.{name}":
            continue
        for params in suite.grid(pg):
            try:
                sec = fn(a.device, **params)
                rec = {"group": group, "name": name, "params": params, "device": a.device, "seconds": sec}
            except Exception as e:  # noqa: BLE001 - a benchmark failing must not stop the sweep
                rec = {"group": group, "name": name, "params": params, "device": a.device, "error": repr(e)}
            results.append(rec)
            print(json.dumps(rec), flush=True)
    if a.out:
        with open(a.out, "w") as f:
            json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
