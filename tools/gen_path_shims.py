"""Generate the alias table (dgl_b200/_alias_table.py) for reference module paths that have no same-named module in dgl_b200.

Usage: python tools/gen_path_shims.py /root/reference/python/dgl [--write]

For every ``*.py`` under the reference package (MXNet/TensorFlow backends and the FFI layer excluded) whose dotted
path does not import from ``dgl_b200``, the public top-level names are read with ``ast`` (names only) and resolved
against a list of candidate dgl_b200 modules (the parent packages of the path plus a few known homes). A shim is
written only when every name resolves; unresolved names are printed so they can be implemented first."""
import argparse
import ast
import importlib
import os
import sys
import warnings

SKIP = ("mxnet", "tensorflow", "_ffi", "_deprecate", "backend/", "_dataloading", "__pycache__")
EXTRA_HOMES = {
    "graphbolt": ["dgl_b200.graphbolt", "dgl_b200.graphbolt.compat", "dgl_b200.graphbolt.impl", "dgl_b200.graphbolt.utils",
                  "dgl_b200.graphbolt.internal", "dgl_b200.graphbolt.datapipes", "dgl_b200.graphbolt.minibatch",
                  "dgl_b200.graphbolt.itemset", "dgl_b200.graphbolt.base"],
    "distributed": ["dgl_b200.distributed", "dgl_b200.distributed.compat", "dgl_b200.distributed.rpc_api",
                    "dgl_b200.distributed.dist_graph", "dgl_b200.distributed.kvstore", "dgl_b200.distributed.partition",
                    "dgl_b200.distributed.optim", "dgl_b200.optim", "dgl_b200.nn"],
    "optim": ["dgl_b200.optim", "dgl_b200.optim.sparse_optim"],
    "distgnn": ["dgl_b200.distgnn", "dgl_b200.partition"],
    "data": ["dgl_b200.data"],
    "nn": ["dgl_b200.nn", "dgl_b200.nn.conv", "dgl_b200.nn.pytorch"],
    "dataloading": ["dgl_b200.dataloading", "dgl_b200.distributed"],
    "sparse": ["dgl_b200.sparse"],
    "": ["dgl_b200"],
}


def public_names(path):
    tree = ast.parse(open(path).read())
    return sorted(n.name for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and not n.name.startswith("_"))


def write_table(path, table):
    """One generated module instead of one stub file per path: groups of source modules + {alias: (group, names)}."""
    groups = {}
    for mod, (sources, _) in table.items():
        groups.setdefault(tuple(sources), []).append(mod)
    gid = {s: f"g{i}" for i, (s, _) in enumerate(sorted(groups.items(), key=lambda kv: -len(kv[1])))}
    with open(path, "w") as fh:
        fh.write('"""Reference module paths that this package does not lay out as files: ``dgl_b200.<key>`` is synthesised on '
                 'import by\n``_pathshim.AliasFinder`` and resolves each listed name lazily in the modules of its source group. '
                 'Generated from the\nreference tree by tools/gen_path_shims.py (names only); checked by '
                 'tests/test_reference_paths.py."""\n\nSOURCE_GROUPS = {\n')
        for s, g in sorted(gid.items(), key=lambda kv: int(kv[1][1:])):
            fh.write(f"    {g!r}: {list(s)!r},\n")
        fh.write("}\n\nALIASES = {\n")
        for mod in sorted(table):
            sources, names = table[mod]
            fh.write(f"    {mod!r}: ({gid[tuple(sources)]!r}, {names!r}),\n")
        fh.write("}\n")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("reference_pkg")
    ap.add_argument("--write", action="store_true")
    args = ap.parse_args()
    warnings.filterwarnings("ignore")
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, here)
    unresolved = {}
    written = []
    table = {}
    for dp, _, files in os.walk(args.reference_pkg):
        for f in sorted(files):
            if not f.endswith(".py") or f == "__init__.py" or f.startswith("_"):
                continue
            rel = os.path.relpath(os.path.join(dp, f), args.reference_pkg)
            if any(s in rel for s in SKIP):
                continue
            dotted = rel[:-3].replace(os.sep, ".")
            target = os.path.join(here, "dgl_b200", rel)
            pkg_dir_as_module = os.path.join(here, "dgl_b200", rel[:-3])
            if os.path.exists(target) or os.path.isdir(pkg_dir_as_module):
                continue
            names = public_names(os.path.join(dp, f))
            # a submodule named like a function of its parent package would replace that function on import
            try:
                import types
                parent_mod = importlib.import_module("dgl_b200" + ("." + ".".join(dotted.split(".")[:-1]) if "." in dotted else ""))
                clash = getattr(parent_mod, dotted.split(".")[-1], None)
                if clash is not None and not isinstance(clash, types.ModuleType):
                    continue
            except Exception:  # noqa: BLE001
                pass
            top = dotted.split(".")[0] if "." in dotted else ""
            parents = []
            parts = dotted.split(".")[:-1]
            while parts:
                parents.append("dgl_b200." + ".".join(parts))
                parts = parts[:-1]
            homes = parents + [h for h in EXTRA_HOMES.get(top, []) if h not in parents] + ["dgl_b200"]
            good_homes = []
            for h in homes:
                try:
                    importlib.import_module(h)
                    good_homes.append(h)
                except Exception:  # noqa: BLE001
                    pass
            missing = []
            for n in names:
                if not any(hasattr(importlib.import_module(h), n) for h in good_homes):
                    missing.append(n)
            if missing:
                unresolved[rel] = missing
                continue
            # every intermediate package must exist
            ok = True
            d = os.path.dirname(target)
            to_make = []
            while not os.path.isdir(d):
                to_make.append(d)
                d = os.path.dirname(d)
            if any(os.path.exists(m + ".py") for m in to_make):
                ok = False                                   # a parent is a plain module here: promote it to a package first
            if not ok:
                unresolved[rel] = ["<parent is a module>"]
                continue
            if args.write:
                for m in reversed(to_make):          # intermediate PACKAGES stay real directories
                    os.makedirs(m, exist_ok=True)
                    init = os.path.join(m, "__init__.py")
                    if not os.path.exists(init):
                        relpkg = os.path.relpath(m, os.path.join(here, "dgl_b200")).replace(os.sep, ".")
                        with open(init, "w") as fh:
                            fh.write(f'"""Reference package path ``dgl.{relpkg}`` (modules are synthesised, see _pathshim)."""\n')
            table[dotted] = (good_homes, names)
            written.append(rel)
    if args.write:
        write_table(os.path.join(here, "dgl_b200", "_alias_table.py"), table)
    print("aliases:", len(written))
    for r in written:
        print("  ", r)
    print("unresolved:")
    for r, m in sorted(unresolved.items()):
        print("  ", r, m)


if __name__ == "__main__":
    main()
