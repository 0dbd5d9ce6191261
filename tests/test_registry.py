"""Global function registry (dgl_b200/_native/registry.py) and the namespace re-exports added next to it."""
import ctypes

import pytest

import dgl_b200 as dgl
from dgl_b200.base import DGLError


def test_native_entry_points_are_listed_and_resolvable():
    names = dgl.list_global_func_names()
    assert len(names) > 50 and all(n.startswith("dglb200_") or "." in n or n.isidentifier() for n in names)
    assert "dglb200_last_error" in names
    f = dgl.get_global_func("dglb200_last_error")
    f.restype = ctypes.c_char_p
    assert f() in (None, b"") or isinstance(f(), bytes)


def test_register_lookup_override_and_missing():
    @dgl.register_func("tests.registry.inc")
    def inc(x):
        return x + 1

    assert dgl.get_global_func("tests.registry.inc")(2) == 3
    assert "tests.registry.inc" in dgl.list_global_func_names()
    with pytest.raises(DGLError):
        dgl.register_func("tests.registry.inc", lambda x: x)
    dgl.register_func("tests.registry.inc", lambda x: x + 2, override=True)
    assert dgl.get_global_func("tests.registry.inc")(2) == 4
    assert dgl.get_global_func("tests.registry.absent", allow_missing=True) is None
    with pytest.raises(DGLError):
        dgl.get_global_func("tests.registry.absent")


def test_extract_ext_funcs_collects_callback_pairs():
    def finit(cb):
        cb("a", len)
        cb("b", abs)
        return 0

    assert dgl.extract_ext_funcs(finit) == {"a": len, "b": abs}


def test_reexported_names_resolve():
    import dgl_b200.distributed as dist
    import dgl_b200.graphbolt as gb
    import dgl_b200.graphbolt.impl as impl
    assert dist.DEFAULT_ETYPE == ("_N", "_E", "_N") and callable(dist.fast_pull) and callable(dist.read_ip_config)
    assert callable(gb.traverse_dps) and callable(gb.replace_dp)
    assert impl.OnDiskMetaData is gb.OnDiskMetaData and impl.ItemSet is gb.ItemSet
    assert dgl.TypeCode.STR == 11 and callable(dgl.apply_each) and callable(dgl.set_edge_lazy_features)
    with pytest.raises(AttributeError):
        impl.no_such_name
