"""Container-aware thread budget (utils/cpu_quota.py)."""
import os

import torch

from dgl_b200.utils import cpu_quota


def test_quota_parsing(monkeypatch):
    files = {"/sys/fs/cgroup/cpu.max": "1600000 100000"}
    monkeypatch.setattr(cpu_quota, "_read", lambda p: files.get(p))
    assert cpu_quota.cgroup_cpu_quota() == 16.0
    monkeypatch.setattr(os, "sched_getaffinity", lambda pid: set(range(128)), raising=False)
    assert cpu_quota.effective_cpu_count() == 16
    files["/sys/fs/cgroup/cpu.max"] = "max 100000"
    assert cpu_quota.cgroup_cpu_quota() is None and cpu_quota.effective_cpu_count() == 128
    files.clear()
    files.update({"/sys/fs/cgroup/cpu/cpu.cfs_quota_us": "250000", "/sys/fs/cgroup/cpu/cpu.cfs_period_us": "100000"})
    assert cpu_quota.cgroup_cpu_quota() == 2.5 and cpu_quota.effective_cpu_count() == 3
    files["/sys/fs/cgroup/cpu/cpu.cfs_quota_us"] = "-1"
    assert cpu_quota.cgroup_cpu_quota() is None


def test_apply_caps_torch_and_native_threads(monkeypatch):
    from dgl_b200._native import loader
    import ctypes
    lib = loader.cpu_lib()
    get = lib.raw("dglb200_cpu_num_threads", ctypes.c_int32, [])
    before_t, before_n = torch.get_num_threads(), get()
    try:
        monkeypatch.delenv("OMP_NUM_THREADS", raising=False)
        monkeypatch.setenv("DGLB200_CPU_QUOTA", "1")
        monkeypatch.setattr(cpu_quota, "effective_cpu_count", lambda: 2)
        assert cpu_quota.apply(force=True) == 2 and torch.get_num_threads() <= 2
        cpu_quota.cap_native(lib)
        assert get() <= 2
        monkeypatch.setenv("DGLB200_CPU_QUOTA", "0")
        assert cpu_quota.apply(force=True) is None
    finally:
        torch.set_num_threads(before_t)
        lib.fn("dglb200_cpu_set_num_threads", "i")(before_n)
        cpu_quota._applied = None
