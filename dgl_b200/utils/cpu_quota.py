"""CPU budget of the process: the smaller of its scheduler affinity and its cgroup CPU quota.

Why it exists: GPU boxes hand a container a *quota* (e.g. 16 CPUs' worth of time on a 128-core host, cgroup
``cpu.max`` = "1600000 100000") while ``nproc`` / ``os.cpu_count()`` still report every core. OpenMP and torch then size
their thread teams by the core count, the team burns the quota in the first fraction of every 100 ms scheduling period
and the whole container is throttled for the rest of it. Measured on the B200 box (128 visible cores, quota 16): one
mini-batch of the host feature path took 90.6 ms with default teams (64 / 128 threads) and 36.9 ms with 8 threads
(``benchmarks/profile_fetch_stages.py``, profiles/r2_fetch_stage_threads.txt).

``apply()`` - called once when the package is imported - caps torch's intra-op threads and the native library's OpenMP
teams to that budget unless the user already chose (``OMP_NUM_THREADS`` set, or ``DGLB200_CPU_QUOTA=0`` to opt out).
The reference has no equivalent (it inherits OpenMP's default); the behaviour mirrors what container-aware runtimes
(JVM, Go) do."""
from __future__ import annotations

import math
import os


def _read(path):
    try:
        with open(path) as f:
            return f.read().strip()
    except OSError:
        return None


def cgroup_cpu_quota():
    """CPUs' worth of time the cgroup allows (float), or None when unlimited / unknown. cgroup v2 then v1."""
    v2 = _read("/sys/fs/cgroup/cpu.max")
    if v2:
        parts = v2.split()
        if len(parts) == 2 and parts[0] != "max":
            try:
                return float(parts[0]) / float(parts[1])
            except (ValueError, ZeroDivisionError):
                return None
        return None
    quota = _read("/sys/fs/cgroup/cpu/cpu.cfs_quota_us") or _read("/sys/fs/cgroup/cpu,cpuacct/cpu.cfs_quota_us")
    period = _read("/sys/fs/cgroup/cpu/cpu.cfs_period_us") or _read("/sys/fs/cgroup/cpu,cpuacct/cpu.cfs_period_us")
    try:
        if quota is not None and period is not None and int(quota) > 0 and int(period) > 0:
            return int(quota) / int(period)
    except ValueError:
        pass
    return None


def effective_cpu_count():
    """Threads worth running: min(affinity, ceil(cgroup quota)), at least 1."""
    try:
        n = len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        n = os.cpu_count() or 1
    q = cgroup_cpu_quota()
    if q is not None:
        n = min(n, max(1, math.ceil(q)))
    return max(1, n)


_applied = None


def apply(force=False):
    """Cap torch and native-library thread teams to the CPU budget. Returns the budget (or None when opted out)."""
    global _applied
    if _applied is not None and not force:
        return _applied
    if os.environ.get("DGLB200_CPU_QUOTA", "1") == "0" or os.environ.get("OMP_NUM_THREADS"):
        _applied = 0
        return None
    budget = effective_cpu_count()
    import torch
    if torch.get_num_threads() > budget:
        torch.set_num_threads(budget)
    _applied = budget
    return budget


def cap_native(lib):
    """Called by the loader right after the host library is loaded: size its OpenMP teams to the same budget."""
    if os.environ.get("DGLB200_CPU_QUOTA", "1") == "0" or os.environ.get("OMP_NUM_THREADS"):
        return
    budget = effective_cpu_count()
    try:
        get = lib.raw("dglb200_cpu_num_threads", __import__("ctypes").c_int32, [])
        if get() > budget:
            lib.fn("dglb200_cpu_set_num_threads", "i")(budget)
    except Exception:  # noqa: BLE001 - sizing is an optimisation, never a reason to fail loading
        pass
