#!/usr/bin/env python
"""Generic multi-machine process launcher for the data-preparation tools (parity: /root/reference/tools/distgraphlaunch.py):
starts ``--num_proc_per_machine`` copies of a command on every host of an ip_config file and gives each copy the
torch.distributed rendezvous variables (MASTER_ADDR / MASTER_PORT / WORLD_SIZE / RANK / LOCAL_RANK) so that the copies can
form one gloo group - this is how ``tools/dispatch_data.py`` runs the distributed partitioning pipeline.

  python tools/distgraphlaunch.py --workspace /data/ws --num_proc_per_machine 2 --ip_config ip_config.txt \
      "python tools/distpartitioning/data_proc_pipeline.py --world-size 4 ..."

Process management (own session + pid file per copy, termination on failure / signal, detached watchdog for the case that
the launcher itself dies) is shared with ``tools/launch.py``.
"""
import argparse
import os
import shlex
import signal
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from launch import is_local, kill_recorded_groups, read_hosts, start_watchdog  # noqa: E402


def build_commands(args, udf_command, job_tag):
    """[(host, argv, pidfile)] - pure function, unit-tested."""
    hosts = read_hosts(args.ip_config)
    per, world = args.num_proc_per_machine, args.num_proc_per_machine * len(hosts)
    master = hosts[0][0]
    out = []
    for node, (ip, _) in enumerate(hosts):
        for local in range(per):
            rank = node * per + local
            env = {"MASTER_ADDR": master, "MASTER_PORT": args.master_port, "WORLD_SIZE": world, "RANK": rank,
                   "LOCAL_RANK": local, "DGL_IP_CONFIG": args.ip_config, "OMP_NUM_THREADS": args.num_omp_threads or 1}
            env.update(dict(kv.split("=", 1) for kv in (args.extra_envs or [])))
            exports = " ".join(f"{k}={shlex.quote(str(v))}" for k, v in env.items())
            pidfile = f"/tmp/dglb200_distgraphlaunch_{job_tag}_{rank}.pid"
            body = f"cd {shlex.quote(args.workspace)} && {exports} {udf_command}"
            wrapped = f"exec setsid -w bash -c {shlex.quote('echo $$ > ' + shlex.quote(pidfile) + '; ' + body)}"
            if is_local(ip):
                argv = ["bash", "-c", wrapped]
            else:
                target = f"{args.ssh_username}@{ip}" if args.ssh_username else ip
                argv = ["ssh", "-o", "StrictHostKeyChecking=no", "-p", str(args.ssh_port), target, wrapped]
            out.append((ip, argv, pidfile))
    return out


def main():
    ap = argparse.ArgumentParser(description="Launch the same command on every machine of a cluster")
    ap.add_argument("--ssh_port", type=int, default=22)
    ap.add_argument("--ssh_username", default="")
    ap.add_argument("--workspace", type=str, required=True)
    ap.add_argument("--num_proc_per_machine", type=int, required=True)
    ap.add_argument("--master_port", type=int, default=1100)
    ap.add_argument("--ip_config", type=str, required=True)
    ap.add_argument("--num_omp_threads", type=int)
    ap.add_argument("--extra_envs", nargs="+", type=str, default=[])
    args, udf = ap.parse_known_args()
    if args.num_proc_per_machine <= 0:
        ap.error("--num_proc_per_machine must be a positive number.")
    if len(udf) != 1:
        ap.error("exactly one quoted command is expected")
    cmds = build_commands(args, udf[0], f"{os.getpid()}_{int(time.time())}")
    targets = [(ip, pf) for ip, _, pf in cmds]
    watchdog = start_watchdog(args.ssh_port, targets)
    procs = [subprocess.Popen(argv, start_new_session=True) for _, argv, _ in cmds]
    state = {"stop": False}

    def cleanup(*_):
        state["stop"] = True
        kill_recorded_groups(targets, args.ssh_port)
        for p in procs:
            try:
                os.killpg(os.getpgid(p.pid), signal.SIGTERM)
            except Exception:  # noqa: BLE001
                pass

    signal.signal(signal.SIGINT, cleanup)
    signal.signal(signal.SIGTERM, cleanup)
    rc = 0
    while not state["stop"]:
        codes = [p.poll() for p in procs]
        failed = [c for c in codes if c not in (None, 0)]
        if failed:
            rc = failed[0]
            cleanup()
        elif all(c == 0 for c in codes):
            break
        time.sleep(0.3)
    kill_recorded_groups(targets, args.ssh_port)
    try:
        watchdog.terminate()
    except Exception:  # noqa: BLE001
        pass
    sys.exit(rc)


if __name__ == "__main__":
    main()
