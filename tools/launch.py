#!/usr/bin/env python
"""Cluster launcher (parity: /root/reference/tools/launch.py:18-33,504-655): starts ``num_servers`` graph-server
processes and ``num_trainers`` trainer processes (torch.distributed.run) on every host of an ip_config file, over
ssh (or locally for 127.0.0.1), exporting the DGL_* role variables. Every launched command runs in its own session and
records its process-group id in a pid file; on SIGINT / SIGTERM / a failed child the launcher terminates those groups, and
a detached WATCHDOG process does the same if the launcher itself dies (kill -9, closed terminal): it polls the launcher's
pid and then signals exactly the recorded groups - locally with ``killpg``, remotely with ``ssh host kill -- -<pgid>``
(the reference's cleanup process greps remote ``ps`` output for the command line instead, tools/launch.py:18-33,626-655).

  python tools/launch.py --workspace /path --num_trainers 8 --num_servers 1 --part_config data/graph.json \
      --ip_config ip_config.txt "python train_dist.py --arg ..."
"""
import argparse
import os
import shlex
import signal
import subprocess
import sys
import threading
import time


def read_hosts(ip_config):
    hosts = []
    with open(ip_config) as f:
        for line in f:
            p = line.split()
            if p:
                hosts.append((p[0], int(p[1]) if len(p) > 1 else 30050))
    return hosts


def is_local(ip):
    return ip in ("127.0.0.1", "localhost", "0.0.0.0")


def wrap_cmd(ip, ssh_port, cmd, env, pidfile=None):
    exports = " ".join(f"{k}={shlex.quote(str(v))}" for k, v in env.items())
    full = f"cd {shlex.quote(env['DGL_WORKSPACE'])} && {exports} {cmd}"
    if pidfile:
        # own session on the target host; the group id (== pid of the session leader) is what cleanup signals
        full = f"exec setsid -w bash -c {shlex.quote('echo $$ > ' + shlex.quote(pidfile) + '; ' + full)}"
    if is_local(ip):
        return ["bash", "-c", full]
    return ["ssh", "-o", "StrictHostKeyChecking=no", "-p", str(ssh_port), ip, full]


def kill_recorded_groups(targets, ssh_port, sig="TERM"):
    """targets: [(ip, pidfile)]. Signals the process group recorded in each pid file, on the host that wrote it."""
    for ip, pidfile in targets:
        snippet = f"test -f {shlex.quote(pidfile)} && kill -{sig} -- -$(cat {shlex.quote(pidfile)}) ; rm -f {shlex.quote(pidfile)}"
        argv = ["bash", "-c", snippet] if is_local(ip) else \
            ["ssh", "-o", "StrictHostKeyChecking=no", "-o", "ConnectTimeout=5", "-p", str(ssh_port), ip, snippet]
        try:
            subprocess.run(argv, timeout=20, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001 - best effort: an unreachable host cannot be cleaned
            pass


def watchdog_main(parent_pid, ssh_port, targets, poll_s=1.0):
    """Body of the detached cleanup process: wait until the launcher is gone, then clean every recorded group."""
    while True:
        try:
            os.kill(parent_pid, 0)
        except ProcessLookupError:
            break
        except PermissionError:
            pass
        time.sleep(poll_s)
    kill_recorded_groups(targets, ssh_port)


def start_watchdog(ssh_port, targets):
    import json
    payload = json.dumps({"parent": os.getpid(), "ssh_port": ssh_port, "targets": targets})
    return subprocess.Popen([sys.executable, os.path.abspath(__file__), "--_watchdog", payload], start_new_session=True,
                            stdin=subprocess.DEVNULL, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def build_commands(args, udf_command, job_tag=None):
    """Pure function (unit-tested): the list of (host, argv) to start; with ``job_tag`` every command records its process
    group in ``/tmp/dglb200_launch_<tag>_<k>.pid`` and the list entries become (host, argv, pidfile)."""
    hosts = read_hosts(args.ip_config)
    cmds = []
    counter = [0]

    def pidfile():
        if job_tag is None:
            return None
        counter[0] += 1
        return f"/tmp/dglb200_launch_{job_tag}_{counter[0]}.pid"
    num_clients = args.num_trainers * (1 + args.num_samplers) * len(hosts)
    base_env = {"DGL_WORKSPACE": args.workspace, "DGL_DIST_MODE": "distributed", "DGL_IP_CONFIG": args.ip_config,
                "DGL_CONF_PATH": args.part_config, "DGL_NUM_SERVER": args.num_servers, "DGL_NUM_CLIENT": num_clients,
                "DGL_NUM_SAMPLER": args.num_samplers, "OMP_NUM_THREADS": args.num_omp_threads or 1,
                "DGL_GRAPH_FORMAT": args.graph_format}
    base_env.update(dict(kv.split("=", 1) for kv in (args.extra_envs or [])))
    sid = 0
    for ip, _ in hosts:
        for _ in range(args.num_servers):
            env = dict(base_env, DGL_ROLE="server", DGL_SERVER_ID=sid)
            pf = pidfile()
            cmds.append((ip, wrap_cmd(ip, args.ssh_port, udf_command, env, pf)) + ((pf,) if pf else ()))
            sid += 1
    master_ip = hosts[0][0]
    for node_rank, (ip, _) in enumerate(hosts):
        env = dict(base_env, DGL_ROLE="client")
        torch_cmd = (f"{sys.executable} -m torch.distributed.run --nproc_per_node={args.num_trainers} "
                     f"--nnodes={len(hosts)} --node_rank={node_rank} --master_addr={master_ip} "
                     f"--master_port={args.master_port} ")
        # torch.distributed.run supplies the interpreter: drop a leading "python" / "/path/to/python3.x" token
        first, _, rest = udf_command.partition(" ")
        trainer = rest if os.path.basename(first).startswith("python") else udf_command
        pf = pidfile()
        cmds.append((ip, wrap_cmd(ip, args.ssh_port, torch_cmd + trainer, env, pf)) + ((pf,) if pf else ()))
    return cmds


def main():
    if len(sys.argv) == 3 and sys.argv[1] == "--_watchdog":
        import json
        cfg = json.loads(sys.argv[2])
        watchdog_main(cfg["parent"], cfg["ssh_port"], [tuple(t) for t in cfg["targets"]])
        return
    ap = argparse.ArgumentParser(description="Launch a distributed job")
    ap.add_argument("--ssh_port", type=int, default=22)
    ap.add_argument("--workspace", type=str, required=True)
    ap.add_argument("--num_trainers", type=int, required=True)
    ap.add_argument("--num_omp_threads", type=int)
    ap.add_argument("--num_samplers", type=int, default=0)
    ap.add_argument("--num_servers", type=int, default=1)
    ap.add_argument("--part_config", type=str, required=True)
    ap.add_argument("--ip_config", type=str, required=True)
    ap.add_argument("--graph_format", type=str, default="csc")
    ap.add_argument("--master_port", type=int, default=1234)
    ap.add_argument("--extra_envs", nargs="+", type=str, default=[])
    args, udf = ap.parse_known_args()
    if len(udf) != 1:
        ap.error("exactly one quoted training command is expected")
    tag = f"{os.getpid()}_{int(time.time())}"
    cmds = build_commands(args, udf[0], job_tag=tag)
    targets = [(ip, pf) for ip, _, pf in cmds]
    procs = []
    stop = threading.Event()
    watchdog = start_watchdog(args.ssh_port, targets)

    def cleanup(*_):
        stop.set()
        kill_recorded_groups(targets, args.ssh_port)       # the recorded groups on every host (remote ones through ssh)
        for p in procs:                                    # and the local ssh / bash clients themselves
            try:
                os.killpg(os.getpgid(p.pid), signal.SIGTERM)
            except Exception:  # noqa: BLE001
                pass

    signal.signal(signal.SIGINT, cleanup)
    signal.signal(signal.SIGTERM, cleanup)
    for ip, argv, _ in cmds:
        procs.append(subprocess.Popen(argv, start_new_session=True))
    rc = 0
    while not stop.is_set():
        alive = False
        for p in procs:
            r = p.poll()
            if r is None:
                alive = True
            elif r != 0 and rc == 0:
                rc = r                      # the FIRST failure is the job's status; the rest die from the cleanup below
                cleanup()
        if not alive:
            break
        time.sleep(0.5)
    kill_recorded_groups(targets, args.ssh_port)            # normal exit: nothing should be left, remove the pid files
    try:
        watchdog.terminate()
    except Exception:  # noqa: BLE001
        pass
    sys.exit(rc)


if __name__ == "__main__":
    main()
