#!/usr/bin/env python
"""Cluster launcher (parity: /root/reference/tools/launch.py:18-33,504-655): starts ``num_servers`` graph-server
processes and ``num_trainers`` trainer processes (torch.distributed.run) on every host of an ip_config file, over
ssh (or locally for 127.0.0.1), exporting the DGL_* role variables; a cleanup thread kills every child on SIGINT /
failure.

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


def wrap_cmd(ip, ssh_port, cmd, env):
    exports = " ".join(f"{k}={shlex.quote(str(v))}" for k, v in env.items())
    full = f"cd {shlex.quote(env['DGL_WORKSPACE'])} && {exports} {cmd}"
    if is_local(ip):
        return ["bash", "-c", full]
    return ["ssh", "-o", "StrictHostKeyChecking=no", "-p", str(ssh_port), ip, full]


def build_commands(args, udf_command):
    """Pure function (unit-tested): the list of (host, argv) to start."""
    hosts = read_hosts(args.ip_config)
    cmds = []
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
            cmds.append((ip, wrap_cmd(ip, args.ssh_port, udf_command, env)))
            sid += 1
    master_ip = hosts[0][0]
    for node_rank, (ip, _) in enumerate(hosts):
        env = dict(base_env, DGL_ROLE="client")
        torch_cmd = (f"{sys.executable} -m torch.distributed.run --nproc_per_node={args.num_trainers} "
                     f"--nnodes={len(hosts)} --node_rank={node_rank} --master_addr={master_ip} "
                     f"--master_port={args.master_port} ")
        trainer = udf_command
        if trainer.startswith("python"):
            trainer = trainer.split(" ", 1)[1] if " " in trainer else ""
        cmds.append((ip, wrap_cmd(ip, args.ssh_port, torch_cmd + trainer, env)))
    return cmds


def main():
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
    cmds = build_commands(args, udf[0])
    procs = []
    stop = threading.Event()

    def cleanup(*_):
        stop.set()
        for p in procs:
            try:
                os.killpg(os.getpgid(p.pid), signal.SIGTERM)
            except Exception:  # noqa: BLE001
                pass

    signal.signal(signal.SIGINT, cleanup)
    signal.signal(signal.SIGTERM, cleanup)
    for ip, argv in cmds:
        procs.append(subprocess.Popen(argv, start_new_session=True))
    rc = 0
    while not stop.is_set():
        alive = False
        for p in procs:
            r = p.poll()
            if r is None:
                alive = True
            elif r != 0:
                rc = r
                cleanup()
        if not alive:
            break
        time.sleep(0.5)
    sys.exit(rc)


if __name__ == "__main__":
    main()
