#!/usr/bin/env python
"""Copy partitions (and scripts) to every machine of a cluster (reference tools/copy_files.py): for machine i of
``ip_config`` copies ``part<i>/``, the partition config and the listed script folders into ``<workspace>/<rel_data_path>``
with scp; local ips are copied with shutil."""
import argparse
import json
import os
import shutil
import subprocess


def copy_file(src, ip, dst_dir, local=False, param=""):
    if local:
        os.makedirs(dst_dir, exist_ok=True)
        (shutil.copytree if os.path.isdir(src) else shutil.copy2)(src, os.path.join(dst_dir, os.path.basename(src)),
                                                                  **({"dirs_exist_ok": True} if os.path.isdir(src) else {}))
        return
    subprocess.check_call(f"ssh -o StrictHostKeyChecking=no {ip} 'mkdir -p {dst_dir}'", shell=True)
    subprocess.check_call(f"scp {param} {'-r' if os.path.isdir(src) else ''} {src} {ip}:{dst_dir}/", shell=True)


def main():
    ap = argparse.ArgumentParser(description="Copy data to the servers.")
    ap.add_argument("--workspace", required=True)
    ap.add_argument("--rel_data_path", required=True)
    ap.add_argument("--part_config", required=True)
    ap.add_argument("--script_folder", default=None)
    ap.add_argument("--ip_config", required=True)
    a = ap.parse_args()
    hosts = [line.split()[0] for line in open(a.ip_config) if line.strip()]
    with open(a.part_config) as f:
        meta = json.load(f)
    assert len(hosts) == meta["num_parts"], "The number of partitions needs to be the same as the number of hosts."
    base = os.path.dirname(os.path.abspath(a.part_config))
    for i, ip in enumerate(hosts):
        local = ip in ("127.0.0.1", "localhost")
        dst = os.path.join(a.workspace, a.rel_data_path)
        copy_file(a.part_config, ip, dst, local)
        copy_file(os.path.join(base, f"part{i}"), ip, dst, local)
        if a.script_folder:
            copy_file(a.script_folder, ip, a.workspace, local)
    print(f"copied {len(hosts)} partitions")


if __name__ == "__main__":
    main()
