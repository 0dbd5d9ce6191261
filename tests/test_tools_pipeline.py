"""chunk_graph -> random_partition -> dispatch_data -> verify_partitions (reference tools/ pipeline) on a toy graph."""
import os
import sys

import torch

import dgl_b200 as dgl

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "tools", "partition_algo"))


def test_chunk_partition_dispatch_verify(tmp_path):
    from chunk_graph import chunk_graph, load_chunked
    from dispatch_data import dispatch
    from random_partition import random_partition
    from verify_partitions import verify
    g = dgl.rand_graph(500, 4000, idtype=torch.int64)
    g.ndata["feat"] = torch.randn(500, 6)
    g.ndata["label"] = torch.randint(0, 4, (500,))
    g.edata["w"] = torch.rand(4000)
    chunked = str(tmp_path / "chunked")
    meta = chunk_graph(g, "toy", num_chunks=3, output_path=chunked)
    assert meta["num_nodes_per_chunk"] == [[167, 167, 166]] and len(meta["edges"]["_N:_E:_N"]["data"]) == 3
    g2, _ = load_chunked(chunked)
    assert torch.equal(g2.ndata["feat"], g.ndata["feat"]) and torch.equal(g2.edges()[0], g.edges()[0])
    parts = str(tmp_path / "parts")
    random_partition(chunked, 3, parts)
    cfg = dispatch(chunked, parts, str(tmp_path / "out"))
    assert verify(chunked, cfg)
    # the result is consumable by the distributed runtime
    from dgl_b200 import distributed as dd
    os.environ["DGL_DIST_MODE"] = "standalone"
    dd.initialize(part_config=cfg)
    dg = dd.DistGraph("toy")
    assert dg.num_nodes() == 500 and dg.num_edges() == 4000
    dd.exit_client()


def _dist_worker(rank, world, port, chunked, parts, out, q):
    import traceback
    try:
        os.environ.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port), "RANK": str(rank), "WORLD_SIZE": str(world)})
        sys.path.insert(0, ROOT)
        import torch.distributed as dist
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from tools.distpartitioning.data_shuffle import gen_dist_partitions
        path = gen_dist_partitions(rank, world, chunked, parts, out)
        dist.destroy_process_group()
        q.put((rank, "ok", path))
    except Exception:  # noqa: BLE001
        q.put((rank, "error", traceback.format_exc()))


def test_distributed_partition_pipeline_gloo(tmp_path):
    """One gloo process per partition; the result must equal partition_graph(node_part=...) run in one process."""
    import socket
    import torch.multiprocessing as mp
    from chunk_graph import chunk_graph
    from random_partition import random_partition
    from dgl_b200 import distributed as dd
    g = dgl.rand_graph(400, 3000, idtype=torch.int64)
    g.ndata["feat"] = torch.randn(400, 5)
    g.edata["w"] = torch.rand(3000)
    chunked, parts, out = str(tmp_path / "chunked"), str(tmp_path / "parts"), str(tmp_path / "out")
    chunk_graph(g, "toy", num_chunks=3, output_path=chunked)
    random_partition(chunked, 2, parts)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_dist_worker, args=(r, 2, port, chunked, parts, out, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(2)]
    for p in procs:
        p.join(30)
    assert all(st == "ok" for _, st, _ in res), [m for _, st, m in res if st != "ok"]
    cfg = os.path.join(out, "toy.json")
    import numpy as np
    node_part = torch.from_numpy(np.loadtxt(os.path.join(parts, "_N.txt"), dtype=np.int64))
    ref_dir = str(tmp_path / "ref")
    dd.partition_graph(g, "toy", 2, ref_dir, node_part=node_part)
    u, v = g.edges()
    for p in range(2):
        pg, nf, ef, gpb, *_ = dd.load_partition(cfg, p)
        rg, rnf, ref_, rgpb, *_ = dd.load_partition(os.path.join(ref_dir, "toy.json"), p)
        assert gpb._node_map["_N"].tolist() == rgpb._node_map["_N"].tolist()
        assert gpb._edge_map["_N:_E:_N"].tolist() == rgpb._edge_map["_N:_E:_N"].tolist()
        inner = pg.ndata["inner_node"].bool()
        orig = pg.ndata["orig_id"]
        assert torch.equal(torch.sort(orig[inner])[0], torch.nonzero(node_part == p).squeeze(1))
        assert torch.equal(nf["_N/feat"], g.ndata["feat"][orig[inner]])                 # features follow the shuffle
        ps, pd = pg.edges()
        eo = pg.edata["orig_id"]
        assert torch.equal(orig[ps.long()], u[eo]) and torch.equal(orig[pd.long()], v[eo])   # edges map back exactly
        assert torch.equal(torch.sort(eo)[0], torch.nonzero(node_part[v] == p).squeeze(1))
        assert torch.allclose(ef["_N:_E:_N/w"], g.edata["w"][eo])
        # same inner node numbering as the single-process partitioner
        assert torch.equal(pg.ndata["orig_id"][inner], rg.ndata["orig_id"][rg.ndata["inner_node"].bool()])
    # consumable by the runtime
    os.environ["DGL_DIST_MODE"] = "standalone"
    dd.initialize(part_config=cfg)
    dg = dd.DistGraph("toy")
    assert dg.num_nodes() == 400 and dg.num_edges() == 3000
    dd.exit_client()


def test_parmetis_pre_post_process(tmp_path):
    import numpy as np
    from chunk_graph import chunk_graph
    sys.path.insert(0, ROOT)
    from tools.distpartitioning.parmetis_postprocess import post_process
    from tools.distpartitioning.parmetis_preprocess import gen_parmetis_input
    hg = dgl.heterograph({("a", "r", "b"): (torch.randint(0, 30, (100,)), torch.randint(0, 20, (100,))),
                          ("b", "q", "a"): (torch.randint(0, 20, (80,)), torch.randint(0, 30, (80,)))})
    chunked = str(tmp_path / "c")
    chunk_graph(hg, "h", num_chunks=2, output_path=chunked)
    nfiles, efiles = gen_parmetis_input(chunked, str(tmp_path / "pm"))
    n = sum(np.loadtxt(f, dtype=np.int64).reshape(-1, 4).shape[0] for f in nfiles)
    e = sum(np.loadtxt(f, dtype=np.int64).reshape(-1, 2).shape[0] for f in efiles)
    assert n == 50 and e == 180
    out = str(tmp_path / "pm" / "h_part.2")
    np.savetxt(out, np.stack([np.arange(50), np.arange(50) % 2], 1), fmt="%d")
    pdir = post_process(chunked, out, str(tmp_path / "parts"))
    assert np.loadtxt(os.path.join(pdir, "a.txt")).shape[0] == 30 and np.loadtxt(os.path.join(pdir, "b.txt")).shape[0] == 20


def test_launcher_watchdog_cleans_recorded_groups_when_launcher_dies(tmp_path):
    """tools/launch.py: a job whose launcher is killed with SIGKILL must not leave its children behind - the detached
    watchdog signals exactly the process groups recorded in the pid files."""
    import signal
    import subprocess
    import sys
    import time
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ipc = tmp_path / "ip_config.txt"
    ipc.write_text("127.0.0.1 30050\n")
    marker = tmp_path / "alive.txt"
    # the "training command": a server-role process that keeps touching a file; trainers exit immediately
    script = tmp_path / "job.py"
    script.write_text(
        "import os, sys, time\n"
        "if os.environ.get('DGL_ROLE') != 'server':\n"
        "    time.sleep(600)\n"
        f"p = {str(marker)!r}\n"
        "while True:\n"
        "    open(p, 'w').write(str(time.time()))\n"
        "    time.sleep(0.1)\n")
    launcher = subprocess.Popen([sys.executable, os.path.join(root, "tools", "launch.py"), "--workspace", str(tmp_path),
                                 "--num_trainers", "1", "--num_servers", "1", "--part_config", "none.json", "--ip_config", str(ipc),
                                 f"{sys.executable} {script}"], start_new_session=True)
    try:
        deadline = time.time() + 30
        while not marker.exists() and time.time() < deadline:
            time.sleep(0.1)
        assert marker.exists(), "the server process never started"
        os.kill(launcher.pid, signal.SIGKILL)            # no chance to run its own cleanup
        launcher.wait()
        # the watchdog notices within ~1 s and terminates the recorded groups: the file stops changing
        time.sleep(4.0)
        a = marker.read_text()
        time.sleep(1.0)
        assert marker.read_text() == a, "a child survived the launcher"
    finally:
        try:
            os.killpg(os.getpgid(launcher.pid), signal.SIGKILL)
        except Exception:  # noqa: BLE001
            pass


def test_cli_recipe_train_apply_roundtrip(tmp_path):
    """dgl_cli: recipes come from tools/recipes/*.yaml, the dataset named in the config is honoured, ``train`` writes a
    checkpoint and ``apply`` reproduces the trained model's predictions from it."""
    import subprocess
    import sys
    import torch
    import yaml
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cli = [sys.executable, os.path.join(root, "tools", "dgl_cli.py")]
    names = subprocess.run(cli + ["recipe", "list"], capture_output=True, text=True, check=True).stdout.split()
    assert "nodepred_citeseer_gcn" in names and "linkpred_cora_sage" in names and len(names) >= 12
    cfg_path = tmp_path / "c.yaml"
    subprocess.run(cli + ["recipe", "get", "nodepred_citeseer_gcn", "--cfg", str(cfg_path)], check=True, capture_output=True)
    cfg = yaml.safe_load(open(cfg_path))
    assert cfg["data"]["name"] == "citeseer" and cfg["model"] == "gcn"
    cfg["general_pipeline"].update(num_epochs=5, save_path=str(tmp_path / "m.pt"))
    yaml.safe_dump(cfg, open(cfg_path, "w"))
    out = subprocess.run(cli + ["train", "--cfg", str(cfg_path)], capture_output=True, text=True, cwd=tmp_path)
    assert out.returncode == 0 and "checkpoint written" in out.stdout, out.stderr[-800:]
    out = subprocess.run(cli + ["apply", "--cfg", str(cfg_path), "--checkpoint", str(tmp_path / "m.pt"), "--output",
                                str(tmp_path / "p.pt")], capture_output=True, text=True, cwd=tmp_path)
    assert out.returncode == 0 and "predictions written" in out.stdout, out.stderr[-800:]
    pred = torch.load(tmp_path / "p.pt")
    import dgl_b200.data as D
    assert pred.shape[0] == D.CiteseerGraphDataset()[0].num_nodes()
    bad = subprocess.run(cli + ["recipe", "get", "nope"], capture_output=True, text=True)
    assert bad.returncode != 0


def test_distgraphlaunch_sets_rendezvous_env_and_propagates_failure(tmp_path):
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ipc = tmp_path / "ip_config.txt"
    ipc.write_text("127.0.0.1\n")
    script = tmp_path / "w.py"
    script.write_text("import os\n"
                      f"open(os.path.join({str(tmp_path)!r}, 'r' + os.environ['RANK']), 'w').write("
                      "','.join(os.environ[k] for k in ('WORLD_SIZE', 'LOCAL_RANK', 'MASTER_ADDR', 'MASTER_PORT', 'FOO')))\n")
    cmd = [sys.executable, os.path.join(root, "tools", "distgraphlaunch.py"), "--workspace", str(tmp_path), "--num_proc_per_machine", "3",
           "--extra_envs", "FOO=bar", "--ip_config", str(ipc), "--master_port", "12345", f"{sys.executable} {script}"]
    assert subprocess.run(cmd, timeout=60).returncode == 0
    for r in range(3):
        assert (tmp_path / f"r{r}").read_text() == f"3,{r},127.0.0.1,12345,bar"
    bad = tmp_path / "bad.py"
    bad.write_text("import os, sys, time\nif os.environ['RANK'] == '1':\n    sys.exit(7)\ntime.sleep(30)\n")
    cmd[-1] = f"{sys.executable} {bad}"
    t0 = __import__("time").time()
    assert subprocess.run(cmd, timeout=60).returncode == 7
    assert __import__("time").time() - t0 < 20          # the surviving copies were terminated, not waited for


def test_launcher_waits_for_the_job_and_returns_its_status(tmp_path):
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ipc = tmp_path / "ip_config.txt"
    ipc.write_text("127.0.0.1 30050\n")
    script = tmp_path / "job.py"
    script.write_text("import os, time\n"
                      "time.sleep(1.5)\n"
                      f"open(os.path.join({str(tmp_path)!r}, 'done_' + os.environ['DGL_ROLE']), 'w').write('x')\n")
    cmd = [sys.executable, os.path.join(root, "tools", "launch.py"), "--workspace", str(tmp_path), "--num_trainers", "1",
           "--num_servers", "1", "--part_config", "none.json", "--ip_config", str(ipc), f"{sys.executable} {script}"]
    out = subprocess.run(cmd, timeout=120)
    assert out.returncode == 0
    assert (tmp_path / "done_server").exists() and (tmp_path / "done_client").exists()
    script.write_text("import os, sys, time\nsys.exit(5 if os.environ['DGL_ROLE'] == 'server' else time.sleep(60))\n")
    assert subprocess.run(cmd, timeout=120).returncode == 5
