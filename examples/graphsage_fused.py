"""The flagship path: whole-step-in-one-CUDA-graph GraphSAGE training (dgl_b200.models.FusedSAGETrainer).
Single GPU:  python examples/graphsage_fused.py
Multi GPU :  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 examples/graphsage_fused.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dgl_b200 as dgl
from dgl_b200.data.synthetic import shaped_dataset
from dgl_b200.models import FusedSAGETrainer, GraphSAGE


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    comm = None
    if world > 1:
        import torch.distributed as dist
        from dgl_b200.parallel import P2PComm
        dist.init_process_group("nccl", device_id=dev)
        comm = P2PComm()
    g, feat, labels, train_idx, C = shaped_dataset("ogbn-arxiv", device=dev)
    with torch.no_grad():  # learnable synthetic labels
        labels = (dgl.ops.copy_u_mean(g, feat) @ torch.randn(feat.shape[1], C, device=dev)).argmax(1)
    tr = FusedSAGETrainer(g, feat, labels, C, hidden=256, fanouts=[10, 10, 10], batch_size=1024, lr=3e-3, dropout=0.2,
                          comm=comm)
    mine = train_idx.cpu()[rank::world]
    for epoch in range(3):
        perm = mine[torch.randperm(mine.numel())]
        batches = [perm[i:i + 1024] for i in range(0, perm.numel() - 1023, 1024)]
        losses = tr.fit(batches)
        if rank == 0:
            print(f"epoch {epoch}: {len(losses)} steps, loss {losses[0]:.3f} -> {losses[-1]:.3f}")
    mod = GraphSAGE(feat.shape[1], 256, C, 3, 0.0).to(dev)
    tr.export_to_module(mod)  # continue with the generic dataloader / inference APIs


if __name__ == "__main__":
    main()
