"""Row-sharded full-graph GraphSAGE training on the GPUs of one box (BASELINE.json config 5).

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 sharded_training.py \
        --dataset products --maxk 32 --epochs 20

Every rank holds its row slab of the synthetic features / labels and a replica of the weights; each layer
all_gathers the compact CBSR slab (5k bytes per node) before its local SpGEMM and reduce_scatters the
sampled gradient in backward (sharded.py); weight gradients are summed with one all_reduce per epoch.
Timing follows all_train.py:118-149 (CUDA events around forward and backward after warm-up epochs),
reported as the max over ranks.
"""
import argparse
import json
import os

import torch
import torch.distributed as dist
import torch.nn.functional as F

from sharded import ShardedMaxKSAGE, allreduce_gradients
from synth_graphs import SHAPES, symmetrize, synth_graph


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dataset", default="products", choices=sorted(SHAPES))
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--maxk", type=int, default=32)
    ap.add_argument("--hidden_dim", type=int, default=256)
    ap.add_argument("--hidden_layers", type=int, default=3)
    ap.add_argument("--in_size", type=int, default=100)
    ap.add_argument("--classes", type=int, default=47)
    ap.add_argument("--epochs", type=int, default=20)
    ap.add_argument("--warmup_epochs", type=int, default=5)
    ap.add_argument("--bwd-mode", default="reduce_scatter")
    ap.add_argument("--partition", default="nnz", choices=["rows", "nnz"],
                    help="row partition: equal row slabs or equal edge counts (the graph here has power-law degrees)")
    a = ap.parse_args()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    world, rank = dist.get_world_size(), dist.get_rank()
    n, e = SHAPES[a.dataset]
    n, e = max(64, int(n * a.scale)), max(64, int(e * a.scale))
    graph = symmetrize(synth_graph(n, max(e // 2, 1), seed=123, kind="powerlaw", device=dev))   # same graph on every rank
    n = graph["v_num"]
    torch.manual_seed(0)                                                # identical initial weights on every rank
    model = ShardedMaxKSAGE(graph, a.in_size, a.hidden_dim, a.hidden_layers, a.classes, maxk=a.maxk,
                            backward_mode=a.bwd_mode, partition=a.partition).to(dev)
    m = model.agg.m                                                     # padded slab height (same on every rank)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    x = torch.randn(m, a.in_size, device=dev, generator=gen)
    y = torch.randint(0, a.classes, (m,), device=dev, generator=gen)
    valid = torch.arange(m, device=dev) < model.agg.valid_rows()
    del graph
    opt = torch.optim.Adam(model.parameters(), lr=0.01)
    fwd_ms = bwd_ms = 0.0
    timed = 0
    loss = None
    for epoch in range(a.epochs):
        timing = epoch >= a.warmup_epochs
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record()
        logits = model(x)
        loss = F.cross_entropy(logits[valid], y[valid], reduction="sum") / n
        ev[1].record()
        opt.zero_grad()
        ev[2].record()
        loss.backward()
        allreduce_gradients(model)
        ev[3].record()
        opt.step()
        if timing:
            torch.cuda.synchronize()
            t = torch.tensor([ev[0].elapsed_time(ev[1]), ev[2].elapsed_time(ev[3])], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            fwd_ms += float(t[0])
            bwd_ms += float(t[1])
            timed += 1
    total = loss.detach().clone()
    dist.all_reduce(total)
    if rank == 0:
        print(json.dumps({"model": "ShardedMaxKSAGE", "dataset": a.dataset, "nodes": n, "world": world, "maxk": a.maxk,
                          "hidden_layers": a.hidden_layers, "avg_forward_ms": fwd_ms / max(timed, 1),
                          "avg_backward_ms": bwd_ms / max(timed, 1), "final_loss": float(total)}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
