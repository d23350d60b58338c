"""Full-graph single-GPU training loop on synthetic data.

The shape of the reference's trainer (all_train.py:93-208): Adam, cross-entropy (or BCE-with-logits
for multi-label sets, :97-100), per-epoch forward and backward CUDA-event timing after warm-up epochs
(:118-149) and the same final timing report fields -- on a synthetic graph of a named shape with
random features / labels (there is no network for the datasets), through the DGL-free models of
maxk_models_integrated.py.

    python maxk_gnn_training.py --dataset yelp --model gcn --maxk 32 --epochs 30 --scale 0.1
"""
import argparse
import json
import time

import torch
import torch.nn as nn
import torch.nn.functional as F

from maxk_models_integrated import CSRGraph, MaxKGCN, MaxKGIN, MaxKSAGE
from synth_graphs import SHAPES, symmetrize, synth_graph

MODELS = {"sage": MaxKSAGE, "gcn": MaxKGCN, "gin": MaxKGIN}
MULTI_LABEL = ("yelp", "proteins")           # all_train.py:97 uses BCE-with-logits for these


def synthetic_task(dataset, scale, in_size, classes, device, seed=0):
    n, e = SHAPES[dataset]
    n, e = max(64, int(n * scale)), max(64, int(e * scale))
    g = symmetrize(synth_graph(n, max(e // 2, 1), seed=123, kind="powerlaw", device=device))   # undirected + self loops
    gen = torch.Generator(device=device).manual_seed(seed)
    x = torch.randn(n, in_size, device=device, generator=gen)
    if dataset in MULTI_LABEL:
        y = (torch.rand(n, classes, device=device, generator=gen) < 0.3).float()
    else:
        y = torch.randint(0, classes, (n,), device=device, generator=gen)
    split = torch.rand(n, device=device, generator=gen)
    masks = (split < 0.6, (split >= 0.6) & (split < 0.8), split >= 0.8)
    return CSRGraph.from_dict(g, dataset), x, y, masks


def train(graph, features, labels, masks, model, epochs=30, lr=0.01, weight_decay=0.0, warmup_epochs=10,
          multi_label=False, log=print):
    loss_fcn = F.binary_cross_entropy_with_logits if multi_label else nn.CrossEntropyLoss()
    optimizer = torch.optim.Adam(model.parameters(), lr=lr, weight_decay=weight_decay)
    train_mask = masks[0]
    fwd_ms = bwd_ms = 0.0
    timed = 0
    losses = []
    for epoch in range(epochs):
        model.train()
        timing = epoch >= warmup_epochs
        if timing:
            torch.cuda.synchronize()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
        logits = model(graph, features)
        loss = loss_fcn(logits[train_mask], labels[train_mask])
        if timing:
            f1.record()
        optimizer.zero_grad()
        if timing:
            b0.record()
        loss.backward()
        if timing:
            b1.record()
            torch.cuda.synchronize()
            fwd_ms += f0.elapsed_time(f1)
            bwd_ms += b0.elapsed_time(b1)
            timed += 1
        optimizer.step()
        losses.append(float(loss.item()))
        if log and (epoch % 10 == 0 or epoch == epochs - 1):
            log("Epoch %04d/%04d| Loss %.4f" % (epoch, epochs, losses[-1]))
    report = {"model": type(model).__name__, "epochs_measured": timed, "losses": losses,
              "avg_forward_ms": fwd_ms / timed if timed else None, "avg_backward_ms": bwd_ms / timed if timed else None}
    if timed:
        report["total_per_epoch_ms"] = report["avg_forward_ms"] + report["avg_backward_ms"]
    return report


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dataset", default="flickr", choices=sorted(SHAPES))
    ap.add_argument("--model", default="sage", choices=sorted(MODELS))
    ap.add_argument("--maxk", type=int, default=32)
    ap.add_argument("--hidden_dim", type=int, default=256)
    ap.add_argument("--hidden_layers", type=int, default=3)
    ap.add_argument("--epochs", type=int, default=30)
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--in_size", type=int, default=128)
    ap.add_argument("--classes", type=int, default=16)
    ap.add_argument("--dropout", type=float, default=0.5)
    ap.add_argument("--norm", action="store_true")
    a = ap.parse_args()
    dev = torch.device("cuda")
    graph, x, y, masks = synthetic_task(a.dataset, a.scale, a.in_size, a.classes, dev)
    model = MODELS[a.model](a.in_size, a.hidden_dim, a.hidden_layers, a.classes, maxk=a.maxk, feat_drop=a.dropout,
                            norm=a.norm, graph_name=a.dataset).to(dev)
    t0 = time.time()
    rep = train(graph, x, y, masks, model, epochs=a.epochs, multi_label=a.dataset in MULTI_LABEL)
    rep.update(dataset=a.dataset, nodes=graph.num_nodes(), edges=graph.num_edges(), maxk=a.maxk,
               hidden_dim=a.hidden_dim, hidden_layers=a.hidden_layers, wall_s=time.time() - t0)
    rep.pop("losses")
    print(json.dumps(rep))


if __name__ == "__main__":
    main()
