"""Developer tool (GPU): BASELINE.json config 3 -- Yelp-shape GCN MaxK k=32, full-graph training epochs -- with the
CUDA-event forward / backward split of the reference trainer (all_train.py:118-149) and, from the torch profiler's
kernel table, the share of an epoch spent in our three hot-path kernels against cuBLAS / LayerNorm / elementwise.

    python tools/epoch_profile.py [--dataset yelp --model gcn --maxk 32 --epochs 16]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "spgemm-prunning_b200")):
    sys.path.insert(0, p)

import torch  # noqa: E402
from maxk_gnn_training import MODELS, MULTI_LABEL, synthetic_task, train  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--dataset", default="yelp")
ap.add_argument("--model", default="gcn")
ap.add_argument("--maxk", type=int, default=32)
ap.add_argument("--epochs", type=int, default=16)
ap.add_argument("--hidden_layers", type=int, default=3)
ap.add_argument("--scale", type=float, default=1.0)
a = ap.parse_args()
dev = torch.device("cuda")
graph, x, y, masks = synthetic_task(a.dataset, a.scale, 128, 16, dev)
model = MODELS[a.model](128, 256, a.hidden_layers, 16, maxk=a.maxk, feat_drop=0.5, norm=True, graph_name=a.dataset).to(dev)
rep = train(graph, x, y, masks, model, epochs=a.epochs, warmup_epochs=6, multi_label=a.dataset in MULTI_LABEL, log=None)
rep.pop("losses")
rep.update(dataset=a.dataset, nodes=graph.num_nodes(), edges=graph.num_edges(), maxk=a.maxk, hidden_layers=a.hidden_layers)
print(json.dumps(rep), flush=True)

# kernel shares of two more epochs
try:
    from torch.profiler import ProfilerActivity, profile
    opt = torch.optim.Adam(model.parameters(), lr=0.01)
    lossf = torch.nn.functional.binary_cross_entropy_with_logits if a.dataset in MULTI_LABEL else torch.nn.CrossEntropyLoss()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(2):
            out = model(graph, x)
            loss = lossf(out[masks[0]], y[masks[0]])
            opt.zero_grad()
            loss.backward()
            opt.step()
        torch.cuda.synchronize()
    rows = [(e.key, e.device_time_total / 2e3, e.count // 2) for e in prof.key_averages() if e.device_time_total > 0]
    rows.sort(key=lambda r: -r[1])
    total = sum(r[1] for r in rows)
    ours = sum(r[1] for r in rows if "maxk::" in r[0])
    print("per-epoch device time %.3f ms; our kernels %.3f ms (%.1f %%)" % (total, ours, 100 * ours / total))
    for key, ms, cnt in rows[:18]:
        print("  %8.3f ms  %5.1f %%  x%-3d %s" % (ms, 100 * ms / total, cnt, key[:110]))
except Exception as ex:
    print("torch profiler unavailable:", repr(ex)[:200])
