"""Synthetic graphs of the shapes BASELINE.json names.

There is no network for the real datasets, so every measurement runs on seeded synthetic
CSR graphs with the node / edge counts of Flickr, Reddit, Yelp, ogbn-proteins, ogbn-products.
Everything is plain torch so it runs on the CPU (tests) and on the GPU (bench, 114.6 M edges).
"""
import torch

# name -> (nodes, edges)   (BASELINE.json configs)
SHAPES = {
    "flickr": (89_250, 899_756),
    "reddit": (232_965, 114_615_892),
    "yelp": (716_847, 13_954_819),
    "proteins": (132_534, 39_561_252),
    "products": (2_449_029, 61_859_140),
}


def _gen(device, seed):
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    return g


def synth_graph(n, e, seed=123, kind="uniform", device="cpu", values="uniform"):
    """Returns dict(indptr int32[n+1], indices int32[e], values fp32[e], v_num, e_num).

    kind="uniform" : every row gets floor(e/n) or ceil(e/n) neighbours, columns iid uniform.
    kind="powerlaw": row degrees ~ lognormal(sigma=1) rescaled to sum e (min 1), columns drawn
                     proportionally to degree (hubs are popular sources too).
    values="uniform": U[0,1) seed-derived (kernels/main.cu:83-84); "ones": all 1 (model path,
                     maxk_models_integrated.py:139).
    Columns are sorted within each row (CSR canonical order, as dataset_gen.py produces).
    """
    device = torch.device(device)
    g = _gen(device, seed)
    if kind == "uniform":
        deg = torch.full((n,), e // n, dtype=torch.int64, device=device)
        deg[: e % n] += 1
        cols = torch.randint(0, n, (e,), generator=g, device=device, dtype=torch.int64)
    elif kind == "powerlaw":
        w = torch.exp(torch.randn(n, generator=g, device=device, dtype=torch.float64))
        deg = torch.clamp((w / w.sum() * e).floor().to(torch.int64), min=1)
        diff = int(e - deg.sum().item())
        if diff > 0:
            top = torch.argsort(w, descending=True)[: max(1, min(n, diff))]
            deg[top] += diff // top.numel()
            deg[top[: diff % top.numel()]] += 1
        elif diff < 0:
            order = torch.argsort(deg, descending=True)
            need, i = -diff, 0
            while need > 0:
                r = order[i % n]
                take = min(need, int(deg[r].item()) - 1)
                deg[r] -= take
                need -= take
                i += 1
        cdf = torch.cumsum(deg.to(torch.float64), 0)
        u = torch.rand(e, generator=g, device=device, dtype=torch.float64) * cdf[-1]
        cols = torch.clamp(torch.searchsorted(cdf, u, right=True), max=n - 1)
    else:
        raise ValueError("unknown kind %r" % kind)
    rows = torch.repeat_interleave(torch.arange(n, device=device, dtype=torch.int64), deg)
    key, _ = torch.sort(rows * n + cols)
    indices = (key % n).to(torch.int32)
    indptr = torch.zeros(n + 1, dtype=torch.int64, device=device)
    indptr[1:] = torch.cumsum(deg, 0)
    if values == "ones":
        vals = torch.ones(e, dtype=torch.float32, device=device)
    else:
        vals = torch.rand(e, generator=g, device=device, dtype=torch.float32)
    return {"indptr": indptr.to(torch.int32), "indices": indices, "values": vals, "v_num": n, "e_num": int(e)}


def symmetrize(graph):
    """A + A^T with self loops, duplicates removed (what dataset_gen.py:44-98 does); values -> ones."""
    n = graph["v_num"]
    indptr, indices = graph["indptr"].long(), graph["indices"].long()
    dev = indices.device
    rows = torch.repeat_interleave(torch.arange(n, device=dev), indptr[1:] - indptr[:-1])
    loops = torch.arange(n, device=dev)
    key = torch.unique(torch.cat([rows * n + indices, indices * n + rows, loops * n + loops]))
    r, c = key // n, key % n
    new_ptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    new_ptr[1:] = torch.cumsum(torch.bincount(r, minlength=n), 0)
    return {"indptr": new_ptr.to(torch.int32), "indices": c.to(torch.int32),
            "values": torch.ones(key.numel(), dtype=torch.float32, device=dev), "v_num": n, "e_num": int(key.numel())}


def shape_graph(name, scale=1.0, **kw):
    """Graph with the node/edge counts of a named dataset, optionally scaled down (same avg degree)."""
    n, e = SHAPES[name]
    n = max(2, int(n * scale))
    e = max(1, int(e * scale))
    return synth_graph(n, e, **kw)
