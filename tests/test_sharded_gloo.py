"""CPU, world_size 2 on gloo: the row-sharded layer (partition + collectives) against the single-process oracle.

The compute backend injected here is the oracle (tests may do that); the product default is the CUDA
backend, covered by tests/test_gpu_sharded.py on real GPUs.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from helpers import assert_close
from sharded import (ShardedMaxKAggregation, ShardedMaxKSAGE, allreduce_gradients, shard_columns, shard_rows,
                     sharded_maxk_spgemm, slab_rows)
from synth_graphs import synth_graph


class OracleCompute:
    def topk(self, x, k):
        v, c = oracle.topk(x.numpy(), k, 2)
        return torch.from_numpy(v), torch.from_numpy(c.astype(np.uint8))

    def spgemm(self, g, vals, sel, row_div=None):
        out = oracle.spgemm_fwd(g["indptr"].numpy(), g["indices"].numpy(), g["values"].numpy(), vals.numpy(), sel.numpy(),
                                deg=None if row_div is None else row_div.numpy())
        return torch.from_numpy(out)

    def sspmm(self, g, grad, sel, row_div=None):
        gs = oracle.sspmm_bwd(g["indptr"].numpy(), g["indices"].numpy(), g["values"].numpy(), grad.numpy(), sel.numpy(),
                              deg=None if row_div is None else row_div.numpy())
        return torch.from_numpy(gs)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _problem(n=203, e=4000, k=16):
    g = synth_graph(n, e, seed=11, kind="powerlaw")
    gen = torch.Generator().manual_seed(3)
    x, grad = torch.randn(n, 256, generator=gen), torch.rand(n, 256, generator=gen)
    deg = torch.clamp((g["indptr"][1:] - g["indptr"][:-1]).float(), min=1)
    return g, x, grad, deg, k


def _worker(rank, world, port, mode, use_div, result_dir, partition="rows"):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g, x, grad, deg, k = _problem()
        layer = ShardedMaxKAggregation(g, k, backward_mode=mode, compute=OracleCompute(), row_div=deg if use_div else None,
                                       partition=partition)
        x_local, g_local = layer.local_slab(x), layer.local_slab(grad)
        xl = x_local.clone().requires_grad_(True)
        out = sharded_maxk_spgemm(xl, layer)
        out.backward(g_local)
        gs = layer.backward(g_local)
        np.savez(os.path.join(result_dir, "rank%d.npz" % rank), out=out.detach().numpy(), gs=gs.numpy(),
                 xgrad=xl.grad.numpy(), wire_fwd=layer.wire_bytes()["forward"], wire_bwd=layer.wire_bytes()["backward"],
                 lo=layer.rows["row_lo"], hi=layer.rows["row_hi"], m=layer.m, edges=layer.rows["e_num"])
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode,use_div", [("reduce_scatter", False), ("reduce_scatter", True), ("allgather", True)])
def test_sharded_layer_matches_single_process_oracle(tmp_path, mode, use_div):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), mode, use_div, str(tmp_path)), nprocs=world, join=True)
    g, x, grad, deg, k = _problem()
    n, m = g["v_num"], slab_rows(g["v_num"], world)
    ip, ix, va = (g[t].numpy() for t in ("indptr", "indices", "values"))
    vals, cols = oracle.topk(x.numpy(), k, 2)
    sel = cols.astype(np.uint8)
    d = deg.numpy() if use_div else None
    exp_out = oracle.spgemm_fwd(ip, ix, va, vals, sel, deg=d)
    exp_gs = oracle.sspmm_bwd(ip, ix, va, grad.numpy(), sel, deg=d)
    exp_xgrad = oracle.scatter_dense(exp_gs, cols)
    for rank in range(world):
        r = np.load(os.path.join(str(tmp_path), "rank%d.npz" % rank))
        lo, hi = rank * m, min(n, rank * m + m)
        assert_close(r["out"][: hi - lo], exp_out[lo:hi], "rank %d forward" % rank)
        assert_close(r["gs"][: hi - lo], exp_gs[lo:hi], "rank %d backward" % rank)
        assert_close(r["xgrad"][: hi - lo], exp_xgrad[lo:hi], "rank %d autograd" % rank)
        assert int(r["wire_fwd"]) == m * k * 5
        assert int(r["wire_bwd"]) == (m * 256 * 4 if mode == "allgather" else m * k * 4)


@pytest.mark.parametrize("mode", ["reduce_scatter", "allgather"])
def test_edge_balanced_partition_matches_single_process_oracle(tmp_path, mode):
    """partition="nnz": slab boundaries from the prefix sum of the degrees, slabs padded to a common height,
    column ids remapped into the padded numbering -- same results as the single-process oracle, and the
    edge counts of the two ranks differ by less than one hub row."""
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), mode, True, str(tmp_path), "nnz"), nprocs=world, join=True)
    g, x, grad, deg, k = _problem()
    ip, ix, va = (g[t].numpy() for t in ("indptr", "indices", "values"))
    vals, cols = oracle.topk(x.numpy(), k, 2)
    sel = cols.astype(np.uint8)
    exp_out = oracle.spgemm_fwd(ip, ix, va, vals, sel, deg=deg.numpy())
    exp_gs = oracle.sspmm_bwd(ip, ix, va, grad.numpy(), sel, deg=deg.numpy())
    exp_xgrad = oracle.scatter_dense(exp_gs, cols)
    res = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % rank)) for rank in range(world)]
    assert int(res[0]["lo"]) == 0 and int(res[0]["hi"]) == int(res[1]["lo"]) and int(res[1]["hi"]) == g["v_num"]
    assert int(res[0]["m"]) == int(res[1]["m"]) == max(int(r["hi"]) - int(r["lo"]) for r in res)
    assert int(res[0]["edges"]) + int(res[1]["edges"]) == g["e_num"]
    assert abs(int(res[0]["edges"]) - int(res[1]["edges"])) <= int(np.diff(ip).max())
    uniform_split = abs(2 * int(ip[slab_rows(g["v_num"], world)]) - g["e_num"])
    assert abs(int(res[0]["edges"]) - int(res[1]["edges"])) <= uniform_split       # never worse than equal rows
    for rank, r in enumerate(res):
        lo, hi = int(r["lo"]), int(r["hi"])
        assert_close(r["out"][: hi - lo], exp_out[lo:hi], "rank %d forward" % rank)
        assert_close(r["gs"][: hi - lo], exp_gs[lo:hi], "rank %d backward" % rank)
        assert_close(r["xgrad"][: hi - lo], exp_xgrad[lo:hi], "rank %d autograd" % rank)
        assert not r["out"][hi - lo:].any() and not r["gs"][hi - lo:].any()       # padding rows stay zero


def test_balanced_bounds_and_padded_positions():
    from sharded import balanced_bounds, padded_position, uniform_bounds
    ip = torch.tensor([0, 0, 0, 50, 50, 51, 52, 100, 100])           # 8 rows, 100 edges, two hubs
    assert balanced_bounds(ip, 1) == [0, 8]
    b = balanced_bounds(ip, 2)
    assert b[0] == 0 and b[-1] == 8 and b == sorted(b)
    assert abs(int(ip[b[1]]) - 50) <= 50
    b4 = balanced_bounds(ip, 4)
    assert len(b4) == 5 and b4 == sorted(b4) and b4[-1] == 8
    # identity for the equal-row partition, owner * m + offset otherwise (empty slabs own nothing)
    ids = torch.arange(10)
    assert torch.equal(padded_position(ids, uniform_bounds(10, 3), 4), ids)
    pos = padded_position(torch.arange(9), [0, 5, 5, 5, 9], 5)
    assert pos.tolist() == [0, 1, 2, 3, 4, 15, 16, 17, 18]
    # a hub-heavy graph: every rank gets its share of edges, all of them exactly once
    g = synth_graph(300, 9000, seed=5, kind="powerlaw")
    for world in (2, 4, 8):
        bounds = balanced_bounds(g["indptr"], world)
        per_rank = [shard_rows(g, world, r, bounds)["e_num"] for r in range(world)]
        assert sum(per_rank) == 9000
        assert max(per_rank) <= 9000 // world + int((g["indptr"][1:] - g["indptr"][:-1]).max())
        m = shard_rows(g, world, 0, bounds)["v_num"]
        assert all(shard_rows(g, world, r, bounds)["v_num"] == m for r in range(world))
        assert sum(shard_columns(g, world, r, bounds)["e_num"] for r in range(world)) == 9000


def test_partition_helpers_cover_the_graph_exactly():
    g = synth_graph(101, 1500, seed=2)
    for world in (1, 2, 4, 8):
        m = slab_rows(101, world)
        edges = 0
        for rank in range(world):
            r = shard_rows(g, world, rank)
            assert r["indptr"].numel() == m + 1 and int(r["indptr"][-1]) == r["e_num"] == r["indices"].numel()
            edges += r["e_num"]
            c = shard_columns(g, world, rank)
            assert c["indptr"].numel() == world * m + 1
            if c["e_num"]:
                assert int(c["indices"].min()) >= 0 and int(c["indices"].max()) < m
        assert edges == 1500
        assert sum(shard_columns(g, world, r)["e_num"] for r in range(world)) == 1500


# ------------------------------------------------------------------------------------ sharded GraphSAGE
def _sage_reference(params, a, deg, x, k):
    """Single-process restatement of ShardedMaxKSAGE.forward in plain torch (CPU)."""
    h = x @ params["lin_in.weight"].t() + params["lin_in.bias"]
    n_layers = sum(1 for key in params if key.startswith("fc_neigh."))
    for i in range(n_layers):
        v, c = oracle.topk(h.detach().numpy(), k, 2)
        mask = torch.zeros_like(h).scatter_(1, torch.from_numpy(c.astype(np.int64)), 1.0)
        hs = h * mask
        agg = (a @ hs) / deg.unsqueeze(1)
        h = hs @ params["fc_self.%d.weight" % i].t() + params["fc_self.%d.bias" % i] + agg @ params["fc_neigh.%d.weight" % i].t()
    return h @ params["lin_out.weight"].t() + params["lin_out.bias"]


def _sage_worker(rank, world, port, result_dir, partition="rows"):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g, x, grad, deg, k = _problem(n=150, e=2500, k=16)
        torch.manual_seed(7)                                   # identical weights on every rank
        model = ShardedMaxKSAGE(g, 256, 256, 2, 5, maxk=k, compute=OracleCompute(), partition=partition)
        rows = model.agg.valid_rows()
        x_local = model.agg.local_slab(x)
        out = model(x_local)
        out[:rows].square().sum().backward()                   # loss over the real rows of this rank
        allreduce_gradients(model)
        np.savez(os.path.join(result_dir, "sage%d.npz" % rank), out=out.detach().numpy(),
                 lo=model.agg.rows["row_lo"], hi=model.agg.rows["row_hi"],
                 **{"g_" + name: p.grad.numpy() for name, p in model.named_parameters()},
                 **{"p_" + name: p.detach().numpy() for name, p in model.named_parameters()})
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("partition", ["rows", "nnz"])
def test_sharded_sage_matches_single_process_model(tmp_path, partition):
    world = 2
    mp.spawn(_sage_worker, args=(world, _free_port(), str(tmp_path), partition), nprocs=world, join=True)
    g, x, grad, deg, k = _problem(n=150, e=2500, k=16)
    n = g["v_num"]
    r0 = np.load(os.path.join(str(tmp_path), "sage0.npz"))
    params = {key[2:]: torch.from_numpy(r0[key]).requires_grad_(True) for key in r0.files if key.startswith("p_")}
    a = torch.sparse_csr_tensor(g["indptr"].long(), g["indices"].long(), g["values"], size=(n, n))
    ref = _sage_reference(params, a, deg, x, k)
    ref.square().sum().backward()
    for rank in range(world):
        r = np.load(os.path.join(str(tmp_path), "sage%d.npz" % rank))
        lo, hi = int(r["lo"]), int(r["hi"])
        np.testing.assert_allclose(r["out"][: hi - lo], ref.detach().numpy()[lo:hi], rtol=2e-4, atol=2e-4)
        for name, p in params.items():                      # summed over ranks == gradient of the full loss
            np.testing.assert_allclose(r["g_" + name], p.grad.numpy(), rtol=2e-3, atol=2e-3, err_msg=name)
