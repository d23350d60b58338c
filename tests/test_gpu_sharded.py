"""GPU: the row-sharded layer on the CUDA backend over NCCL: world 1 always, worlds 2 / 4 / 8 when that many GPUs
are visible (gpurun --gpus N); both forward exchanges (top-k writing into peer memory, NCCL all_gather)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from helpers import assert_close
from sharded import ShardedMaxKAggregation, sharded_maxk_spgemm
from synth_graphs import synth_graph

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _problem(n=3001, e=150000, k=32):
    g = synth_graph(n, e, seed=21, kind="powerlaw")
    gen = torch.Generator().manual_seed(5)
    x, grad = torch.randn(n, 256, generator=gen), torch.rand(n, 256, generator=gen)
    deg = torch.clamp((g["indptr"][1:] - g["indptr"][:-1]).float(), min=1)
    return g, x, grad, deg, k


def _expected(g, x, grad, deg, k):
    ip, ix, va = (g[t].numpy() for t in ("indptr", "indices", "values"))
    vals, cols = oracle.topk(x.numpy(), k, 2)
    sel = cols.astype(np.uint8)
    out = oracle.spgemm_fwd(ip, ix, va, vals, sel, deg=deg.numpy())
    gs = oracle.sspmm_bwd(ip, ix, va, grad.numpy(), sel, deg=deg.numpy())
    return out, gs, oracle.scatter_dense(gs, cols)


def _run_rank(rank, world, port, mode, result_dir, partition="rows", gather="auto"):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        g, x, grad, deg, k = _problem()
        gc = {key: (v.cuda() if isinstance(v, torch.Tensor) else v) for key, v in g.items()}
        layer = ShardedMaxKAggregation(gc, k, backward_mode=mode, row_div=deg.cuda(),
                                       partition=partition, gather=gather)             # default compute = CUDA kernels
        x_local, g_local = layer.local_slab(x.cuda()), layer.local_slab(grad.cuda())
        xl = x_local.clone().requires_grad_(True)
        out = sharded_maxk_spgemm(xl, layer)
        out.backward(g_local)
        gs = layer.backward(g_local)
        # a second and third step through the recycled peer buffers must give the same rows
        out2 = layer.forward(x_local)
        out3 = layer.forward(x_local)
        torch.cuda.synchronize()
        assert torch.equal(out2, out3) and torch.allclose(out2, out.detach(), rtol=1e-6, atol=1e-7)
        np.savez(os.path.join(result_dir, "rank%d.npz" % rank), out=out.detach().cpu().numpy(), gs=gs.cpu().numpy(),
                 xgrad=xl.grad.cpu().numpy(), lo=layer.rows["row_lo"], hi=layer.rows["row_hi"], edges=layer.rows["e_num"],
                 gather=layer.gather + " / " + layer.reduce)
    finally:
        dist.destroy_process_group()


def _check(tmp_path, world):
    g, x, grad, deg, k = _problem()
    exp_out, exp_gs, exp_xgrad = _expected(g, x, grad, deg, k)
    for rank in range(world):
        r = np.load(os.path.join(str(tmp_path), "rank%d.npz" % rank))
        lo, hi = int(r["lo"]), int(r["hi"])
        assert_close(r["out"][: hi - lo], exp_out[lo:hi], "rank %d forward" % rank)
        assert_close(r["gs"][: hi - lo], exp_gs[lo:hi], "rank %d backward" % rank, rtol=2e-5)
        assert_close(r["xgrad"][: hi - lo], exp_xgrad[lo:hi], "rank %d autograd" % rank, rtol=2e-5)


@pytest.mark.parametrize("mode", ["reduce_scatter", "allgather"])
def test_world1_nccl(tmp_path, mode):
    mp.spawn(_run_rank, args=(1, _free_port(), mode, str(tmp_path)), nprocs=1, join=True)
    _check(tmp_path, 1)


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("mode,gather", [("reduce_scatter", "auto"), ("allgather", "auto"), ("reduce_scatter", "nccl")])
def test_multi_gpu_nccl(tmp_path, world, mode, gather):
    """3001 rows over 2 / 4 / 8 ranks: the last slab is padded (3001 % 8 = 1), every rank's rows against the oracle."""
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs (gpurun --gpus %d)" % (world, world))
    mp.spawn(_run_rank, args=(world, _free_port(), mode, str(tmp_path), "rows", gather), nprocs=world, join=True)
    _check(tmp_path, world)
    print("forward exchange:", str(np.load(os.path.join(str(tmp_path), "rank0.npz"))["gather"]))


@pytest.mark.parametrize("world", [2, 8])
@pytest.mark.parametrize("mode", ["reduce_scatter", "allgather"])
def test_multi_gpu_nccl_edge_balanced_partition(tmp_path, world, mode):
    """partition="nnz" on a power-law graph: same results, and the ranks hold (almost) the same number of edges."""
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs (gpurun --gpus %d)" % (world, world))
    mp.spawn(_run_rank, args=(world, _free_port(), mode, str(tmp_path), "nnz"), nprocs=world, join=True)
    _check(tmp_path, world)
    e = [int(np.load(os.path.join(str(tmp_path), "rank%d.npz" % r))["edges"]) for r in range(world)]
    g = _problem()[0]
    assert sum(e) == g["e_num"] and max(e) - min(e) <= 2 * int((g["indptr"][1:] - g["indptr"][:-1]).max())
