"""Host-buffer entry point of the MaxK layer: features and upstream gradient arrive in pinned HOST
memory, the aggregate and the sampled gradient go back to pinned HOST memory.

The three stages of a call overlap on three CUDA streams, slab by slab:

    h2d stream : x slab 0..S-1, then grad slab 0..S-1                      (PCIe host->device)
    compute    : top-k(x slab j) as soon as slab j landed (top-k is row local)
                 forward SpGEMM over row slab j (needs the whole CBSR)      -> out slab j
                 backward SSpMM over source-row slab j as soon as grad slab j landed (accumulating)
    d2h stream : out slab j as soon as its forward finished, gs at the end  (PCIe device->host)

so one call costs about max(PCIe in, compute, PCIe out) instead of their sum, and consecutive calls
overlap as well (double-buffered device staging).  Every kernel is the same C-ABI call the
device-resident path uses; nothing here computes on the host.
"""
import torch

import maxk_cuda_kernels as K


def bind_host_to_gpu(nvml_index):
    """Pin this process to the CPU cores next to its GPU (NVML's ideal affinity) BEFORE it allocates pinned host
    buffers: first-touch places them on the GPU's NUMA node, so the host<->device copies of the pipeline do not
    cross sockets (one process per GPU under torchrun is otherwise scheduled anywhere).  Returns the core count
    it bound to, or 0 when NVML / sched_setaffinity is unavailable."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(nvml_index))
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return 0


class HostStagedMaxKLayer:
    def __init__(self, indptr, indices, values, k, dim=256, slabs=8, device=None):
        self.dev = indices.device if device is None else device
        self.ip, self.ix, self.va = indptr, indices, values
        self.n = indptr.numel() - 1
        self.k, self.dim = int(k), int(dim)
        s = max(1, min(int(slabs), self.n))
        step = (self.n + s - 1) // s
        self.bounds = [(lo, min(lo + step, self.n)) for lo in range(0, self.n, step)]
        self.h2d, self.comp, self.d2h = (torch.cuda.Stream(self.dev) for _ in range(3))
        mk = lambda *shape, dt=torch.float32: torch.empty(*shape, dtype=dt, device=self.dev)
        # two sets of device staging buffers so that call i+1 can start copying while call i computes
        self.sets = [{"x": mk(self.n, dim), "g": mk(self.n, dim), "out": mk(self.n, dim), "gs": mk(self.n, self.k),
                      "vals": mk(self.n, self.k), "sel": mk(self.n, self.k, dt=torch.uint8),
                      "done": None} for _ in range(2)]
        # row plans of the slabs (the forward walks its rows in plan order, csrc/plan.cu)
        self.plans = [K.build_plan(indptr[lo:hi], indptr[lo + 1:hi + 1]) for lo, hi in self.bounds]
        self.calls = 0

    def run(self, hx, hg, hout, hgs, block_current_stream=True):
        """hx, hg: pinned [N, dim] fp32 host tensors; hout [N, dim], hgs [N, k]: pinned host outputs.
        Asynchronous: returns an event that is recorded when hout/hgs are complete.  With
        block_current_stream=False consecutive calls overlap (the caller waits on the returned event)."""
        b = self.sets[self.calls % 2]
        self.calls += 1
        cur = torch.cuda.current_stream(self.dev)
        for s in (self.h2d, self.comp, self.d2h):
            s.wait_stream(cur)
        if b["done"] is not None:                      # this buffer set's previous call must have drained
            self.h2d.wait_event(b["done"])
        ev = lambda: torch.cuda.Event()
        ex, eg, ef = [ev() for _ in self.bounds], [ev() for _ in self.bounds], [ev() for _ in self.bounds]
        with torch.cuda.stream(self.h2d):
            for j, (lo, hi) in enumerate(self.bounds):
                b["x"][lo:hi].copy_(hx[lo:hi], non_blocking=True)
                ex[j].record(self.h2d)
            for j, (lo, hi) in enumerate(self.bounds):
                b["g"][lo:hi].copy_(hg[lo:hi], non_blocking=True)
                eg[j].record(self.h2d)
        with torch.cuda.stream(self.comp):
            for j, (lo, hi) in enumerate(self.bounds):
                self.comp.wait_event(ex[j])
                K.topk_cbsr(b["x"][lo:hi], self.k, order=K.ORDER_BANKED, out_values=b["vals"][lo:hi],
                            out_sel=b["sel"][lo:hi])
            for j, (lo, hi) in enumerate(self.bounds):
                K.spgemm_forward_csr(self.ip[lo:hi], self.ip[lo + 1:hi + 1], self.ix, self.va, b["vals"], b["sel"],
                                     out_dim=self.dim, out=b["out"][lo:hi], plan=self.plans[j])
                ef[j].record(self.comp)
            b["gs"].zero_()
            for j, (lo, hi) in enumerate(self.bounds):
                self.comp.wait_event(eg[j])
                K.sspmm_backward_csr(self.ip[lo:hi], self.ip[lo + 1:hi + 1], self.ix, self.va, b["g"][lo:hi], b["sel"],
                                     out=b["gs"], accumulate=True)
            e_bwd = ev()
            e_bwd.record(self.comp)
        with torch.cuda.stream(self.d2h):
            for j, (lo, hi) in enumerate(self.bounds):
                self.d2h.wait_event(ef[j])
                hout[lo:hi].copy_(b["out"][lo:hi], non_blocking=True)
            self.d2h.wait_event(e_bwd)
            hgs.copy_(b["gs"], non_blocking=True)
            done = ev()
            done.record(self.d2h)
        b["done"] = done
        if block_current_stream:
            cur.wait_event(done)
        return done

    launches_per_call = property(lambda self: len(self.bounds) * 4)   # top-k + fwd + (bwd, long-row bwd) per slab


class ShardedHostStagedLayer:
    """The same three-stream slab pipeline for ONE RANK of the row-sharded layer (sharded.py): every GPU has its
    own PCIe link, so each rank streams its row slab of the features / upstream gradient in and its slab of the
    aggregate / sampled gradient out while it computes.

        h2d stream : x chunk 0..S-1, then grad chunk 0..S-1
        compute    : top-k(x chunk j) as its chunk lands -> all_gather of the rank's CBSR slab (NCCL, on this stream)
                     forward SpGEMM over row chunk j -> out chunk j
                     SSpMM over source-row chunk j as its gradient lands (accumulating into the full-size partial)
                     reduce_scatter(sum) of the partial -> this rank's gs slab
        d2h stream : out chunk j as soon as its forward finished, gs at the end
    """

    def __init__(self, layer, dim=256, slabs=4):
        import torch.distributed as dist
        self.dist = dist
        self.layer = layer
        rows = layer.rows
        self.ip, self.ix, self.va = rows["indptr"], rows["indices"], rows["values"]
        self.dev = self.ix.device
        self.m, self.k, self.world, self.dim = layer.m, layer.k, layer.world, int(dim)
        s = max(1, min(int(slabs), self.m))
        step = (self.m + s - 1) // s
        self.bounds = [(lo, min(lo + step, self.m)) for lo in range(0, self.m, step)]
        self.plans = [K.build_plan(self.ip[lo:hi], self.ip[lo + 1:hi + 1]) for lo, hi in self.bounds]
        self.h2d, self.comp, self.d2h = (torch.cuda.Stream(self.dev) for _ in range(3))
        mk = lambda *shape, dt=torch.float32: torch.empty(*shape, dtype=dt, device=self.dev)
        m, k, w = self.m, self.k, self.world
        self.sets = [{"x": mk(m, dim), "g": mk(m, dim), "out": mk(m, dim), "gs": mk(m, k), "vals": mk(m, k),
                      "sel": mk(m, k, dt=torch.uint8), "vals_full": mk(w * m, k), "sel_full": mk(w * m, k, dt=torch.uint8),
                      "partial": mk(w * m, k), "done": None} for _ in range(2)]
        self.calls = 0

    def run(self, hx, hg, hout, hgs, block_current_stream=True):
        """hx, hg: pinned [m, dim] host tensors (this rank's slab); hout [m, dim], hgs [m, k]: pinned host outputs."""
        dist, group = self.dist, self.layer.group
        b = self.sets[self.calls % 2]
        self.calls += 1
        cur = torch.cuda.current_stream(self.dev)
        for s in (self.h2d, self.comp, self.d2h):
            s.wait_stream(cur)
        if b["done"] is not None:
            self.h2d.wait_event(b["done"])
        ev = lambda: torch.cuda.Event()
        ex, eg, ef = [ev() for _ in self.bounds], [ev() for _ in self.bounds], [ev() for _ in self.bounds]
        with torch.cuda.stream(self.h2d):
            for j, (lo, hi) in enumerate(self.bounds):
                b["x"][lo:hi].copy_(hx[lo:hi], non_blocking=True)
                ex[j].record(self.h2d)
            for j, (lo, hi) in enumerate(self.bounds):
                b["g"][lo:hi].copy_(hg[lo:hi], non_blocking=True)
                eg[j].record(self.h2d)
        with torch.cuda.stream(self.comp):
            for j, (lo, hi) in enumerate(self.bounds):
                self.comp.wait_event(ex[j])
                K.topk_cbsr(b["x"][lo:hi], self.k, order=K.ORDER_BANKED, out_values=b["vals"][lo:hi], out_sel=b["sel"][lo:hi])
            dist.all_gather_into_tensor(b["vals_full"], b["vals"], group=group)
            dist.all_gather_into_tensor(b["sel_full"], b["sel"], group=group)
            for j, (lo, hi) in enumerate(self.bounds):
                K.spgemm_forward_csr(self.ip[lo:hi], self.ip[lo + 1:hi + 1], self.ix, self.va, b["vals_full"], b["sel_full"],
                                     out_dim=self.dim, out=b["out"][lo:hi], plan=self.plans[j])
                ef[j].record(self.comp)
            b["partial"].zero_()
            for j, (lo, hi) in enumerate(self.bounds):
                self.comp.wait_event(eg[j])
                K.sspmm_backward_csr(self.ip[lo:hi], self.ip[lo + 1:hi + 1], self.ix, self.va, b["g"][lo:hi], b["sel_full"],
                                     out=b["partial"], accumulate=True)
            dist.reduce_scatter_tensor(b["gs"], b["partial"], op=dist.ReduceOp.SUM, group=group)
            e_bwd = ev()
            e_bwd.record(self.comp)
        with torch.cuda.stream(self.d2h):
            for j, (lo, hi) in enumerate(self.bounds):
                self.d2h.wait_event(ef[j])
                hout[lo:hi].copy_(b["out"][lo:hi], non_blocking=True)
            self.d2h.wait_event(e_bwd)
            hgs.copy_(b["gs"], non_blocking=True)
            done = ev()
            done.record(self.d2h)
        b["done"] = done
        if block_current_stream:
            cur.wait_event(done)
        return done

    launches_per_call = property(lambda self: len(self.bounds) * 4)
