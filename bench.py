"""bench.py -- the contract benchmark of the MaxK aggregation hot path.

    python bench.py --gpus 1 --steps K --warmup W              (ours, one B200)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   (ours, row-sharded)
    python bench.py --impl reference ...                       (the reference's CPU path, host cores)

A step = one pass of the hot path over one batch of synthetic input on the Reddit-shape graph
(232,965 nodes, 114.6 M edges, hidden 256, k = 32; BASELINE.json configs[1]):
    top-k -> CBSR, forward SpGEMM, backward SSpMM.
`value` is the whole-job algorithmic throughput: the compulsory bytes of SURVEY.md 8(d)
(B_topk + B_fwd + B_bwd, every operand read once, every result written once) divided by the device
time of a step with inputs resident in HBM.  `e2e` is the same through the operator API with HOST
buffers (pinned): features and upstream gradient copied in, aggregate and sampled gradient copied out.
Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "spgemm-prunning_b200"), os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

DIM = 256
KERNELS_PER_STEP = 5      # topk_cbsr, spgemm_fwd + its long-row kernel, sspmm_bwd + its long-row kernel
METRIC = "MaxK top-k + fwd SpGEMM + bwd SSpMM algorithmic HBM throughput, Reddit shape k=32"


def metric_name(args):
    return METRIC if (args.shape == "reddit" and args.k == 32) else METRIC.replace("Reddit shape k=32", "%s shape k=%d" % (args.shape, args.k))


def layer_bytes(n, e, k, dim=DIM):
    """compulsory traffic of one layer (top-k, forward, backward)."""
    b_topk = n * dim * 4 + n * k * 5
    b_fwd = (n + 1) * 4 + e * 8 + n * k * 5 + n * dim * 4
    b_bwd = (n + 1) * 4 + e * 8 + n * dim * 4 + n * k * 1 + n * k * 4
    return b_topk, b_fwd, b_bwd


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(workload, kernel):
    """dram read+write bytes per launch from the committed ncu --set full capture, or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            return json.load(f).get(workload, {}).get(kernel)
    except Exception:
        return None


class ClockSampler:
    """Samples SM clock + throttle reasons during the timed region (NVML, 5 ms period)."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {}
        for attr, label in (("nvmlClocksEventReasonHwSlowdown", "hw_slowdown"),
                            ("nvmlClocksThrottleReasonHwSlowdown", "hw_slowdown"),
                            ("nvmlClocksEventReasonHwThermalSlowdown", "hw_thermal_slowdown"),
                            ("nvmlClocksThrottleReasonHwThermalSlowdown", "hw_thermal_slowdown"),
                            ("nvmlClocksEventReasonSwThermalSlowdown", "sw_thermal_slowdown"),
                            ("nvmlClocksThrottleReasonSwThermalSlowdown", "sw_thermal_slowdown"),
                            ("nvmlClocksEventReasonSwPowerCap", "sw_power_cap"),
                            ("nvmlClocksThrottleReasonSwPowerCap", "sw_power_cap")):
            if hasattr(nv, attr):
                names[getattr(nv, attr)] = label
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons", None)
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                if get_reasons is not None:
                    mask = get_reasons(self.h)
                    for bit, label in names.items():
                        if mask & bit:
                            self.reasons.add(label)
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()

    def stop(self):
        if self._thread is not None:
            self._stop.set()
            self._thread.join()
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def dist_info():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return world, rank, local


# --------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline leg: the reference's CPU path (port: torch.topk + torch.sparse CSR mm)
# --------------------------------------------------------------------------------------------------
def cpu_leg(n, e, k, seconds_budget, steps, warmup):
    import cpu_baseline
    avg_deg = max(1, e // n)
    threads = os.cpu_count() or 1
    probe_rows = 512
    p = cpu_baseline.sample_problem(n, avg_deg, probe_rows, k)
    t_probe, _ = cpu_baseline.time_layer(p, 1, 0, threads)        # includes the full top-k
    # split the probe into the fixed top-k part and the per-row part with a second, larger probe
    p2 = cpu_baseline.sample_problem(n, avg_deg, 4 * probe_rows, k)
    t_probe2, _ = cpu_baseline.time_layer(p2, 1, 0, threads)
    per_row = max((t_probe2 - t_probe) / (3 * probe_rows), 1e-7)
    fixed = max(t_probe - per_row * probe_rows, 0.0)
    per_step = seconds_budget / max(steps + warmup, 1)
    rows = int(max(probe_rows, min(n, (per_step - fixed) / per_row)))
    p = cpu_baseline.sample_problem(n, avg_deg, rows, k)
    sec, threads = cpu_baseline.time_layer(p, steps, warmup, threads)
    gbs = p["bytes"] / sec / 1e9
    sample = "first %d of %d rows (%d of %d edges) of a uniform Reddit-shape graph, full top-k over all %d nodes; " \
             "torch.topk + scatter + torch.sparse_csr@dense fwd, A^T@g + gather bwd" % (rows, n, p["n_edges"], e, n)
    return {"value": gbs, "unit": "GB/s", "cores": threads, "kind": "port", "sample": sample,
            "ms_per_step": sec * 1e3, "rows": rows}


def run_reference(args, n, e):
    world, rank, _ = dist_info()
    if rank != 0:
        return
    leg = cpu_leg(n, e, args.k, seconds_budget=150.0, steps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": metric_name(args), "value": leg["value"], "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": leg["ms_per_step"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, n, e),
        "cpu_baseline": {k: leg[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": leg["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, n, e):
    return {"workload": "synthetic %s-shape graph (%d nodes, %d edges, uniform kind, seed 123, U[0,1) edge values), "
                        "hidden %d, k=%d: top-k + fwd SpGEMM + bwd SSpMM" % (args.shape, n, e, DIM, args.k),
            "shape": args.shape, "nodes": n, "edges": e, "hidden": DIM, "k": args.k,
            "l2_note": "inputs per step (CSR 917 MB + features 239 MB + gradient 239 MB) exceed the 126 MB L2; no flush needed",
            "parallelism": "1 GPU" if args.gpus == 1 else "1-D row sharding over %d GPUs (%s partition), all_gather(CBSR) fwd, %s bwd" % (args.gpus, "equal-row" if args.partition == "rows" else "equal-edge", args.bwd_mode)}


# --------------------------------------------------------------------------------------------------
# ours
# --------------------------------------------------------------------------------------------------
def run_ours(args, n, e):
    import torch.distributed as dist
    import maxk_cuda_kernels as K          # raises if libmaxk_b200.so is missing: no fallback
    from synth_graphs import synth_graph

    world, rank, local = dist_info()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    k = args.k
    graph = synth_graph(n, e, seed=123, kind="uniform", device=dev)
    gen = torch.Generator(device=dev).manual_seed(123)
    b_topk, b_fwd, b_bwd = layer_bytes(n, e, k)
    total_bytes = b_topk + b_fwd + b_bwd

    ev = lambda: torch.cuda.Event(enable_timing=True)
    nvml_index = local
    try:
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            nvml_index = int(vis.split(",")[local])
    except (ValueError, IndexError):
        nvml_index = local
    sampler = ClockSampler(nvml_index)

    if world == 1:
        ip = graph["indptr"]
        rb, re_, ix, va = ip[:-1], ip[1:], graph["indices"], graph["values"]
        x = torch.rand(n, DIM, device=dev, generator=gen)
        grad = torch.rand(n, DIM, device=dev, generator=gen)
        out = torch.empty(n, DIM, device=dev)
        gs = torch.empty(n, k, device=dev)

        def step(marks=None):
            r = K.topk_cbsr(x, k, order=K.ORDER_BANKED)
            if marks:
                marks[0].record()
            K.spgemm_forward_csr(rb, re_, ix, va, r["values"], r["sel"], out=out)
            if marks:
                marks[1].record()
            K.sspmm_backward_csr(rb, re_, ix, va, grad, r["sel"], out=gs)

        for _ in range(max(args.warmup, 3)):
            step()
        torch.cuda.synchronize()
        marks = [(ev(), ev(), ev(), ev()) for _ in range(args.steps)]
        sampler.start()
        torch.cuda.synchronize()
        for s in range(args.steps):
            marks[s][0].record()
            step(marks[s][1:3])
            marks[s][3].record()
        torch.cuda.synchronize()
        clocks = sampler.stop()
        t_step = sum(m[0].elapsed_time(m[3]) for m in marks) / args.steps
        t_topk = sum(m[0].elapsed_time(m[1]) for m in marks) / args.steps
        t_fwd = sum(m[1].elapsed_time(m[2]) for m in marks) / args.steps
        t_bwd = sum(m[2].elapsed_time(m[3]) for m in marks) / args.steps

        # ---- e2e: host buffers in, host buffers out, through the host-buffer operator API --------
        # (maxk_host_pipeline.HostStagedMaxKLayer: same kernels, copies overlapped with compute slab by slab;
        #  every step copies its inputs host->device and its results device->host inside the timed region)
        from maxk_host_pipeline import HostStagedMaxKLayer
        hx = torch.empty(n, DIM, pin_memory=True).copy_(x)
        hg = torch.empty(n, DIM, pin_memory=True).copy_(grad)
        hout = torch.empty(n, DIM, pin_memory=True)
        hgs = torch.empty(n, k, pin_memory=True)
        del x, grad, out, gs
        staged = HostStagedMaxKLayer(ip, ix, va, k, dim=DIM, slabs=8)
        for _ in range(3):
            staged.run(hx, hg, hout, hgs)
        torch.cuda.synchronize()
        e2e_steps = max(3, min(args.steps, 10))
        a, b = ev(), ev()
        a.record()
        for _ in range(e2e_steps):
            done = staged.run(hx, hg, hout, hgs, block_current_stream=False)
        torch.cuda.current_stream().wait_event(done)
        b.record()
        torch.cuda.synchronize()
        t_e2e = a.elapsed_time(b) / e2e_steps
        e2e_launches = staged.launches_per_call * e2e_steps
        h2d, d2h = 2 * n * DIM * 4, n * DIM * 4 + n * k * 4
        launches = KERNELS_PER_STEP * args.steps
        parts = {"topk_ms": t_topk, "fwd_ms": t_fwd, "bwd_ms": t_bwd, "e2e_kernel_launches": e2e_launches,
                 "e2e_note": "8 row slabs, h2d / compute / d2h on three streams, consecutive steps double-buffered"}
        roof_bytes, roof_ms = b_fwd, t_fwd
        scaling = "strong"
    else:
        from sharded import ShardedMaxKAggregation
        layer = ShardedMaxKAggregation(graph, k, backward_mode=args.bwd_mode, partition=args.partition)
        m = layer.m
        x = torch.rand(m, DIM, device=dev, generator=gen)
        grad = torch.rand(m, DIM, device=dev, generator=gen)
        del graph
        torch.cuda.empty_cache()

        def step():
            out_l = layer.forward(x)
            gs_l = layer.backward(grad)
            return out_l, gs_l

        for _ in range(max(args.warmup, 3)):
            step()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sampler.start()
        a, b = ev(), ev()
        a.record()
        for _ in range(args.steps):
            step()
        b.record()
        torch.cuda.synchronize()
        dist.barrier()
        clocks = sampler.stop()
        t = torch.tensor([a.elapsed_time(b) / args.steps], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_step = float(t.item())

        hx = torch.empty(m, DIM, pin_memory=True).copy_(x)
        hg = torch.empty(m, DIM, pin_memory=True).copy_(grad)
        hout = torch.empty(m, DIM, pin_memory=True)
        hgs = torch.empty(m, k, pin_memory=True)
        dx, dg = torch.empty_like(x), torch.empty_like(grad)

        def e2e_step():
            dx.copy_(hx, non_blocking=True)
            dg.copy_(hg, non_blocking=True)
            out_l = layer.forward(dx)
            gs_l = layer.backward(dg)
            hout.copy_(out_l, non_blocking=True)
            hgs.copy_(gs_l, non_blocking=True)

        for _ in range(2):
            e2e_step()
        torch.cuda.synchronize()
        dist.barrier()
        e2e_steps = max(3, min(args.steps, 10))
        a, b = ev(), ev()
        a.record()
        for _ in range(e2e_steps):
            e2e_step()
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / e2e_steps], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_e2e = float(t.item())
        h2d, d2h = world * 2 * m * DIM * 4, world * (m * DIM * 4 + m * k * 4)
        launches = KERNELS_PER_STEP * args.steps * world
        parts = {"wire_bytes_per_rank": layer.wire_bytes()}
        roof_bytes, roof_ms = None, None
        scaling = "strong"

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak_gbs()
    value = total_bytes / (t_step * 1e-3) / 1e9
    line = {
        "metric": metric_name(args), "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": t_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(args, n, e), "clocks": clocks,
        "e2e": {"value": total_bytes / (t_e2e * 1e-3) / 1e9, "unit": "GB/s", "ms_per_step": t_e2e,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": launches, "breakdown": parts, "algorithmic_bytes_per_step": total_bytes,
    }
    if roof_bytes is not None:
        achieved = roof_bytes / (roof_ms * 1e-3) / 1e9
        line["roofline"] = {"bound": "hbm", "kernel": "spgemm_fwd_kernel<%d>" % k, "achieved": achieved, "peak": peak,
                            "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                            "algorithmic_bytes_per_launch": roof_bytes, "ms_per_launch": roof_ms,
                            "traffic": ncu_traffic("%s_k%d" % (args.shape, k), "spgemm_fwd_kernel"),
                            "layer_frac": value / peak,
                            "note": "not HBM-limited: the kernel saturates the SM LSU data pipe "
                                    "(l1tex__data_pipe_lsu_wavefronts ~96% of peak), see profiles/ and DESIGN.md"}
    if world == 1 and not args.no_cpu_baseline:
        leg = cpu_leg(n, e, k, seconds_budget=20.0, steps=1, warmup=0)
        line["cpu_baseline"] = {kk: leg[kk] for kk in ("value", "unit", "cores", "kind", "sample")}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--shape", default="reddit")
    ap.add_argument("--k", type=int, default=32)
    ap.add_argument("--scale", type=float, default=1.0, help="developer knob: shrink the graph (not a contract bench)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--bwd-mode", default="reduce_scatter", choices=["reduce_scatter", "allgather", "overlap"],
                    help="multi-GPU backward exchange (sharded.py)")
    ap.add_argument("--partition", default="rows", choices=["rows", "nnz"],
                    help="multi-GPU row partition: equal row slabs, or equal edge counts (identical on the uniform contract graph)")
    args = ap.parse_args()
    from synth_graphs import SHAPES
    n, e = SHAPES[args.shape]
    n, e = int(n * args.scale), int(e * args.scale)
    if args.impl == "reference":
        run_reference(args, n, e)
    else:
        run_ours(args, n, e)


if __name__ == "__main__":
    main()
