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
KERNELS_PER_STEP = 4      # topk_banked, spgemm_fwd_slots, sspmm_bwd + its long-row kernel
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
    """Figures of the committed ncu --set full capture of `workload` (dram read+write bytes per launch, LSU
    wavefronts per edge ...): one kernel's traffic, or the whole record when kernel is None."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            rec = json.load(f).get(workload, {})
            return rec if kernel is None else rec.get(kernel)
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
    return {"workload": "synthetic %s-shape graph (%d nodes, %d edges, %s kind, seed 123, U[0,1) edge values), "
                        "hidden %d, k=%d: top-k + fwd SpGEMM + bwd SSpMM" % (args.shape, n, e, args.kind, DIM, args.k),
            "shape": args.shape, "nodes": n, "edges": e, "hidden": DIM, "k": args.k,
            "l2_note": "inputs per step (CSR 917 MB + features 239 MB + gradient 239 MB) exceed the 126 MB L2; no flush needed",
            "parallelism": "1 GPU" if args.gpus == 1 else "1-D row sharding over %d GPUs (%s partition), CBSR slabs exchanged fwd (top-k writes into peer memory, else NCCL all_gather), %s bwd" % (args.gpus, "equal-row" if args.partition == "rows" else "equal-edge", args.bwd_mode)}


# --------------------------------------------------------------------------------------------------
# ours
# --------------------------------------------------------------------------------------------------
SWEEP = [("reddit", "uniform", 8), ("reddit", "uniform", 16), ("reddit", "uniform", 64), ("reddit", "powerlaw", 32),
         ("flickr", "uniform", 32), ("yelp", "uniform", 32), ("proteins", "uniform", 64), ("products", "uniform", 32)]
L2_GATHER_PEAK_TBS = 18.5      # tools/l2_bw.cu on B200: random 160-byte CBSR rows out of L2, 16-byte-per-lane loads
L2_RED_PEAK_TBS = 6.2          # tools/l2_bw.cu: 128-byte red.global.add rows into an L2-resident target


def _ev():
    return torch.cuda.Event(enable_timing=True)


def _median_ms(fn, warm=2, reps=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    evs = [(_ev(), _ev()) for _ in range(reps)]
    for a, b in evs:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2]


def parity_block(K, ip, ix, va, x, grad, vals, sel, out, gs, k, seed, rows_out=None, rows_gs=None, row_offset=0,
                 n_samples=1024):
    """Sampled-row check of one step's results against the C oracle (oracle/sampled_parity.py), outside the
    timed region.  (ip, ix, va, x, grad, vals, sel) describe the WHOLE problem in global numbering; out / gs are
    this process's result rows [row_offset, row_offset + rows)."""
    import sampled_parity as sp
    dev = ix.device
    n_local = out.size(0) if rows_out is None else rows_out
    idx = sp.sample_ids(n_local, n_samples, seed, dev)
    sets_equal, vals_equal = sp.check_topk(x[idx + row_offset], vals[idx + row_offset], sel[idx + row_offset], k)
    fv, fr = sp.check_forward(ip, ix, va, vals, sel, idx + row_offset, out[idx])
    n_gs = gs.size(0) if rows_gs is None else rows_gs
    idx2 = sp.sample_ids(n_gs, n_samples, seed + 1, dev)
    bv, br, used = sp.check_backward(ip, ix, va, grad, sel, idx2 + row_offset, gs[idx2])
    return {"topk_sets_equal": bool(sets_equal and vals_equal), "fwd_max_rel": fr, "bwd_max_rel": br,
            "fwd_violation": fv, "bwd_violation": bv, "rows_checked": int(idx.numel()), "dst_rows_checked": used,
            "tolerance": "index sets bit-exact; |got-exp| <= 1e-6 + 1e-5|exp| (violation <= 1)",
            "ok": bool(sets_equal and vals_equal and fv <= 1.0 and bv <= 1.0)}


def sweep_group(K, shape, kind, ks, dev, peak, ref):
    """Secondary configurations on one graph (BASELINE.json configs 1, 3, 4, 5 and the k sweep of config 2): our
    three kernels, the reference's kernels recompiled for sm_100a on the same inputs (checker leg, oracle/_ref),
    and a sampled parity check."""
    from synth_graphs import SHAPES, synth_graph
    n, e = SHAPES[shape]
    g = synth_graph(n, e, seed=123, kind=kind, device=dev)
    ip, ix, va = g["indptr"], g["indices"], g["values"]
    rb, re_ = ip[:-1], ip[1:]
    gen = torch.Generator(device=dev).manual_seed(123)
    x = torch.rand(n, DIM, device=dev, generator=gen)
    grad = torch.rand(n, DIM, device=dev, generator=gen)
    plan = K.build_plan(rb, re_)
    out = torch.empty(n, DIM, device=dev)
    max_deg = int((re_ - rb).max())
    w4 = None
    res = []
    for k in ks:
        gs = torch.empty(n, k, device=dev)
        r = K.topk_cbsr(x, k, order=K.ORDER_BANKED)
        t_topk = _median_ms(lambda: K.topk_cbsr(x, k, order=K.ORDER_BANKED, out_values=r["values"], out_sel=r["sel"]))
        t_fwd = _median_ms(lambda: K.spgemm_forward_csr(rb, re_, ix, va, r["values"], r["sel"], out=out, plan=plan))
        t_bwd = _median_ms(lambda: K.sspmm_backward_csr(rb, re_, ix, va, grad, r["sel"], out=gs))
        b_topk, b_fwd, b_bwd = layer_bytes(n, e, k)
        frac = lambda byt, ms: byt / (ms * 1e-3) / 1e9 / peak
        ent = {"shape": shape, "kind": kind, "k": k, "nodes": n, "edges": e, "max_degree": max_deg,
               "topk_ms": t_topk, "fwd_ms": t_fwd, "bwd_ms": t_bwd, "layer_ms": t_topk + t_fwd + t_bwd,
               "topk_frac": frac(b_topk, t_topk), "fwd_frac": frac(b_fwd, t_fwd), "bwd_frac": frac(b_bwd, t_bwd),
               "layer_frac": frac(b_topk + b_fwd + b_bwd, t_topk + t_fwd + t_bwd)}
        try:
            ent["parity"] = parity_block(K, ip, ix, va, x, grad, r["values"], r["sel"], out, gs, k, seed=7, n_samples=256)
        except Exception as ex:   # the checker must not take the measurement down
            ent["parity"] = {"ok": False, "error": repr(ex)[:200]}
        if ref is not None:
            try:
                if w4 is None:
                    w4 = K.build_warp4(ip)
                ent["ref_fwd_ms"] = _median_ms(lambda: ref.ref_cuda_forward(w4[0], ix, va, r["values"], r["sel"], w4[1]), 1, 3)
                ent["ref_bwd_ms"] = _median_ms(lambda: ref.ref_cuda_backward(w4[0], ix, va, grad, r["sel"], w4[1]), 1, 3)
            except Exception as ex:
                ent["ref_error"] = repr(ex)[:200]
        res.append(ent)
        del gs, r
    return res


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
    graph = synth_graph(n, e, seed=123, kind=args.kind, device=dev)
    gen = torch.Generator(device=dev).manual_seed(123)
    b_topk, b_fwd, b_bwd = layer_bytes(n, e, k)
    total_bytes = b_topk + b_fwd + b_bwd

    nvml_index = local
    try:
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            nvml_index = int(vis.split(",")[local])
    except (ValueError, IndexError):
        nvml_index = local
    sampler = ClockSampler(nvml_index)
    parity = None
    sweep = None
    from maxk_host_pipeline import bind_host_to_gpu
    host_cores = bind_host_to_gpu(nvml_index) if world > 1 else 0     # pinned buffers land on the GPU's NUMA node

    if world == 1:
        ip = graph["indptr"]
        rb, re_, ix, va = ip[:-1], ip[1:], graph["indices"], graph["values"]
        plan = K.build_plan(rb, re_)                       # once per graph, like the reference's warp4 metadata
        x = torch.rand(n, DIM, device=dev, generator=gen)
        grad = torch.rand(n, DIM, device=dev, generator=gen)
        out = torch.empty(n, DIM, device=dev)
        gs = torch.empty(n, k, device=dev)
        cb = {"values": torch.empty(n, k, device=dev), "sel": torch.empty(n, k, dtype=torch.uint8, device=dev)}

        def step(marks=None):
            K.topk_cbsr(x, k, order=K.ORDER_BANKED, out_values=cb["values"], out_sel=cb["sel"])
            if marks:
                marks[0].record()
            K.spgemm_forward_csr(rb, re_, ix, va, cb["values"], cb["sel"], out=out, plan=plan)
            if marks:
                marks[1].record()
            K.sspmm_backward_csr(rb, re_, ix, va, grad, cb["sel"], out=gs)

        for _ in range(max(args.warmup, 3)):
            step()
        torch.cuda.synchronize()
        marks = [(_ev(), _ev(), _ev(), _ev()) for _ in range(args.steps)]
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

        if not args.no_parity:
            parity = parity_block(K, ip, ix, va, x, grad, cb["values"], cb["sel"], out, gs, k, seed=11)

        # ---- e2e: host buffers in, host buffers out, through the host-buffer operator API --------
        # (maxk_host_pipeline.HostStagedMaxKLayer: same kernels, copies overlapped with compute slab by slab;
        #  every step copies its inputs host->device and its results device->host inside the timed region)
        from maxk_host_pipeline import HostStagedMaxKLayer
        hx = torch.empty(n, DIM, pin_memory=True).copy_(x)
        hg = torch.empty(n, DIM, pin_memory=True).copy_(grad)
        hout = torch.empty(n, DIM, pin_memory=True)
        hgs = torch.empty(n, k, pin_memory=True)
        ref_out_rows = out[:4096].cpu()
        del x, grad, out, gs
        staged = HostStagedMaxKLayer(ip, ix, va, k, dim=DIM, slabs=8)
        for _ in range(3):
            staged.run(hx, hg, hout, hgs)
        torch.cuda.synchronize()
        e2e_steps = max(3, min(args.steps, 10))
        a, b = _ev(), _ev()
        a.record()
        for _ in range(e2e_steps):
            done = staged.run(hx, hg, hout, hgs, block_current_stream=False)
        torch.cuda.current_stream().wait_event(done)
        b.record()
        torch.cuda.synchronize()
        t_e2e = a.elapsed_time(b) / e2e_steps
        if parity is not None:       # the host-buffer path must deliver the same rows as the device-resident one
            parity["e2e_matches_device_path"] = bool(torch.allclose(hout[:4096], ref_out_rows, rtol=1e-5, atol=1e-6))
            parity["ok"] = bool(parity["ok"] and parity["e2e_matches_device_path"])
        e2e_launches = staged.launches_per_call * e2e_steps
        h2d, d2h = 2 * n * DIM * 4, n * DIM * 4 + n * k * 4
        launches = KERNELS_PER_STEP * args.steps
        parts = {"topk_ms": t_topk, "fwd_ms": t_fwd, "bwd_ms": t_bwd, "e2e_kernel_launches": e2e_launches,
                 "e2e_note": "8 row slabs, h2d / compute / d2h on three streams, consecutive steps double-buffered"}
        roof_bytes, roof_ms = b_fwd, t_fwd
        scaling = "strong"
        del staged, hx, hg, hout, hgs, graph, ip, ix, va, rb, re_, plan
        torch.cuda.empty_cache()
        if not args.no_sweep and args.shape == "reddit" and args.k == 32 and args.scale == 1.0:
            peak0, _ = measured_peak_gbs()
            ref = None
            try:
                import oracle
                if oracle.ref_cuda_available():
                    ref = oracle
            except Exception:
                ref = None
            sweep = []
            groups = []
            for shape, kind, kk in SWEEP:
                if groups and groups[-1][0] == (shape, kind):
                    groups[-1][1].append(kk)
                else:
                    groups.append(((shape, kind), [kk]))
            for (shape, kind), kks in groups:
                try:
                    sweep.extend(sweep_group(K, shape, kind, kks, dev, peak0, ref))
                except Exception as ex:
                    sweep.append({"shape": shape, "kind": kind, "k": kks, "error": repr(ex)[:200]})
                torch.cuda.empty_cache()
    else:
        from sharded import ShardedMaxKAggregation, _all_gather, padded_position
        layer = ShardedMaxKAggregation(graph, k, backward_mode=args.bwd_mode, partition=args.partition, gather=args.gather)
        m = layer.m
        x = torch.rand(m, DIM, device=dev, generator=torch.Generator(device=dev).manual_seed(123 + rank))
        grad = torch.rand(m, DIM, device=dev, generator=torch.Generator(device=dev).manual_seed(1123 + rank))

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
        a, b = _ev(), _ev()
        a.record()
        for _ in range(args.steps):
            out_l, gs_l = step()
        b.record()
        torch.cuda.synchronize()
        dist.barrier()
        clocks = sampler.stop()
        t = torch.tensor([a.elapsed_time(b) / args.steps], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_step = float(t.item())

        # phase split of the step (outside the timed region; max over ranks): top-k + exchange, forward kernel,
        # backward (kernel + exchange), backward kernel alone
        phases = None
        if args.phases:
            marks = [[_ev() for _ in range(5)] for _ in range(10)]
            ipl = layer.rows["indptr"]
            scratch = torch.empty(layer.world * m, k, device=dev)
            for mk in marks:
                mk[0].record()
                vf, sf, _ = layer.gather_cbsr(x)
                layer.sel_full = sf
                mk[1].record()
                layer.compute.spgemm(layer.rows, vf, sf, layer.row_div)
                mk[2].record()
                layer.backward(grad)
                mk[3].record()
                K.sspmm_backward_csr(ipl[:-1], ipl[1:], layer.rows["indices"], layer.rows["values"], grad, sf, out=scratch)
                mk[4].record()
            torch.cuda.synchronize()
            ph = torch.tensor([sum(mk[i].elapsed_time(mk[i + 1]) for mk in marks) / len(marks) for i in range(4)], device=dev)
            dist.all_reduce(ph, op=dist.ReduceOp.MAX)
            phases = dict(zip(("topk_and_exchange_ms", "fwd_kernel_ms", "bwd_total_ms", "bwd_kernel_ms"), ph.tolist()))
            del scratch

        if not args.no_parity:
            # every rank checks sampled rows of ITS output slabs against the oracle on the whole problem
            # (features, gradients and CBSR all-gathered outside the timed region; the full graph is still resident)
            pos = padded_position(torch.arange(n, device=dev), layer.bounds, m)
            x_g = _all_gather(x, None)[pos]
            g_g = _all_gather(grad, None)[pos]
            vals_l, sel_l = layer.compute.topk(x, k)
            vals_g = _all_gather(vals_l, None)[pos]
            sel_g = _all_gather(sel_l, None)[pos]
            valid = layer.valid_rows()
            p = parity_block(K, graph["indptr"], graph["indices"], graph["values"], x_g, g_g, vals_g, sel_g, out_l, gs_l, k,
                             seed=11 + rank, rows_out=valid, rows_gs=valid, row_offset=layer.rows["row_lo"]) if valid > 0 else \
                {"topk_sets_equal": True, "fwd_violation": 0.0, "bwd_violation": 0.0, "fwd_max_rel": 0.0, "bwd_max_rel": 0.0,
                 "rows_checked": 0, "dst_rows_checked": 0, "ok": True}
            red = torch.tensor([p["fwd_violation"], p["bwd_violation"], p["fwd_max_rel"], p["bwd_max_rel"],
                                0.0 if p["topk_sets_equal"] else 1.0, 0.0 if p["ok"] else 1.0], device=dev, dtype=torch.float64)
            dist.all_reduce(red, op=dist.ReduceOp.MAX)
            cnt = torch.tensor([p["rows_checked"], p["dst_rows_checked"]], device=dev, dtype=torch.int64)
            dist.all_reduce(cnt)
            red = red.tolist()
            parity = {"topk_sets_equal": red[4] == 0.0, "fwd_max_rel": red[2], "bwd_max_rel": red[3], "fwd_violation": red[0],
                      "bwd_violation": red[1], "rows_checked": int(cnt[0]), "dst_rows_checked": int(cnt[1]), "ranks": world,
                      "tolerance": "index sets bit-exact; |got-exp| <= 1e-6 + 1e-5|exp| (violation <= 1), max over ranks",
                      "ok": red[5] == 0.0}
            del x_g, g_g, vals_g, sel_g, pos
        # the backward variant BASELINE.json names ("gradient rows exchanged with an NCCL allgather before local
        # aggregation"), timed beside the default (reduce of the compact partial): same forward, same graph
        ag_ms = None
        if not args.no_allgather_variant and args.bwd_mode != "allgather":
            layer_ag = ShardedMaxKAggregation(graph, k, backward_mode="allgather", partition=args.partition, gather=args.gather)
            for _ in range(3):
                layer_ag.forward(x)
                layer_ag.backward(grad)
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()
            a, b = _ev(), _ev()
            a.record()
            for _ in range(10):
                layer_ag.forward(x)
                layer_ag.backward(grad)
            b.record()
            torch.cuda.synchronize()
            t = torch.tensor([a.elapsed_time(b) / 10], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ag_ms = float(t.item())
            del layer_ag
        del graph, out_l, gs_l
        torch.cuda.empty_cache()

        # ---- e2e: every rank streams its slab through its own PCIe link (three-stream chunk pipeline) ----
        from maxk_host_pipeline import ShardedHostStagedLayer
        hx = torch.empty(m, DIM, pin_memory=True).copy_(x)
        hg = torch.empty(m, DIM, pin_memory=True).copy_(grad)
        hout = torch.empty(m, DIM, pin_memory=True)
        hgs = torch.empty(m, k, pin_memory=True)
        del x, grad
        staged = ShardedHostStagedLayer(layer, dim=DIM, slabs=4)
        for _ in range(3):
            staged.run(hx, hg, hout, hgs)
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        e2e_steps = max(3, min(args.steps, 10))
        a, b = _ev(), _ev()
        a.record()
        for _ in range(e2e_steps):
            done = staged.run(hx, hg, hout, hgs, block_current_stream=False)
        torch.cuda.current_stream().wait_event(done)
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / e2e_steps], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_e2e = float(t.item())
        h2d, d2h = world * 2 * m * DIM * 4, world * (m * DIM * 4 + m * k * 4)
        launches = KERNELS_PER_STEP * args.steps * world
        parts = {"wire_bytes_per_rank": layer.wire_bytes(), "forward_exchange": layer.gather, "backward_exchange": layer.reduce,
                 "forward_exchange_fallback_reason": layer.gather_error,
                 "e2e_note": "per rank: 4 row chunks, h2d / compute+NCCL / d2h on three streams, steps double-buffered; "
                             "process bound to the %d cores next to its GPU" % host_cores}
        if phases:
            parts["phases"] = phases
        if ag_ms is not None:
            parts["allgather_backward_variant_ms_per_step"] = ag_ms
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
    if parity is not None:
        line["parity"] = parity
    if roof_bytes is not None:
        achieved = roof_bytes / (roof_ms * 1e-3) / 1e9
        prof = ncu_traffic("%s_k%d" % (args.shape, k), None) or {}
        gathered = e * (8 + 5 * k) + n * DIM * 4
        line["roofline"] = {"bound": "hbm", "kernel": "spgemm_fwd_slots_kernel<%d, 2>" % k, "achieved": achieved, "peak": peak,
                            "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                            "algorithmic_bytes_per_launch": roof_bytes, "ms_per_launch": roof_ms,
                            "traffic": prof.get("spgemm_fwd_kernel"),
                            "layer_frac": value / peak,
                            "note": "not HBM-limited: the kernel saturates the SM LSU data pipe "
                                    "(l1tex__data_pipe_lsu_wavefronts %s%% of peak in the committed ncu capture), see "
                                    "roofline.l2, profiles/ and DESIGN.md 3.1-3.2" % prof.get("fwd_lsu_pct_of_peak", "~96"),
                            "l2": {"what": "the roofline this kernel is actually on: every edge gathers a 160-byte CBSR row "
                                           "out of L2 and scatters 32 products into shared memory",
                                   "gathered_bytes_per_launch": gathered,
                                   "achieved_tbs": gathered / (roof_ms * 1e-3) / 1e12,
                                   "peak_gather_tbs": L2_GATHER_PEAK_TBS, "peak_red_add_tbs": L2_RED_PEAK_TBS,
                                   "frac_of_gather_peak": gathered / (roof_ms * 1e-3) / 1e12 / L2_GATHER_PEAK_TBS,
                                   "lsu_wavefronts_per_edge": prof.get("fwd_lsu_wavefronts_per_edge"),
                                   "smem_conflict_wavefronts_per_edge": prof.get("fwd_smem_conflicts_per_edge"),
                                   "lsu_pipe_pct_of_peak": prof.get("fwd_lsu_pct_of_peak"),
                                   "peaks_source": "tools/l2_bw.cu (profiles/r01_microbench_l2.txt); ncu figures from the "
                                                   "committed capture named in profiles/roofline_traffic.json"}}
    if sweep is not None:
        line["extra"] = {"sweep": sweep,
                         "sweep_note": "per-kernel ms (median of 5, CUDA events) and fraction of the measured HBM peak on "
                                       "SURVEY 8(d) algorithmic bytes; ref_* = the reference's kernels on the same inputs and GPU"}
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
    ap.add_argument("--no-parity", action="store_true", help="skip the sampled-row oracle check after the timed region")
    ap.add_argument("--no-sweep", action="store_true", help="skip the secondary configurations (extra.sweep)")
    ap.add_argument("--no-allgather-variant", action="store_true",
                    help="multi-GPU: do not also time the all_gather(gradient rows) backward variant")
    ap.add_argument("--phases", action="store_true", help="multi-GPU: add a per-phase split of the step to breakdown")
    ap.add_argument("--kind", default="uniform", choices=["uniform", "powerlaw"], help="degree distribution of the synthetic graph")
    ap.add_argument("--gather", default="auto", choices=["auto", "peer", "nccl"],
                    help="multi-GPU forward exchange: top-k writing into peer memory, or NCCL all_gather")
    ap.add_argument("--bwd-mode", default="reduce_scatter", choices=["reduce_scatter", "allgather"],
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
