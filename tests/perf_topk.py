"""Developer timing of the top-k -> CBSR kernel alone (not collected by pytest): rows/s and the fraction of the
HBM copy peak for the shapes of BASELINE.json, uniform and normal data, every order; index sets are checked
against torch.topk on the tie-free rows."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "spgemm-prunning_b200"))
import maxk_cuda_kernels as K  # noqa: E402


def timeit(fn, warm=3, reps=7):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


def main():
    peak = 6538.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    for name, n in (("reddit", 232965), ("yelp", 716847), ("products", 2449029)):
        for dist in ("uniform", "normal"):
            torch.manual_seed(123)
            x = torch.rand(n, 256, device="cuda") if dist == "uniform" else torch.randn(n, 256, device="cuda")
            for k in ((8, 16, 32, 64) if name == "reddit" else (32,)):
                r = K.topk_cbsr(x, k, order=2)
                ref = torch.topk(x[:20000], k, dim=1).indices.sort(dim=1).values
                got = r["sel"][:20000].long().sort(dim=1).values
                ok = bool((got == ref).all())
                t2 = timeit(lambda: K.topk_cbsr(x, k, order=2))
                t1 = timeit(lambda: K.topk_cbsr(x, k, order=1))
                t0 = timeit(lambda: K.topk_cbsr(x, k, order=0))
                gb = (n * 256 * 4 + n * k * 5) / 1e9
                print("%-8s %-7s k=%-2d banked %.3f ms (%.0f GB/s, %.0f%% of %.0f) | column %.3f | value %.3f | sets==torch.topk: %s"
                      % (name, dist, k, t2, gb / t2 * 1e3, 100 * gb / t2 * 1e3 / peak, peak, t1, t0, ok), flush=True)
            del x


if __name__ == "__main__":
    main()
