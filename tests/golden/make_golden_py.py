"""Generates tests/golden/ref_py.npz by running the REFERENCE's own Python code on the CPU.

Run in the build container only (needs /root/reference; never at test time):
    python tests/golden/make_golden_py.py

What is executed from /root/reference (nothing is copied into this repo):
  * maxk_spgemm_function.py  MaxKSpGEMMFunction.forward through maxk_spgemm(): with no
    `maxk_cuda_kernels` importable it takes its pure-PyTorch branch (:96-126): torch.topk ->
    scatter_ -> torch.sparse.mm -> / in_degrees.  (Its backward returns Nones, :182-184.)
  * utils/models.py MaxK (:11-25) and model_integrated_v3.py OPTMaxK (:28-43): the class sources
    are pulled out with `ast` (the modules import dgl, which is not installed) and executed.
  * kernels/generate_meta.py, run as the script it is, in a temp dir holding graphs/*.indptr.
  * graph_loader.py GraphDataLoader.load_graph (seed-123 edge values).
Inputs are seeded and tie-free (torch.topk's tie order is unspecified).
"""
import ast
import os
import subprocess
import sys
import tempfile

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "spgemm-prunning_b200"))
from synth_graphs import synth_graph  # noqa: E402  (only the graph generator; NOT our maxk_cuda_kernels)

sys.path.remove(os.path.join(ROOT, "spgemm-prunning_b200"))
sys.modules.pop("maxk_cuda_kernels", None)


def class_from(path, name):
    tree = ast.parse(open(path).read())
    node = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == name)
    ns = {"torch": torch, "Function": torch.autograd.Function}
    exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
    return ns[name]


def main():
    out = {}
    torch.manual_seed(0)
    n, e, k, d = 200, 3000, 32, 256
    g = synth_graph(n, e, seed=7, kind="powerlaw")
    indptr, indices, values = g["indptr"], g["indices"], g["values"]
    x = torch.rand(n, d)
    in_deg = torch.clamp((indptr[1:] - indptr[:-1]).float(), min=1)
    out.update(indptr=indptr.numpy(), indices=indices.numpy(), values=values.numpy(), x=x.numpy(),
               in_deg=in_deg.numpy(), k=np.int32(k))

    # --- MaxKSpGEMMFunction.forward, CPU branch --------------------------------------------------
    sys.path.insert(0, REF)
    import io
    import contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        import maxk_spgemm_function as ref_fn
        assert not ref_fn.MAXK_KERNELS_AVAILABLE
        y = ref_fn.maxk_spgemm(indices, values, x.clone(), k, None, 0, indptr, in_deg, None, None, None)
        y_raw = ref_fn.maxk_spgemm(indices, values, x.clone(), k, None, 0, indptr, None, None, None, None)
    out.update(spgemm_fwd_norm=y.numpy(), spgemm_fwd_raw=y_raw.numpy())

    # --- MaxK / OPTMaxK -----------------------------------------------------------------------------
    MaxK = class_from(os.path.join(REF, "utils", "models.py"), "MaxK")
    OPTMaxK = class_from(os.path.join(REF, "model_integrated_v3.py"), "OPTMaxK")
    xs = torch.randn(64, d, requires_grad=True)
    upstream = torch.randn(64, d)
    for kk in (8, 32):
        ym = MaxK.apply(xs, kk)
        (gm,) = torch.autograd.grad(ym, xs, upstream)
        out["maxk_x"] = xs.detach().numpy()
        out["maxk_up"] = upstream.numpy()
        out["maxk_fwd_k%d" % kk] = ym.detach().numpy()
        out["maxk_bwd_k%d" % kk] = gm.numpy()
        yo, tv, ti = OPTMaxK.apply(xs, kk)
        up_v = torch.randn(64, kk, generator=torch.Generator().manual_seed(kk))
        (go,) = torch.autograd.grad([yo, tv], xs, [upstream, up_v])
        out["optmaxk_fwd_k%d" % kk] = yo.detach().numpy()
        out["optmaxk_vals_k%d" % kk] = tv.detach().numpy()
        out["optmaxk_idx_k%d" % kk] = ti.numpy()
        out["optmaxk_upv_k%d" % kk] = up_v.numpy()
        out["optmaxk_bwd_k%d" % kk] = go.numpy()      # reference drops grad_topk_values

    # --- generate_meta.py + graph_loader.py ------------------------------------------------------------
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "graphs"))
        big = synth_graph(500, 40000, seed=3, kind="powerlaw")     # rows longer than 64 -> several quads per row
        for name, gr in (("small", g), ("big", big)):
            gr["indptr"].numpy().astype(np.int32).tofile(os.path.join(tmp, "graphs", name + ".indptr"))
            gr["indices"].numpy().astype(np.int32).tofile(os.path.join(tmp, "graphs", name + ".indices"))
        subprocess.run([sys.executable, os.path.join(REF, "kernels", "generate_meta.py")], cwd=tmp, check=True,
                       stdout=subprocess.DEVNULL)
        out["warp4_small"] = np.fromfile(os.path.join(tmp, "w12_nz64_warp_4", "small.warp4"), dtype=np.int32)
        out["warp4_big"] = np.fromfile(os.path.join(tmp, "w12_nz64_warp_4", "big.warp4"), dtype=np.int32)
        out["big_indptr"] = big["indptr"].numpy()
        with contextlib.redirect_stdout(io.StringIO()):
            import graph_loader as ref_loader
            loaded = ref_loader.GraphDataLoader(os.path.join(tmp, "graphs") + "/").load_graph("small")
        out["loader_values"] = loaded["values"]
        assert np.array_equal(loaded["indptr"], g["indptr"].numpy())

    np.savez_compressed(os.path.join(HERE, "ref_py.npz"), **out)
    print("wrote ref_py.npz:", {k_: v.shape for k_, v in out.items()})


if __name__ == "__main__":
    main()
