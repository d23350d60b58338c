"""graph_loader -- same surface as the reference's graph_loader.py (GraphDataLoader :13-157).

Reads / writes the reference's on-disk graph format: `<name>.indptr` and `<name>.indices`, raw
little-endian int32 arrays (graph_loader.py:19-68, kernels/data.h:8-24), and `<name>.warp4`
quads (kernels/generate_meta.py:46-48).  Edge values are not stored; like the reference they
are U[0,1) from numpy seed 123 (graph_loader.py:71-72).
"""
import os
from pathlib import Path

import numpy as np
import torch


class GraphDataLoader:
    def __init__(self, base_dir="kernels/graphs/"):
        self.base_dir = base_dir

    def read_binary_array(self, filepath, dtype=np.int32):
        if not os.path.exists(filepath):
            raise FileNotFoundError("Graph file not found: %s" % filepath)
        if dtype not in (np.int32, np.float32):
            raise ValueError("Unsupported dtype: %s" % dtype)
        return np.fromfile(filepath, dtype=dtype)

    def load_graph(self, graph_name):
        stem = Path(graph_name).stem
        indptr = self.read_binary_array(os.path.join(self.base_dir, stem + ".indptr"), np.int32)
        indices = self.read_binary_array(os.path.join(self.base_dir, stem + ".indices"), np.int32)
        v_num, e_num = len(indptr) - 1, len(indices)
        np.random.seed(123)
        values = np.random.uniform(0, 1, e_num).astype(np.float32)
        return {"graph_name": stem, "indptr": indptr, "indices": indices, "values": values,
                "v_num": v_num, "e_num": e_num}

    def save_graph(self, graph_name, indptr, indices):
        """Inverse of load_graph: writes <name>.indptr / <name>.indices (dataset_gen.py:100-118)."""
        os.makedirs(self.base_dir, exist_ok=True)
        stem = Path(graph_name).stem
        _np(indptr).astype(np.int32).tofile(os.path.join(self.base_dir, stem + ".indptr"))
        _np(indices).astype(np.int32).tofile(os.path.join(self.base_dir, stem + ".indices"))

    def to_cuda_tensors(self, graph_data, device="cuda"):
        out = {}
        for key, value in graph_data.items():
            out[key] = torch.from_numpy(np.ascontiguousarray(value)).to(device) if isinstance(value, np.ndarray) else value
        return out

    def generate_test_features(self, v_num, dim_origin=256, dim_k_limit=64, device="cuda"):
        """graph_loader.py:101-141 without the per-row Python loop."""
        torch.manual_seed(123)
        vin_sparse = torch.rand(v_num, dim_origin, device=device, dtype=torch.float32)
        vin_sparse_data = torch.rand(v_num, dim_k_limit, device=device, dtype=torch.float32)
        sel = torch.rand(v_num, dim_origin, device=device).argsort(dim=1)[:, :dim_k_limit].to(torch.uint8)
        return {
            "vin_sparse": vin_sparse, "vin_sparse_data": vin_sparse_data, "vin_sparse_selector": sel.contiguous(),
            "vout_ref": torch.zeros(v_num, dim_origin, device=device),
            "vout_maxk": torch.zeros(v_num, dim_origin, device=device),
            "vout_maxk_backward": torch.zeros(v_num, dim_k_limit, device=device),
            "dim_origin": dim_origin, "dim_k_limit": dim_k_limit,
        }

    def get_available_graphs(self):
        if not os.path.exists(self.base_dir):
            return []
        names = [f[:-7] for f in os.listdir(self.base_dir) if f.endswith(".indptr")]
        return sorted(n for n in names if os.path.exists(os.path.join(self.base_dir, n + ".indices")))


def clean_edges(src, dst, num_nodes, undirected=True, self_loops=True):
    """dataset_gen.py:44-98 as tensor ops on the edge list's device: add reverse edges, add one self loop per
    node, drop duplicate (src, dst) pairs, return CSR (indptr int32[N+1], indices int32[E]) with src as the row.
    The reference dedups with a Python set over all edges (dataset_gen.py:72-80: minutes on Reddit); here it is
    one sort of 64-bit keys.  Columns come out ascending inside a row (DGL keeps insertion order; the operator
    does not depend on it beyond the fp32 summation order)."""
    src = torch.as_tensor(src).long().reshape(-1)
    dst = torch.as_tensor(dst).long().reshape(-1).to(src.device)
    n = int(num_nodes)
    if src.numel() != dst.numel():
        raise ValueError("src and dst must have the same length")
    if src.numel() and (int(torch.minimum(src.min(), dst.min())) < 0 or int(torch.maximum(src.max(), dst.max())) >= n):
        raise ValueError("edge endpoint outside [0, num_nodes)")
    parts = [src * n + dst]
    if undirected:
        parts.append(dst * n + src)
    if self_loops:
        loops = torch.arange(n, device=src.device)
        parts.append(loops * n + loops)
    key = torch.unique(torch.cat(parts))          # sorted: row-major, columns ascending
    rows, cols = key // n, key % n
    indptr = torch.zeros(n + 1, dtype=torch.int64, device=src.device)
    indptr[1:] = torch.cumsum(torch.bincount(rows, minlength=n), 0)
    if key.numel() >= 2 ** 31:
        raise ValueError("more than 2^31-1 edges do not fit the int32 graph format")
    return indptr.to(torch.int32), cols.to(torch.int32)


def csr_to_csc(indptr, indices, values=None):
    """Transpose of a square CSR matrix (what generate_meta_csc.py:134-141 asks scipy for): returns
    (csc_indptr, csc_indices[, csc_values]); row ids ascending inside a column (stable sort), same device."""
    indptr = torch.as_tensor(indptr).long()
    indices = torch.as_tensor(indices).long().to(indptr.device)
    n = indptr.numel() - 1
    rows = torch.repeat_interleave(torch.arange(n, device=indptr.device), indptr[1:] - indptr[:-1])
    order = torch.sort(indices, stable=True).indices
    t_ptr = torch.zeros(n + 1, dtype=torch.int64, device=indptr.device)
    t_ptr[1:] = torch.cumsum(torch.bincount(indices, minlength=n), 0)
    out = (t_ptr.to(torch.int32), rows[order].to(torch.int32))
    if values is not None:
        out += (torch.as_tensor(values).to(indptr.device)[order],)
    return out


def generate_meta(graph_name, base_dir="kernels/graphs/", root="kernels", num_warps=12, warp_max_nz=64,
                  device="cuda"):
    """kernels/generate_meta.py + generate_meta_csc.py:97-173 for one graph: reads <name>.indptr/.indices,
    builds the CSR and the CSC warp4 quads with the GPU scan/fill kernels (maxk_warp4_scan / maxk_warp4_fill)
    and writes <root>/w12_nz64_warp_4/<name>.warp4 and <root>/w12_nz64_warp_4_csc/<name>.warp4_csc.
    Returns (csr_path, n_csr_quads, csc_path, n_csc_quads)."""
    import maxk_cuda_kernels as mk
    g = GraphDataLoader(base_dir).load_graph(graph_name)
    indptr = torch.from_numpy(g["indptr"]).to(device)
    indices = torch.from_numpy(g["indices"]).to(device)
    w_csr, n_csr = mk.build_warp4(indptr, warp_max_nz)
    t_ptr, _ = csr_to_csc(indptr, indices)
    w_csc, n_csc = mk.build_warp4(t_ptr, warp_max_nz)
    p_csr = warp4_path(g["graph_name"], num_warps, warp_max_nz, False, root)
    p_csc = warp4_path(g["graph_name"], num_warps, warp_max_nz, True, root)
    save_warp4(p_csr, w_csr)
    save_warp4(p_csc, w_csc)
    return p_csr, n_csr, p_csc, n_csc


def _np(a):
    return a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)


def warp4_path(graph_name, num_warps=12, warp_max_nz=64, csc=False, root="kernels"):
    """Path convention of load_warp4_metadata (cuda_kernel_bindings.cpp:290-293, binding_v2.py:323-326)."""
    d = "w%d_nz%d_warp_4%s" % (num_warps, warp_max_nz, "_csc" if csc else "")
    return os.path.join(root, d, graph_name + (".warp4_csc" if csc else ".warp4"))


def save_warp4(path, warp4):
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    _np(warp4).astype(np.int32).tofile(path)
