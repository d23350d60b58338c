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


def _np(a):
    return a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)


def warp4_path(graph_name, num_warps=12, warp_max_nz=64, csc=False, root="kernels"):
    """Path convention of load_warp4_metadata (cuda_kernel_bindings.cpp:290-293, binding_v2.py:323-326)."""
    d = "w%d_nz%d_warp_4%s" % (num_warps, warp_max_nz, "_csc" if csc else "")
    return os.path.join(root, d, graph_name + (".warp4_csc" if csc else ".warp4"))


def save_warp4(path, warp4):
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    _np(warp4).astype(np.int32).tofile(path)
