"""oracle.py -- numpy-facing wrapper of the CPU oracle (maxk_oracle.c) + CPU baseline.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product path (spgemm-prunning_b200/) never imports it.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libmaxk_oracle.so")
_REF_SO = os.path.join(_HERE, "_ref", "libmaxk_ref.so")


def build(verbose=False):
    """Compiles maxk_oracle.c (gcc) and, when /root/reference is present, oracle/_ref (nvcc)."""
    res = subprocess.run(["make", "-C", _HERE, "all"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
    if res.returncode != 0:
        raise RuntimeError("building the oracle failed")
    return _SO


def _lib():
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "maxk_oracle.c")):
        build()
    lib = ctypes.CDLL(_SO)
    lib.oracle_warp4.restype = ctypes.c_int64
    lib.oracle_num_threads.restype = ctypes.c_int
    return lib


_L = None


def lib():
    global _L
    if _L is None:
        _L = _lib()
    return _L


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else ctypes.c_void_p(0)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def num_threads():
    return int(lib().oracle_num_threads())


def banked_modulus(k):
    return 4 if k in (8, 16, 32, 64, 96, 128) else 1


def topk(x, k, order=0):
    """(values fp32 [N,k], columns int32 [N,k]); order 0 = (value desc, col asc), 1 = column asc,
    2 = MAXK_ORDER_BANKED (classes col mod 4 by size, see maxk_oracle.c)."""
    x = _f32(x)
    n, d = x.shape
    vals = np.empty((n, k), np.float32)
    cols = np.empty((n, k), np.int32)
    lib().oracle_topk(_p(x), ctypes.c_int64(n), ctypes.c_int(d), ctypes.c_int(k), _p(vals), _p(cols), ctypes.c_int(order),
                      ctypes.c_int(banked_modulus(k)))
    return vals, cols


def warp4(indptr, max_nz=64):
    indptr = _i32(indptr)
    n = indptr.size - 1
    w = lib().oracle_warp4(_p(indptr), ctypes.c_int64(n), ctypes.c_int(max_nz), ctypes.c_void_p(0))
    out = np.empty(4 * w, np.int32)
    lib().oracle_warp4(_p(indptr), ctypes.c_int64(n), ctypes.c_int(max_nz), _p(out))
    return out, int(w)


def spgemm_fwd(indptr, indices, values, data, sel, dim=256, deg=None):
    indptr, indices, values, data, sel = _i32(indptr), _i32(indices), _f32(values), _f32(data), _u8(sel)
    n_rows, k = indptr.size - 1, data.shape[1]
    out = np.empty((n_rows, dim), np.float32)
    deg = _f32(deg) if deg is not None else None
    lib().oracle_spgemm_fwd(_p(indptr), _p(indices), _p(values), _p(data), _p(sel), ctypes.c_int64(n_rows),
                            ctypes.c_int(k), ctypes.c_int(dim), _p(deg), _p(out))
    return out


def _u16(a):
    return np.ascontiguousarray(a, dtype=np.uint16)


def spgemm_fwd16(indptr, indices, values, data, sel16, dim, deg=None):
    """Forward with uint16 selectors (feature widths above 256)."""
    indptr, indices, values, data, sel16 = _i32(indptr), _i32(indices), _f32(values), _f32(data), _u16(sel16)
    n_rows, k = indptr.size - 1, data.shape[1]
    out = np.empty((n_rows, dim), np.float32)
    deg = _f32(deg) if deg is not None else None
    lib().oracle_spgemm_fwd16(_p(indptr), _p(indices), _p(values), _p(data), _p(sel16), ctypes.c_int64(n_rows),
                              ctypes.c_int(k), ctypes.c_int(dim), _p(deg), _p(out))
    return out


def sspmm_bwd16(indptr, indices, values, g, sel16, deg=None):
    indptr, indices, values, g, sel16 = _i32(indptr), _i32(indices), _f32(values), _f32(g), _u16(sel16)
    n_rows, dim = g.shape
    n_cols, k = sel16.shape
    gs = np.empty((n_cols, k), np.float32)
    deg = _f32(deg) if deg is not None else None
    lib().oracle_sspmm_bwd16(_p(indptr), _p(indices), _p(values), _p(g), _p(sel16), ctypes.c_int64(n_rows),
                             ctypes.c_int64(n_cols), ctypes.c_int(k), ctypes.c_int(dim), _p(deg), _p(gs))
    return gs


def spgemm_fwd_warp4(warp4_quads, indices, values, data, sel, n_rows, dim=256):
    warp4_quads, indices, values, data, sel = _i32(warp4_quads), _i32(indices), _f32(values), _f32(data), _u8(sel)
    k = data.shape[1]
    out = np.empty((n_rows, dim), np.float32)
    scratch = np.empty(n_rows * dim, np.float64)
    lib().oracle_spgemm_fwd_warp4(_p(warp4_quads), ctypes.c_int64(warp4_quads.size // 4), _p(indices), _p(values),
                                  _p(data), _p(sel), ctypes.c_int64(n_rows), ctypes.c_int(k), ctypes.c_int(dim),
                                  _p(scratch), _p(out))
    return out


def sspmm_bwd(indptr, indices, values, g, sel, deg=None):
    indptr, indices, values, g, sel = _i32(indptr), _i32(indices), _f32(values), _f32(g), _u8(sel)
    n_rows, dim = g.shape
    n_cols, k = sel.shape
    gs = np.empty((n_cols, k), np.float32)
    deg = _f32(deg) if deg is not None else None
    lib().oracle_sspmm_bwd(_p(indptr), _p(indices), _p(values), _p(g), _p(sel), ctypes.c_int64(n_rows),
                           ctypes.c_int64(n_cols), ctypes.c_int(k), ctypes.c_int(dim), _p(deg), _p(gs))
    return gs


def maxk_act_fwd(x, cols):
    x, cols = _f32(x), _i32(cols)
    n, d = x.shape
    out = np.empty_like(x)
    lib().oracle_maxk_act_fwd(_p(x), _p(cols), ctypes.c_int64(n), ctypes.c_int(d), ctypes.c_int(cols.shape[1]), _p(out))
    return out


def maxk_act_bwd(grad, cols):
    grad, cols = _f32(grad), _i32(cols)
    n, d = grad.shape
    out = np.empty_like(grad)
    lib().oracle_maxk_act_bwd(_p(grad), _p(cols), ctypes.c_int64(n), ctypes.c_int(d), ctypes.c_int(cols.shape[1]), _p(out))
    return out


def scatter_dense(vals, cols, dim=256):
    """dense[r, cols[r,l]] = vals[r,l] (zeros elsewhere) -- pure numpy."""
    vals = np.asarray(vals)
    out = np.zeros((vals.shape[0], dim), vals.dtype)
    np.put_along_axis(out, np.asarray(cols, dtype=np.int64), vals, axis=1)
    return out


# ---------------------------------------------------------------------------------------------
# Independent numpy/scipy restatements, used to cross-check the C oracle itself
# ---------------------------------------------------------------------------------------------
def spgemm_fwd_scipy(indptr, indices, values, data, sel, dim=256):
    import scipy.sparse as sp
    n_src = data.shape[0]
    a = sp.csr_matrix((np.asarray(values, np.float64), np.asarray(indices), np.asarray(indptr)),
                      shape=(len(indptr) - 1, n_src))
    return (a @ scatter_dense(np.asarray(data, np.float64), sel, dim)).astype(np.float32)


def sspmm_bwd_scipy(indptr, indices, values, g, sel):
    import scipy.sparse as sp
    n_dst = sel.shape[0]
    a = sp.csr_matrix((np.asarray(values, np.float64), np.asarray(indices), np.asarray(indptr)),
                      shape=(len(indptr) - 1, n_dst))
    full = a.T @ np.asarray(g, np.float64)
    return np.take_along_axis(full, np.asarray(sel, np.int64), axis=1).astype(np.float32)


# ---------------------------------------------------------------------------------------------
# The reference's own CUDA kernels (oracle/_ref), GPU only
# ---------------------------------------------------------------------------------------------
def ref_cuda_available():
    return os.path.exists(_REF_SO)


_R = None


def ref_lib():
    global _R
    if _R is None:
        _R = ctypes.CDLL(_REF_SO)
        _R.ref_spmm_maxk_forward.restype = ctypes.c_int
        _R.ref_spmm_maxk_backward.restype = ctypes.c_int
    return _R


def ref_cuda_forward(warp4_t, indices_t, values_t, data_t, sel_t, num_warps, dim=256):
    """Runs the reference's spmm_kernel_opt2_sparse_v3 (unmodified, sm_100a) on torch CUDA tensors."""
    import torch
    n, k = data_t.shape
    out = torch.empty((n, dim), dtype=torch.float32, device=data_t.device)
    st = ref_lib().ref_spmm_maxk_forward(
        ctypes.c_void_p(warp4_t.data_ptr()), ctypes.c_void_p(indices_t.data_ptr()), ctypes.c_void_p(values_t.data_ptr()),
        ctypes.c_void_p(data_t.data_ptr()), ctypes.c_void_p(sel_t.data_ptr()), ctypes.c_void_p(out.data_ptr()),
        ctypes.c_int(n), ctypes.c_int(indices_t.numel()), ctypes.c_int(dim), ctypes.c_int(k), ctypes.c_int(num_warps),
        ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    if st != 0:
        raise RuntimeError("reference forward kernel launch failed: cudaError %d" % st)
    return out


def ref_cuda_backward(warp4_t, indices_t, values_t, grad_t, sel_t, num_warps):
    import torch
    n, dim = grad_t.shape
    k = sel_t.shape[1]
    gs = torch.empty((n, k), dtype=torch.float32, device=grad_t.device)
    st = ref_lib().ref_spmm_maxk_backward(
        ctypes.c_void_p(warp4_t.data_ptr()), ctypes.c_void_p(indices_t.data_ptr()), ctypes.c_void_p(values_t.data_ptr()),
        ctypes.c_void_p(grad_t.data_ptr()), ctypes.c_void_p(sel_t.data_ptr()), ctypes.c_void_p(gs.data_ptr()),
        ctypes.c_int(n), ctypes.c_int(indices_t.numel()), ctypes.c_int(dim), ctypes.c_int(k), ctypes.c_int(num_warps),
        ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    if st != 0:
        raise RuntimeError("reference backward kernel launch failed: cudaError %d" % st)
    return gs
