"""maxk_cuda_kernels -- drop-in for the reference's pybind11 extension of the same name.

The reference builds `maxk_cuda_kernels` from cuda_kernel_bindings.cpp (exports at :429-490,
newer copy binding_v2.py:488-561).  This module provides the same Python-visible functions with
the same argument meaning, over the C ABI of include/maxk_b200.h (libmaxk_b200.so, hand-written
sm_100a kernels) through ctypes.  PyTorch is used only for device memory and the current stream.

Differences that are deliberate:
  * launches go to torch's CURRENT stream on the tensors' device, never synchronise, never print;
  * errors raise RuntimeError -- there is no cuSPARSE / CPU fallback anywhere;
  * cuda_topk_maxk_float is an exact fp32 top-k (the reference quantises to uint8);
  * `cusparse_spmm` keeps its name for the validators but is our own dense SpMM kernel.
"""
import ctypes
import os
import weakref

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.environ.get("MAXK_B200_LIB") or os.path.join(_HERE, "lib", "libmaxk_b200.so")   # env: developer A/B builds

FULL_DIM = 256          # cuda_kernel_bindings.cpp:70
WARPS_PER_BLOCK = 12    # kernels/generate_meta.py:8 (metadata contract only)
WARP_MAX_NZ = 64        # kernels/generate_meta.py:9

ORDER_VALUE_DESC = 0
ORDER_COLUMN_ASC = 1
ORDER_BANKED = 2          # residue classes mod 4 by size (include/maxk_b200.h): fewest bank conflicts in the forward

_c_i64 = ctypes.c_int64
_c_int = ctypes.c_int
_c_ptr = ctypes.c_void_p
_c_size = ctypes.c_size_t

# name -> (restype, argtypes); mirrors include/maxk_b200.h one to one (tests check this).
_SIGNATURES = {
    "maxk_abi_version": (_c_int, []),
    "maxk_status_string": (ctypes.c_char_p, [_c_int]),
    "maxk_banked_modulus": (_c_int, [_c_int]),
    "maxk_topk_cbsr": (_c_int, [_c_ptr, _c_i64, _c_int, _c_int, _c_int, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr]),
    "maxk_topk_cbsr_peers": (_c_int, [_c_ptr, _c_i64, _c_int, _c_int, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_i64, _c_ptr, _c_ptr]),
    "maxk_nvls_reduce": (_c_int, [_c_ptr, _c_ptr, _c_i64, _c_ptr]),
    "maxk_spgemm_forward": (_c_int, [_c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_i64, _c_i64, _c_int,
                                     _c_int, _c_ptr, _c_ptr, _c_size, _c_ptr]),
    "maxk_sspmm_backward": (_c_int, [_c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_i64, _c_i64, _c_i64,
                                     _c_int, _c_int, _c_ptr, _c_ptr, _c_size, _c_ptr]),
    "maxk_sspmm_backward_accumulate": (_c_int, [_c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_i64, _c_i64,
                                                _c_i64, _c_int, _c_int, _c_ptr, _c_ptr, _c_size, _c_ptr]),
    "maxk_topk_cbsr16": (_c_int, [_c_ptr, _c_i64, _c_int, _c_int, _c_ptr, _c_ptr, _c_ptr, _c_ptr]),
    "maxk_wide_workspace_bytes": (_c_size, [_c_i64, _c_i64, _c_int, _c_int]),
    "maxk_spgemm_forward16": (_c_int, [_c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_i64, _c_i64, _c_i64, _c_int, _c_int,
                                       _c_ptr, _c_ptr, _c_size, _c_ptr]),
    "maxk_sspmm_backward16": (_c_int, [_c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_i64, _c_i64, _c_i64, _c_int,
                                       _c_int, _c_ptr, _c_ptr, _c_size, _c_ptr]),
    "maxk_spgemm_workspace_bytes": (_c_size, [_c_i64]),
    "maxk_plan_bytes": (_c_size, [_c_i64]),
    "maxk_plan_workspace_bytes": (_c_size, [_c_i64]),
    "maxk_plan_build": (_c_int, [_c_ptr, _c_ptr, _c_i64, _c_ptr, _c_size, _c_ptr, _c_size, _c_ptr]),
    "maxk_spgemm_forward_planned": (_c_int, [_c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_i64, _c_i64, _c_int, _c_int,
                                             _c_ptr, _c_ptr]),
    "maxk_warp4_workspace_bytes": (_c_size, [_c_i64]),
    "maxk_warp4_scan": (_c_int, [_c_ptr, _c_i64, _c_int, _c_ptr, _c_ptr, _c_size, _c_ptr]),
    "maxk_warp4_fill": (_c_int, [_c_ptr, _c_ptr, _c_i64, _c_int, _c_ptr, _c_ptr]),
    "maxk_warp4_to_rows": (_c_int, [_c_ptr, _c_i64, _c_i64, _c_ptr, _c_ptr, _c_ptr]),
    "maxk_cbsr_scatter": (_c_int, [_c_ptr, _c_ptr, _c_i64, _c_int, _c_int, _c_ptr, _c_ptr]),
    "maxk_mask_apply": (_c_int, [_c_ptr, _c_ptr, _c_ptr, _c_i64, _c_int, _c_int, _c_ptr, _c_ptr]),
    "maxk_dense_spmm": (_c_int, [_c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_i64, _c_int, _c_ptr, _c_ptr]),
}


def _load_library():
    if not os.path.exists(_LIB_PATH):
        raise ImportError(
            "maxk_cuda_kernels: %s is missing. Build it with `python __graft_entry__.py` "
            "(nvcc, sm_100a). There is no CPU or cuSPARSE fallback." % _LIB_PATH)
    lib = ctypes.CDLL(_LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError here == the .so does not match the header
        fn.restype = res
        fn.argtypes = args
    if lib.maxk_abi_version() != 1:
        raise ImportError("maxk_cuda_kernels: ABI version mismatch in %s" % _LIB_PATH)
    return lib


_lib = _load_library()
LIBRARY_PATH = _LIB_PATH


# ----------------------------------------------------------------------------------------------
# plumbing
# ----------------------------------------------------------------------------------------------
def _check(cond, msg):
    if not cond:
        raise RuntimeError(msg)


def _status(code, what):
    if code != 0:
        raise RuntimeError("%s failed: %s (status %d)" % (what, _lib.maxk_status_string(code).decode(), code))


def _stream(t):
    return _c_ptr(torch.cuda.current_stream(t.device).cuda_stream)


def _ptr(t):
    return _c_ptr(t.data_ptr()) if t is not None else _c_ptr(0)


def _cuda(t, name, dtype=None):
    _check(isinstance(t, torch.Tensor), "%s must be a torch.Tensor" % name)
    _check(t.is_cuda, "%s must be CUDA tensor" % name)          # cuda_kernel_bindings.cpp:52-56
    if dtype is not None:
        _check(t.dtype == dtype, "%s must be %s" % (name, str(dtype).replace("torch.", "")))  # :58-62
    return t if t.is_contiguous() else t.contiguous()


# One scratch buffer per (device, stream): calls on a stream are ordered, so they can share it (the
# reference allocates nothing per call either; round 1 did a torch.empty per op).
_ws_cache = {}


def _workspace(n_rows, device):
    nbytes = _lib.maxk_spgemm_workspace_bytes(n_rows)
    key = (device.index if device.index is not None else torch.cuda.current_device(),
           torch.cuda.current_stream(device).cuda_stream)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty((nbytes + 255) // 256 * 256, dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf, nbytes


class RowPlan:
    """The forward kernel's row plan (csrc/plan.cu) of one CSR row structure: build once per graph, pass as
    `plan=` to spgemm_forward_csr.  Holds the device buffer and the row count it was built for."""

    def __init__(self, buf, n_rows, device):
        self.buf, self.n_rows, self.device = buf, n_rows, device

    def header(self):
        """(n_rows, shared long rows, separate groups, shared last-wave rows, trailing separate groups, items) --
        one device->host read, for reports and tests."""
        h = self.buf[:64].view(torch.int32).tolist()
        return {"n_rows": h[1], "long_rows": h[2], "groups": h[3], "last_wave_rows": h[4], "tail_groups": h[5], "items": h[8]}


def build_plan(row_begin, row_end):
    """GPU-built row plan for (row_begin, row_end) (for a plain CSR: indptr[:-1], indptr[1:])."""
    row_begin = _cuda(row_begin, "row_begin", torch.int32)
    row_end = _cuda(row_end, "row_end", torch.int32)
    n_rows = row_begin.numel()
    _check(row_end.numel() == n_rows, "row_begin / row_end length mismatch")
    dev = row_begin.device
    pb, wb = _lib.maxk_plan_bytes(n_rows), _lib.maxk_plan_workspace_bytes(n_rows)
    buf = torch.empty((pb + 255) // 256 * 256, dtype=torch.uint8, device=dev)
    ws = torch.empty((wb + 255) // 256 * 256, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _status(_lib.maxk_plan_build(_ptr(row_begin), _ptr(row_end), n_rows, _ptr(buf), pb, _ptr(ws), wb, _stream(row_begin)),
                "maxk_plan_build")
    ws.record_stream(torch.cuda.current_stream(dev))
    return RowPlan(buf, n_rows, dev)


# indptr tensor -> RowPlan, rebuilt only when the tensor object or its version changes (the operators receive
# graph_indptr on every call; a graph object keeps it alive, so the weak reference identifies it safely)
_indptr_plans = {}


def plan_for_indptr(indptr):
    """(row_begin, row_end, RowPlan) of a CSR indptr tensor, cached per tensor object."""
    if indptr.dtype != torch.int32:
        indptr = indptr.to(torch.int32)            # a temporary: planned for this call only
    key = id(indptr)
    hit = _indptr_plans.get(key)
    if hit is not None:
        ref, version, plan, ready = hit
        if ref() is indptr and version == indptr._version:
            torch.cuda.current_stream(indptr.device).wait_event(ready)
            return indptr[:-1], indptr[1:], plan
    _check(indptr.is_cuda and indptr.dim() == 1 and indptr.numel() >= 1, "graph_indptr must be a 1-D CUDA tensor")
    indptr = indptr if indptr.is_contiguous() else indptr.contiguous()
    plan = build_plan(indptr[:-1], indptr[1:])
    if len(_indptr_plans) > 64:
        for k in [k for k, v in _indptr_plans.items() if v[0]() is None]:
            del _indptr_plans[k]
    try:
        _indptr_plans[key] = (weakref.ref(indptr), indptr._version, plan, _ready_event(indptr.device))
    except TypeError:
        pass
    return indptr[:-1], indptr[1:], plan


def rows_and_plan(warp4_metadata, num_warps, graph_indptr, n_rows):
    """Row edge ranges + row plan from the CSR indptr when given, else from the warp4 quads."""
    if graph_indptr is not None:
        return plan_for_indptr(graph_indptr)
    if warp4_metadata is None:
        raise RuntimeError("maxk_spgemm needs warp4_metadata or graph_indptr (there is no fallback path)")
    rows = _rows_from_warp4(warp4_metadata, int(num_warps), n_rows)
    torch.cuda.current_stream(warp4_metadata.device).wait_event(rows[3])
    return rows[0], rows[1], rows[2]


# warp4 tensor -> (row_begin, row_end), rebuilt only when the tensor object or its version changes
_rows_cache = {}


def _rows_from_warp4(warp4, num_warps, n_rows):
    key = id(warp4)
    hit = _rows_cache.get(key)
    if hit is not None:
        ref, version, nw, nr, rows = hit
        if ref() is warp4 and version == warp4._version and nw == num_warps and nr == n_rows:
            return rows
    _check(warp4.numel() >= 4 * num_warps, "warp4_metadata holds fewer than num_warps quads")
    rows = torch.empty((2, max(n_rows, 1)), dtype=torch.int32, device=warp4.device)
    with torch.cuda.device(warp4.device):
        _status(_lib.maxk_warp4_to_rows(_ptr(warp4), num_warps, n_rows, _ptr(rows[0]), _ptr(rows[1]),
                                        _stream(warp4)), "maxk_warp4_to_rows")
    rows = (rows[0], rows[1], build_plan(rows[0], rows[1]) if n_rows > 0 else None, _ready_event(warp4.device))
    if len(_rows_cache) > 64:
        for k in [k for k, v in _rows_cache.items() if v[0]() is None]:
            del _rows_cache[k]
    try:
        _rows_cache[key] = (weakref.ref(warp4), warp4._version, num_warps, n_rows, rows)
    except TypeError:
        pass
    return rows


# ----------------------------------------------------------------------------------------------
# CSR-native entry points (additive API; the warp4 entry points below reduce to these)
# ----------------------------------------------------------------------------------------------
def _ready_event(device):
    """Event recorded on the current stream: cached per-graph data built here is waited for by other streams."""
    ev = torch.cuda.Event()
    ev.record(torch.cuda.current_stream(device))
    return ev


def _check_graph(indices, n_src, row_end, where):
    """Opt-in (MAXK_VALIDATE=1) bounds check of the graph arrays: malformed input otherwise becomes out-of-bounds
    device gathers / reductions instead of an exception."""
    if os.environ.get("MAXK_VALIDATE", "0") != "1" or indices.numel() == 0:
        return
    _check(int(indices.min()) >= 0 and int(indices.max()) < n_src, "%s: a column index is outside [0, %d)" % (where, n_src))
    _check(int(row_end.max()) <= indices.numel(), "%s: a row range ends past the edge arrays" % where)


def spgemm_forward_csr(row_begin, row_end, indices, values, cbsr_val, cbsr_sel, out_dim=FULL_DIM,
                       row_div=None, out=None, plan=None):
    """out[n_rows, out_dim] = A_csr x scatter(CBSR), optionally / row_div (fused).
    plan: the RowPlan of (row_begin, row_end) (build_plan); without it the plan is rebuilt inside the call."""
    row_begin = _cuda(row_begin, "row_begin", torch.int32)
    row_end = _cuda(row_end, "row_end", torch.int32)
    indices = _cuda(indices, "indices", torch.int32)
    values = _cuda(values, "values", torch.float32)
    cbsr_val = _cuda(cbsr_val, "input_data", torch.float32)
    cbsr_sel = _cuda(cbsr_sel, "sparse_selector", torch.uint8)
    _check(cbsr_val.dim() == 2 and cbsr_sel.shape == cbsr_val.shape, "input_data / sparse_selector must both be [N, k]")
    _check(indices.numel() == values.numel(), "indices and values must have the same length")
    n_rows, k = row_begin.numel(), cbsr_val.size(1)
    _check(row_end.numel() == n_rows, "row_begin / row_end length mismatch")
    if row_div is not None:
        row_div = _cuda(row_div, "row_div", torch.float32)
        _check(row_div.numel() == n_rows, "row_div must have one entry per row")
    if out is None:
        out = torch.empty((n_rows, out_dim), dtype=torch.float32, device=cbsr_val.device)
    else:
        _check(out.is_cuda and out.is_contiguous() and out.dtype == torch.float32 and
               tuple(out.shape) == (n_rows, out_dim), "out must be a contiguous fp32 CUDA [n_rows, out_dim] tensor")
    _check_graph(indices, cbsr_val.size(0), row_end, "spgemm_forward")
    with torch.cuda.device(cbsr_val.device):
        if plan is not None:
            _check(isinstance(plan, RowPlan) and plan.n_rows == n_rows and plan.device == cbsr_val.device,
                   "plan was built for another row structure / device")
            _status(_lib.maxk_spgemm_forward_planned(_ptr(plan.buf), _ptr(indices), _ptr(values), _ptr(cbsr_val),
                                                     _ptr(cbsr_sel), _ptr(out), n_rows, indices.numel(), out_dim, k,
                                                     _ptr(row_div), _stream(cbsr_val)), "maxk_spgemm_forward_planned")
        else:
            ws, ws_bytes = _workspace(n_rows, cbsr_val.device)
            _status(_lib.maxk_spgemm_forward(_ptr(row_begin), _ptr(row_end), _ptr(indices), _ptr(values), _ptr(cbsr_val),
                                             _ptr(cbsr_sel), _ptr(out), n_rows, indices.numel(), out_dim, k,
                                             _ptr(row_div), _ptr(ws), ws_bytes, _stream(cbsr_val)),
                    "maxk_spgemm_forward")
    return out


def sspmm_backward_csr(row_begin, row_end, indices, values, grad_output, cbsr_sel, row_div=None, out=None,
                       accumulate=False):
    """gs[n_dst, k] = sample_sel(A_csr^T (grad_output / row_div)); accumulate=True adds into `out` instead."""
    row_begin = _cuda(row_begin, "row_begin", torch.int32)
    row_end = _cuda(row_end, "row_end", torch.int32)
    indices = _cuda(indices, "indices", torch.int32)
    values = _cuda(values, "values", torch.float32)
    grad_output = _cuda(grad_output, "grad_output", torch.float32)
    cbsr_sel = _cuda(cbsr_sel, "sparse_selector", torch.uint8)
    _check(grad_output.dim() == 2 and cbsr_sel.dim() == 2, "grad_output must be [N, D], sparse_selector [N, k]")
    n_rows, dim = row_begin.numel(), grad_output.size(1)
    n_dst, k = cbsr_sel.shape
    _check(grad_output.size(0) == n_rows, "grad_output must have one row per CSR row")
    if row_div is not None:
        row_div = _cuda(row_div, "row_div", torch.float32)
        _check(row_div.numel() == n_rows, "row_div must have one entry per row")
    _check(not (accumulate and out is None), "accumulate=True needs an `out` tensor to add into")
    if out is None:
        out = torch.empty((n_dst, k), dtype=torch.float32, device=grad_output.device)
    else:
        _check(out.is_cuda and out.is_contiguous() and out.dtype == torch.float32 and tuple(out.shape) == (n_dst, k),
               "out must be a contiguous fp32 CUDA [n_dst, k] tensor")
    fn = _lib.maxk_sspmm_backward_accumulate if accumulate else _lib.maxk_sspmm_backward
    _check_graph(indices, n_dst, row_end, "sspmm_backward")
    with torch.cuda.device(grad_output.device):
        ws, ws_bytes = _workspace(n_rows, grad_output.device)
        _status(fn(_ptr(row_begin), _ptr(row_end), _ptr(indices), _ptr(values),
                                         _ptr(grad_output), _ptr(cbsr_sel), _ptr(out), n_rows, n_dst,
                                         indices.numel(), dim, k, _ptr(row_div), _ptr(ws), ws_bytes,
                                         _stream(grad_output)), "maxk_sspmm_backward")
    return out


def topk_cbsr(x, k, order=ORDER_BANKED, want_sel=True, want_i32=False, want_i64=False, want_masked=False,
              out_values=None, out_sel=None):
    """Exact row-wise top-k of x[N, D<=256] -> dict(values, sel, i32, i64, masked).
    out_values / out_sel: optional preallocated contiguous [N, k] destinations (e.g. row slabs of a larger CBSR)."""
    x = _cuda(x, "input", torch.float32)
    _check(x.dim() == 2, "Input must be 2D tensor")
    n, d = x.shape
    _check(0 < k <= d, "Invalid k value")
    _check(d <= FULL_DIM, "feature dim must be <= 256 (uint8 column selectors)")
    dev = x.device
    vals = out_values if out_values is not None else torch.empty((n, k), dtype=torch.float32, device=dev)
    sel = out_sel if out_sel is not None else (torch.empty((n, k), dtype=torch.uint8, device=dev) if want_sel else None)
    for t, dt, nm in ((vals, torch.float32, "out_values"), (sel, torch.uint8, "out_sel")):
        _check(t is None or (t.is_cuda and t.is_contiguous() and t.dtype == dt and tuple(t.shape) == (n, k)),
               nm + " must be a contiguous CUDA [N, k] tensor of the right dtype")
    i32 = torch.empty((n, k), dtype=torch.int32, device=dev) if want_i32 else None
    i64 = torch.empty((n, k), dtype=torch.int64, device=dev) if want_i64 else None
    masked = torch.empty((n, d), dtype=torch.float32, device=dev) if want_masked else None
    with torch.cuda.device(dev):
        _status(_lib.maxk_topk_cbsr(_ptr(x), n, d, k, order, _ptr(vals), _ptr(sel), _ptr(i32), _ptr(i64),
                                    _ptr(masked), _stream(x)), "maxk_topk_cbsr")
    return {"values": vals, "sel": sel, "i32": i32, "i64": i64, "masked": masked}


# ----------------------------------------------------------------------------------------------
# Feature widths above 256: uint16 selectors (csrc/wide.cu).  torch has no uint16 arithmetic, so the selector
# tensors are int16 views of the same bits (use .view(torch.uint16) / .to(torch.int32) & 0xffff to read them).
# ----------------------------------------------------------------------------------------------
WIDE_MAX_DIM = 1024


def topk_cbsr16(x, k, want_masked=False):
    """Exact row-wise top-k of x[N, D <= 1024] -> dict(values fp32 [N, k], sel16 int16-typed uint16 bits [N, k] in
    column-ascending order, masked [N, D] or None)."""
    x = _cuda(x, "input", torch.float32)
    _check(x.dim() == 2, "Input must be 2D tensor")
    n, d = x.shape
    _check(0 < k <= min(d, FULL_DIM), "Invalid k value (1 <= k <= min(D, 256))")
    _check(d <= WIDE_MAX_DIM, "feature dim must be <= 1024")
    vals = torch.empty((n, k), dtype=torch.float32, device=x.device)
    sel = torch.empty((n, k), dtype=torch.int16, device=x.device)
    masked = torch.empty_like(x) if want_masked else None
    with torch.cuda.device(x.device):
        _status(_lib.maxk_topk_cbsr16(_ptr(x), n, d, int(k), _ptr(vals), _ptr(sel), _ptr(masked), _stream(x)), "maxk_topk_cbsr16")
    return {"values": vals, "sel16": sel, "masked": masked}


def _wide_workspace(n_rows, n_src, dim, k, device):
    nbytes = _lib.maxk_wide_workspace_bytes(n_rows, n_src, dim, k)
    return torch.empty(nbytes + 256, dtype=torch.uint8, device=device), nbytes


def spgemm_forward16_csr(row_begin, row_end, indices, values, cbsr_val, cbsr_sel16, dim, row_div=None, plan=None):
    """out[n_rows, dim] = A_csr x scatter(CBSR with uint16 selectors), optionally / row_div."""
    row_begin = _cuda(row_begin, "row_begin", torch.int32)
    row_end = _cuda(row_end, "row_end", torch.int32)
    indices = _cuda(indices, "indices", torch.int32)
    values = _cuda(values, "values", torch.float32)
    cbsr_val = _cuda(cbsr_val, "input_data", torch.float32)
    cbsr_sel16 = _cuda(cbsr_sel16, "sparse_selector", torch.int16)
    _check(cbsr_val.dim() == 2 and cbsr_sel16.shape == cbsr_val.shape, "input_data / sparse_selector must both be [N, k]")
    n_rows, (n_src, k) = row_begin.numel(), cbsr_val.shape
    _check(0 < dim <= WIDE_MAX_DIM, "dim must be in [1, 1024]")
    if row_div is not None:
        row_div = _cuda(row_div, "row_div", torch.float32)
    if plan is None:
        plan = build_plan(row_begin, row_end)
    _check(plan.n_rows == n_rows, "plan was built for another row structure")
    _check_graph(indices, n_src, row_end, "spgemm_forward16")
    out = torch.empty((n_rows, dim), dtype=torch.float32, device=cbsr_val.device)
    with torch.cuda.device(cbsr_val.device):
        ws, ws_bytes = _wide_workspace(n_rows, n_src, dim, k, cbsr_val.device)
        _status(_lib.maxk_spgemm_forward16(_ptr(plan.buf), _ptr(indices), _ptr(values), _ptr(cbsr_val), _ptr(cbsr_sel16), _ptr(out),
                                           n_rows, n_src, indices.numel(), int(dim), k, _ptr(row_div), _ptr(ws), ws_bytes,
                                           _stream(cbsr_val)), "maxk_spgemm_forward16")
    return out


def sspmm_backward16_csr(row_begin, row_end, indices, values, grad_output, cbsr_sel16, row_div=None):
    """gs[n_dst, k] = sample_sel(A_csr^T (grad_output / row_div)) for uint16 selectors, grad_output [n_rows, D <= 1024]."""
    row_begin = _cuda(row_begin, "row_begin", torch.int32)
    row_end = _cuda(row_end, "row_end", torch.int32)
    indices = _cuda(indices, "indices", torch.int32)
    values = _cuda(values, "values", torch.float32)
    grad_output = _cuda(grad_output, "grad_output", torch.float32)
    cbsr_sel16 = _cuda(cbsr_sel16, "sparse_selector", torch.int16)
    n_rows, dim = row_begin.numel(), grad_output.size(1)
    n_dst, k = cbsr_sel16.shape
    _check(grad_output.size(0) == n_rows and dim <= WIDE_MAX_DIM, "grad_output must be [n_rows, D <= 1024]")
    if row_div is not None:
        row_div = _cuda(row_div, "row_div", torch.float32)
    _check_graph(indices, n_dst, row_end, "sspmm_backward16")
    gs = torch.empty((n_dst, k), dtype=torch.float32, device=grad_output.device)
    with torch.cuda.device(grad_output.device):
        ws, ws_bytes = _wide_workspace(n_rows, n_dst, dim, k, grad_output.device)
        _status(_lib.maxk_sspmm_backward16(_ptr(row_begin), _ptr(row_end), _ptr(indices), _ptr(values), _ptr(grad_output),
                                           _ptr(cbsr_sel16), _ptr(gs), n_rows, n_dst, indices.numel(), dim, k, _ptr(row_div),
                                           _ptr(ws), ws_bytes, _stream(grad_output)), "maxk_sspmm_backward16")
    return gs


class WideMaxKSpGEMMFunction(torch.autograd.Function):
    """x[N, D in (256, 1024]] -> A @ maxk(x) / in_degrees with gradients to x: the generation-1 operator
    (maxk_spgemm_function.py) for feature widths its uint8 selectors cannot address."""

    @staticmethod
    def forward(ctx, indptr, indices, values, x, k, in_degrees=None, out_degrees=None):
        row_begin, row_end, plan = plan_for_indptr(indptr)
        r = topk_cbsr16(x, int(k))
        ctx.save_for_backward(row_begin, row_end, indices, values, r["sel16"],
                              out_degrees if out_degrees is not None else torch.empty(0, device=x.device))
        ctx.has_div, ctx.dim = out_degrees is not None, x.size(1)
        return spgemm_forward16_csr(row_begin, row_end, indices, values, r["values"], r["sel16"], x.size(1),
                                    row_div=in_degrees, plan=plan)

    @staticmethod
    def backward(ctx, grad_output):
        row_begin, row_end, indices, values, sel16, out_degrees = ctx.saved_tensors
        gs = sspmm_backward16_csr(row_begin, row_end, indices, values, grad_output.contiguous(), sel16,
                                  row_div=out_degrees if ctx.has_div else None)
        grad_x = torch.zeros(gs.size(0), ctx.dim, dtype=gs.dtype, device=gs.device)
        grad_x.scatter_(1, sel16.to(torch.int64) & 0xffff, gs)
        return None, None, None, grad_x, None, None, None


MAX_PEERS = 8


def topk_cbsr_to_peers(x, k, peer_val_ptrs, peer_sel_ptrs, row_offset, want_masked=False, mc_val_ptr=0, mc_sel_ptr=0):
    """Row-sharded top-k: row r of x -> row (row_offset + r) of every peer's gathered CBSR buffers (raw device
    pointers of peer-mapped [P*m, k] fp32 / uint8 buffers, own rank included; mc_*_ptr: their NVLS multicast
    mappings, 0 = none).  Returns the masked rows or None."""
    x = _cuda(x, "input", torch.float32)
    _check(x.dim() == 2 and x.size(1) == FULL_DIM, "the peer-writing top-k needs [rows, 256] features")
    _check(len(peer_val_ptrs) == len(peer_sel_ptrs) and 1 <= len(peer_val_ptrs) <= MAX_PEERS, "1..8 peers")
    n = x.size(0)
    pv = (ctypes.c_void_p * len(peer_val_ptrs))(*[int(p) for p in peer_val_ptrs])
    ps = (ctypes.c_void_p * len(peer_sel_ptrs))(*[int(p) for p in peer_sel_ptrs])
    masked = torch.empty_like(x) if want_masked else None
    with torch.cuda.device(x.device):
        _status(_lib.maxk_topk_cbsr_peers(_ptr(x), n, int(k), len(peer_val_ptrs), pv, ps, _c_ptr(int(mc_val_ptr)),
                                          _c_ptr(int(mc_sel_ptr)), int(row_offset), _ptr(masked), _stream(x)),
                "maxk_topk_cbsr_peers")
    return masked


def nvls_reduce(mc_src_ptr, dst):
    """dst[...] = sum over the ranks of the multicast group of their buffers at mc_src_ptr (a raw multicast address),
    reduced inside the NVSwitch."""
    _check(dst.is_cuda and dst.is_contiguous() and dst.dtype == torch.float32 and dst.numel() % 4 == 0,
           "dst must be a contiguous fp32 CUDA tensor with a multiple of 4 elements")
    with torch.cuda.device(dst.device):
        _status(_lib.maxk_nvls_reduce(_c_ptr(int(mc_src_ptr)), _ptr(dst), dst.numel(), _stream(dst)), "maxk_nvls_reduce")
    return dst


def cbsr_scatter(vals, sel, dim=FULL_DIM):
    """dense[N, dim] with dense[r, sel[r,l]] = vals[r,l] and zeros elsewhere."""
    vals = _cuda(vals, "vals", torch.float32)
    sel = _cuda(sel, "sel", torch.uint8)
    _check(vals.shape == sel.shape and vals.dim() == 2, "vals / sel must both be [N, k]")
    n, k = vals.shape
    out = torch.empty((n, dim), dtype=torch.float32, device=vals.device)
    with torch.cuda.device(vals.device):
        _status(_lib.maxk_cbsr_scatter(_ptr(vals), _ptr(sel), n, dim, k, _ptr(out), _stream(vals)),
                "maxk_cbsr_scatter")
    return out


def mask_apply(dense, sel, add_vals=None):
    """out = dense * mask(sel) (+ add_vals scattered at the selected positions)."""
    dense = _cuda(dense, "dense", torch.float32)
    sel = _cuda(sel, "sel", torch.uint8)
    n, d = dense.shape
    k = sel.size(1)
    if add_vals is not None:
        add_vals = _cuda(add_vals, "add_vals", torch.float32)
        _check(add_vals.shape == sel.shape, "add_vals must be [N, k]")
    out = torch.empty_like(dense)
    with torch.cuda.device(dense.device):
        _status(_lib.maxk_mask_apply(_ptr(dense), _ptr(sel), _ptr(add_vals), n, d, k, _ptr(out), _stream(dense)),
                "maxk_mask_apply")
    return out


def maxk_layer_forward(indptr, indices, values, x, k, row_div=None, want_masked=False, plan=None):
    """The layer's forward in one call (additive): top-k -> CBSR -> SpGEMM with the fused divisor.
    Returns (out [N, 256], cbsr_val [N, k], cbsr_sel [N, k] uint8, masked [N, D] or None); keep cbsr_sel for
    maxk_layer_backward.  What MaxK.forward + MaxKSpGEMMFunction.forward do in the reference
    (maxk_models_integrated.py:28-37 + maxk_spgemm_function.py:51-86: two torch.topk, a scatter, the kernel,
    a division) as two kernel launches."""
    indptr = _cuda(indptr, "indptr", torch.int32)
    r = topk_cbsr(x, k, order=ORDER_BANKED, want_masked=want_masked)
    if plan is None:
        plan = plan_for_indptr(indptr)[2]
    out = spgemm_forward_csr(indptr[:-1], indptr[1:], indices, values, r["values"], r["sel"], row_div=row_div, plan=plan)
    return out, r["values"], r["sel"], r["masked"]


def maxk_layer_backward(indptr, indices, values, grad_output, cbsr_sel, row_div=None, dense_dim=None):
    """The layer's backward in one call: gs [N, k] = sample_sel(A^T (grad_output / row_div)); with dense_dim the
    gradient is also scattered to [N, dense_dim] (what maxk_spgemm_function.py:152-175 returns for
    input_features).  Returns gs, or (gs, dense)."""
    indptr = _cuda(indptr, "indptr", torch.int32)
    gs = sspmm_backward_csr(indptr[:-1], indptr[1:], indices, values, grad_output, cbsr_sel, row_div=row_div)
    if dense_dim is None:
        return gs
    return gs, cbsr_scatter(gs, cbsr_sel, dim=int(dense_dim))


def build_warp4(indptr, warp_max_nz=WARP_MAX_NZ):
    """GPU replacement of kernels/generate_meta.py: returns (warp4 int32[4W], W)."""
    indptr = _cuda(indptr, "indptr", torch.int32)
    n_rows = indptr.numel() - 1
    _check(n_rows >= 0, "indptr must have at least one entry")
    dev = indptr.device
    seg = torch.empty(n_rows + 1, dtype=torch.int32, device=dev)
    nbytes = _lib.maxk_warp4_workspace_bytes(n_rows)
    ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _status(_lib.maxk_warp4_scan(_ptr(indptr), n_rows, warp_max_nz, _ptr(seg), _ptr(ws), nbytes,
                                     _stream(indptr)), "maxk_warp4_scan")
        num_warps = int(seg[n_rows].item())   # the one host read-back: W sizes the allocation
        warp4 = torch.empty(4 * num_warps, dtype=torch.int32, device=dev)
        if num_warps:
            _status(_lib.maxk_warp4_fill(_ptr(indptr), _ptr(seg), n_rows, warp_max_nz, _ptr(warp4),
                                         _stream(indptr)), "maxk_warp4_fill")
    return warp4, num_warps


# ----------------------------------------------------------------------------------------------
# The reference's exports (cuda_kernel_bindings.cpp:429-490), same names and argument order
# ----------------------------------------------------------------------------------------------
def spmm_maxk_forward(warp4_metadata, indices, values, input_data, sparse_selector, num_warps, dim_sparse):
    """cuda_kernel_bindings.cpp:42-104.  Returns fp32 [N, 256] (FULL_DIM hard-wired, :70)."""
    warp4_metadata = _cuda(warp4_metadata, "warp4_metadata", torch.int32)
    input_data = _cuda(input_data, "input_data", torch.float32)
    _check(input_data.dim() == 2, "input_data must be [N, k]")
    _check(int(dim_sparse) == input_data.size(1), "dim_sparse must equal input_data.size(1)")
    rows = _rows_from_warp4(warp4_metadata, int(num_warps), input_data.size(0))
    torch.cuda.current_stream(input_data.device).wait_event(rows[3])     # the cache may have been filled on another stream
    return spgemm_forward_csr(rows[0], rows[1], indices, values, input_data, sparse_selector, FULL_DIM, plan=rows[2])


def spmm_maxk_backward(warp4_metadata, indices, values, grad_output, sparse_selector, num_warps, dim_sparse):
    """cuda_kernel_bindings.cpp:106-161.  Returns fp32 [N, dim_sparse]."""
    warp4_metadata = _cuda(warp4_metadata, "warp4_metadata", torch.int32)
    grad_output = _cuda(grad_output, "grad_output", torch.float32)
    sparse_selector = _cuda(sparse_selector, "sparse_selector", torch.uint8)
    _check(grad_output.dim() == 2, "grad_output must be [N, D]")
    _check(int(dim_sparse) == sparse_selector.size(1), "dim_sparse must equal sparse_selector.size(1)")
    rows = _rows_from_warp4(warp4_metadata, int(num_warps), grad_output.size(0))
    torch.cuda.current_stream(grad_output.device).wait_event(rows[3])
    return sspmm_backward_csr(rows[0], rows[1], indices, values, grad_output, sparse_selector)


def cuda_topk_maxk(input, k):
    """cuda_kernel_bindings.cpp:164-201: uint8 [N, D] -> (uint8 values, uint8 indices), exact."""
    _check(isinstance(input, torch.Tensor) and input.is_cuda, "Input must be on CUDA")
    _check(input.dim() == 2, "Input must be 2D tensor")
    _check(input.dtype == torch.uint8, "Input must be uint8 tensor")
    _check(0 < k <= input.size(1), "Invalid k value")
    r = topk_cbsr(input.to(torch.float32), k, order=ORDER_VALUE_DESC)
    return r["values"].to(torch.uint8), r["sel"]


def cuda_topk_maxk_float(input, k):
    """cuda_kernel_bindings.cpp:203-238: (values fp32 [N,k], indices int32 [N,k]).

    Exact fp32 selection in torch.topk order (value desc, lowest column first on ties); the
    reference quantises to round(x*255) uint8 first (:214-215) and is only valid for k = 32."""
    _check(isinstance(input, torch.Tensor) and input.is_cuda, "Input must be on CUDA")
    _check(input.dim() == 2, "Input must be 2D tensor")
    _check(0 < k <= input.size(1), "Invalid k value")
    _check(input.dtype in (torch.float32, torch.uint8), "Input must be float32 or uint8")
    if input.dtype == torch.uint8:
        r = topk_cbsr(input.to(torch.float32), k, order=ORDER_VALUE_DESC, want_sel=False, want_i32=True)
        return r["values"].to(torch.uint8), r["i32"]
    r = topk_cbsr(input, k, order=ORDER_VALUE_DESC, want_sel=False, want_i32=True)
    return r["values"], r["i32"]


def prepare_cbsr_format_maxk(features, maxk):
    """cuda_kernel_bindings.cpp:240-251."""
    _check(isinstance(features, torch.Tensor) and features.is_cuda, "Features must be on CUDA")
    _check(features.dim() == 2, "Features must be 2D tensor")
    _check(0 < maxk <= features.size(1), "Invalid maxk value")
    return cuda_topk_maxk_float(features, maxk)


def _load_quads(path):
    if not os.path.exists(path):
        raise RuntimeError("Cannot open warp4 file: " + path)     # cuda_kernel_bindings.cpp:296-298
    return torch.from_numpy(np.fromfile(path, dtype=np.int32).copy()).cuda()


def load_warp4_metadata(graph_name, num_warps=WARPS_PER_BLOCK, warp_max_nz=WARP_MAX_NZ):
    """cuda_kernel_bindings.cpp:287-317: reads kernels/w{12}_nz{64}_warp_4/<graph>.warp4."""
    return _load_quads("kernels/w%d_nz%d_warp_4/%s.warp4" % (num_warps, warp_max_nz, graph_name))


def load_warp4_metadata_csc(graph_name, num_warps=WARPS_PER_BLOCK, warp_max_nz=WARP_MAX_NZ):
    """binding_v2.py:320-351: reads kernels/w{12}_nz{64}_warp_4_csc/<graph>.warp4_csc."""
    return _load_quads("kernels/w%d_nz%d_warp_4_csc/%s.warp4_csc" % (num_warps, warp_max_nz, graph_name))


def cusparse_spmm(indptr, indices, values, input_features, timing=False):
    """cuda_kernel_bindings.cpp:253-284 by name; a hand-written CSR x dense kernel, no cuSPARSE."""
    for t, n in ((indptr, "indptr"), (indices, "indices"), (values, "values"), (input_features, "input_features")):
        _check(isinstance(t, torch.Tensor) and t.is_cuda and t.is_contiguous(), n + " must be CUDA and contiguous")
    _check(indptr.dtype == torch.int32 and indices.dtype == torch.int32, "indptr / indices must be int32")
    _check(values.dtype == torch.float32 and input_features.dtype == torch.float32, "values / features must be float32")
    n_rows, dim = indptr.numel() - 1, input_features.size(1)
    out = torch.empty((n_rows, dim), dtype=torch.float32, device=input_features.device)
    with torch.cuda.device(input_features.device):
        _status(_lib.maxk_dense_spmm(_ptr(indptr), _ptr(indices), _ptr(values), _ptr(input_features), n_rows, dim,
                                     _ptr(out), _stream(input_features)), "maxk_dense_spmm")
    return out


def generate_sparse_selector(num_v, dim_origin, dim_sparse):
    """cuda_kernel_bindings.cpp:320-340: k distinct random columns per row, seed 123."""
    _check(0 < dim_sparse <= dim_origin <= FULL_DIM, "need 0 < dim_sparse <= dim_origin <= 256")
    gen = torch.Generator(device="cuda")
    gen.manual_seed(123)
    keys = torch.rand((num_v, dim_origin), device="cuda", generator=gen)
    return keys.argsort(dim=1)[:, :dim_sparse].to(torch.uint8).contiguous()


class CudaTimer:
    """cuda_kernel_bindings.cpp:343-369 (SimpleCudaTimer), on the current stream."""

    def __init__(self):
        self._start = torch.cuda.Event(enable_timing=True)
        self._stop = torch.cuda.Event(enable_timing=True)

    def start(self):
        self._start.record()

    def stop(self):
        self._stop.record()
        self._stop.synchronize()
        return self._start.elapsed_time(self._stop)


def benchmark_spmm_maxk(warp4_metadata, indices, values, input_data, sparse_selector, num_warps, dim_sparse,
                        num_runs=4):
    """cuda_kernel_bindings.cpp:372-402: num_runs warm-up + num_runs timed launches, ms each."""
    for _ in range(num_runs):
        spmm_maxk_forward(warp4_metadata, indices, values, input_data, sparse_selector, num_warps, dim_sparse)
    torch.cuda.synchronize()
    timer, times = CudaTimer(), []
    for _ in range(num_runs):
        timer.start()
        spmm_maxk_forward(warp4_metadata, indices, values, input_data, sparse_selector, num_warps, dim_sparse)
        times.append(timer.stop())
    return times


def validate_spmm_maxk(warp4_metadata, indices, values, input_data, sparse_selector, reference_output, num_warps,
                       dim_sparse, tolerance=0.001):
    """cuda_kernel_bindings.cpp:405-427: mean |out - reference| < tolerance."""
    out = spmm_maxk_forward(warp4_metadata, indices, values, input_data, sparse_selector, num_warps, dim_sparse)
    return bool((out - reference_output).abs().mean().item() < tolerance)


def validate_spmm_maxk_backward(warp4_metadata, indices, values, grad_output, sparse_selector, reference_output,
                                num_warps, dim_sparse, tolerance=0.001):
    """binding_v2.py:464-486."""
    out = spmm_maxk_backward(warp4_metadata, indices, values, grad_output, sparse_selector, num_warps, dim_sparse)
    return bool((out - reference_output).abs().mean().item() < tolerance)
