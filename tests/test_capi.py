"""CPU: the C-ABI library loads and exports every symbol include/maxk_b200.h declares (no compute calls)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "maxk_b200.h")
LIB = os.path.join(ROOT, "spgemm-prunning_b200", "lib", "libmaxk_b200.so")


def declared_functions():
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    return sorted(set(re.findall(r"\b(maxk_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_three_hot_path_operations():
    names = declared_functions()
    for required in ("maxk_topk_cbsr", "maxk_spgemm_forward", "maxk_sspmm_backward", "maxk_warp4_scan",
                     "maxk_warp4_fill", "maxk_warp4_to_rows"):
        assert required in names


def test_library_exports_every_declared_symbol():
    assert os.path.exists(LIB), "build first: python __graft_entry__.py"
    lib = ctypes.CDLL(LIB)
    for name in declared_functions():
        assert hasattr(lib, name), "libmaxk_b200.so does not export " + name


def test_python_binding_binds_exactly_the_header():
    import maxk_cuda_kernels as k
    assert sorted(k._SIGNATURES) == declared_functions()
    assert k._lib.maxk_abi_version() == 1
    assert k._lib.maxk_status_string(-2).decode().startswith("dim must")
    assert [k._lib.maxk_banked_modulus(x) for x in (8, 16, 32, 64, 96, 128, 19)] == [4, 4, 4, 4, 4, 4, 1]
    assert k._lib.maxk_spgemm_workspace_bytes(1000) >= 4000


def test_reference_export_names_are_present():
    """Python-visible surface of the reference extension (cuda_kernel_bindings.cpp:429-490, binding_v2.py:488-561)."""
    import maxk_cuda_kernels as k
    for name in ("spmm_maxk_forward", "spmm_maxk_backward", "cuda_topk_maxk", "cuda_topk_maxk_float",
                 "prepare_cbsr_format_maxk", "load_warp4_metadata", "load_warp4_metadata_csc", "cusparse_spmm",
                 "generate_sparse_selector", "benchmark_spmm_maxk", "validate_spmm_maxk",
                 "validate_spmm_maxk_backward", "CudaTimer"):
        assert hasattr(k, name), name


def test_product_code_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under spgemm-prunning_b200/ may import or load it."""
    pkg = os.path.join(ROOT, "spgemm-prunning_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "libmaxk_oracle" not in text and "libmaxk_ref" not in text, f


def test_header_is_plain_c_and_a_c_host_links_against_the_library(tmp_path):
    """The boundary is a C ABI, not a C++ one: the header compiles as C99 on its own, and the plain-C host in
    examples/ (the binding a cgo / JNI / Rust caller would write) compiles and links against libmaxk_b200.so
    with nothing but the CUDA runtime.  (It is not run here: no GPU.)"""
    import shutil
    import subprocess
    gcc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else shutil.which("gcc")
    assert gcc, "no C compiler"
    probe = tmp_path / "probe.c"
    probe.write_text('#include "maxk_b200.h"\nint main(void) { int (*f)(void) = maxk_abi_version; return f == 0; }\n')
    subprocess.run([gcc, "-std=c99", "-pedantic", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"),
                    str(probe)], check=True)
    cuda = "/usr/local/cuda"
    if not os.path.exists(os.path.join(cuda, "include", "cuda_runtime_api.h")):
        import pytest
        pytest.skip("CUDA toolkit headers not found")
    exe = tmp_path / "c_host_layer"
    subprocess.run([gcc, "-std=c99", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(cuda, "include"),
                    os.path.join(ROOT, "examples", "c_host_layer.c"), "-L", os.path.dirname(LIB), "-lmaxk_b200",
                    "-L", os.path.join(cuda, "lib64"), "-lcudart", "-lm", "-Wl,-rpath," + os.path.dirname(LIB),
                    "-o", str(exe)], check=True, env={k: v for k, v in os.environ.items() if k not in ("CC", "CXX")})
    assert exe.exists()
