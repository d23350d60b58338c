import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "spgemm-prunning_b200")
for p in (ROOT, PKG, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # a fresh checkout has no built artefacts (they are git-ignored): build the C-ABI library once so that
    # the suite does not depend on `python __graft_entry__.py` having been run first (nvcc cross-compiles
    # without a GPU; this is a build step, not a fallback -- the product still raises if the .so is missing)
    import _build
    if _build.needs_build():
        try:
            _build.build()
        except Exception as ex:   # no nvcc on this machine: the pure-CPU tests (oracle, host logic) still run;
            config._maxk_build_error = repr(ex)    # tests that load the library fail with the ImportError it raises


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
