import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")
    # the oracle port is test infrastructure: build it on demand (gcc, <1 s)
    if not os.path.exists(os.path.join(ROOT, "oracle", "libgrab_oracle.so")):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "port"], check=True,
                       stdout=subprocess.DEVNULL)


def pytest_collection_modifyitems(config, items):
    # a pathological random pattern must fail its test, not hold the suite (pytest-timeout is installed in this image)
    if config.pluginmanager.hasplugin("timeout"):
        for it in items:
            if not any(m.name == "timeout" for m in it.iter_markers()):
                it.add_marker(pytest.mark.timeout(600))
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)
