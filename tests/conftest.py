import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "bundle-adjusting-gaussian-splatting_b200")
for p in (PKG, ROOT, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # BAGS_TEST_TUNE="8=1,1=4": run the whole suite with tuning knobs set (bags_set_tuning knob=value), so that every
    # kernel variant behind a knob can be put through the same parity tests as the default
    spec = os.environ.get("BAGS_TEST_TUNE", "")
    if spec:
        import bags_capi
        for kv in spec.split(","):
            k, v = kv.split("=")
            bags_capi.check(bags_capi.lib().bags_set_tuning(int(k), int(v)), "bags_set_tuning")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
