import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _device_count():
    try:
        from bild_b200 import _lib
        return int(_lib.load().bildk_device_count())
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    """`gpu`-marked tests are skipped (not failed) when no CUDA device is visible or the library is not built."""
    if not any("gpu" in item.keywords for item in items):
        return
    if _device_count() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device visible (bildk_device_count() == 0)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
