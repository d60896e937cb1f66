import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "live_reference: needs /root/reference (build container only)")


def pytest_collection_modifyitems(config, items):
    have_ref = os.path.isfile("/root/reference/surrDAMH/modules/classes_SAMPLER.py")
    skip_ref = pytest.mark.skip(reason="/root/reference not present on this machine")
    for item in items:
        if "live_reference" in item.keywords and not have_ref:
            item.add_marker(skip_ref)
