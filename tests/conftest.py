import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "live_reference: needs the reference modules (/root/reference, or oracle/_ref made by oracle.make_ref)")


def pytest_collection_modifyitems(config, items):
    from oracle import ref_import
    have_ref = ref_import.available()
    skip_ref = pytest.mark.skip(reason="reference modules not present on this machine (/root/reference or oracle/_ref)")
    for item in items:
        if "live_reference" in item.keywords and not have_ref:
            item.add_marker(skip_ref)
