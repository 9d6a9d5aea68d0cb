import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run with -m gpu on the B200 box)')
    config.addinivalue_line('markers', 'ref_tree: needs the reference tree mounted at /root/reference')


def pytest_collection_modifyitems(config, items):
    from oracle import ref_shim
    if not ref_shim.reference_available():
        skip = pytest.mark.skip(reason='reference tree not mounted')
        for item in items:
            if 'ref_tree' in item.keywords:
                item.add_marker(skip)
