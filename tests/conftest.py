import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')
    config.addinivalue_line('markers', 'refonly: needs /root/reference (build container only)')


def pytest_collection_modifyitems(config, items):
    import torch
    has_gpu = torch.cuda.is_available()
    has_ref = os.path.isdir(os.path.join(os.environ.get('GGA_REFERENCE_ROOT', '/root/reference'), 'mmdet3d'))
    for it in items:
        if 'gpu' in it.keywords and not has_gpu:
            it.add_marker(pytest.mark.skip(reason='no CUDA device'))
        if 'refonly' in it.keywords and not has_ref:
            it.add_marker(pytest.mark.skip(reason='reference tree not present (GGA_REFERENCE_ROOT)'))


@pytest.fixture(scope='session')
def golden_dir():
    return os.path.join(ROOT, 'tests', 'golden')
