import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_contact():
    from tests.helpers import load_contact_golden
    return load_contact_golden()


@pytest.fixture(scope="session")
def cuda_ctx():
    from assemble_them_all_b200 import contact
    ctx = contact.Context(0)  # raises loudly if the library or the device is missing
    yield ctx
    ctx.close()
