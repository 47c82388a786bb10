import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE = "/root/reference/ga3c"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")
    config.addinivalue_line("markers", "reference: imports the reference's own Python from /root/reference "
                                       "(build container only; skipped where it is absent)")


def pytest_collection_modifyitems(config, items):
    have_ref = os.path.isdir(REFERENCE)
    for item in items:
        if "reference" in item.keywords and not have_ref:
            item.add_marker(pytest.mark.skip(reason="/root/reference not present on this box"))


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
