import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def kats():
    import json
    with open(os.path.join(GOLDEN, "reference_kats.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def fixture_ctx():
    with open(os.path.join(GOLDEN, "two_short_contigs.ctx"), "rb") as f:
        return f.read()


@pytest.fixture(scope="session")
def fixture_fa():
    seqs = []
    with open(os.path.join(GOLDEN, "two_short_contigs.fa")) as f:
        for line in f:
            if not line.startswith(">"):
                seqs.append(line.strip().encode())
    return seqs
