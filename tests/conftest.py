import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def native_lib():
    """Build (if stale) and load libkassign.so. Never falls back to anything else."""
    import kafka_assigner_b200 as kab
    kab.build_mod.build()
    return kab.load_native()


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle_lib
    oracle_lib.build()
    return oracle_lib
