import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _has_gpu() -> bool:
    try:
        from chebpol_b200 import _lib

        return _lib.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a box without a GPU must fail loudly, not skip: the product has no fallback.
    pass


@pytest.fixture(scope="session")
def port():
    from oracle.oracle import Oracle

    return Oracle("port")


@pytest.fixture(scope="session")
def ref():
    from oracle.oracle import Oracle

    if not Oracle.available("reference"):
        try:
            from oracle import oracle as o

            o.build()
        except Exception:
            pass
    if not Oracle.available("reference"):
        pytest.skip("oracle/_ref/libchebpol_ref.so not present (reference sources unavailable here)")
    return Oracle("reference")


@pytest.fixture(scope="session")
def gpu_ok():
    assert _has_gpu(), "no CUDA device visible: chebpol_b200 has no CPU fallback, gpu tests cannot run"
    return True
