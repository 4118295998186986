"""-m gpu: assembled MMDiT / denoise loop / VAE decode / pipeline API against the fp32 CPU oracle."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_generate_tests(metafunc):
    if "mcheck" in metafunc.fixturenames:
        import model_checks as mc

        metafunc.parametrize("mcheck", mc.ALL_CHECKS, ids=[c.__name__ for c in mc.ALL_CHECKS])


def test_model(cuda, mcheck):
    assert isinstance(mcheck(), dict)
