"""-m gpu: every hand-written kernel against an fp32 torch reference of the same op (through the C ABI)."""
import pytest

pytestmark = pytest.mark.gpu


def _checks():
    import importlib

    kc = importlib.import_module("kernel_checks")
    return kc.ALL_CHECKS


def pytest_generate_tests(metafunc):
    if "check" in metafunc.fixturenames:
        import os
        import sys

        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        import kernel_checks as kc

        metafunc.parametrize("check", kc.ALL_CHECKS, ids=[c.__name__ for c in kc.ALL_CHECKS])


def test_kernel(cuda, check):
    res = check()
    assert isinstance(res, dict)
