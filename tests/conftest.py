import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def dev():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def pytest_runtest_logstart(nodeid, location):
    # one flushed line per test start: lets tools/gpu_pytest.py name the test that hung a kernel
    if os.environ.get("MPL_TEST_TRACE"):
        sys.stderr.write(f"START {nodeid}\n")
        sys.stderr.flush()


def pytest_runtest_logreport(report):
    # immediate, flushed failure text (a later hung kernel would otherwise swallow pytest's end-of-run summary)
    if os.environ.get("MPL_TEST_TRACE") and report.failed:
        sys.stderr.write(f"\nRESULT FAILED {report.nodeid} [{report.when}]\n{report.longreprtext[-1500:]}\n")
        sys.stderr.flush()
