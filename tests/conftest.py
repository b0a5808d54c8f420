import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)   # helpers.py, simt_emu/ ("tests" itself collides with an installed package)

GOLDEN = os.path.join(ROOT, "tests", "golden")
# tests/golden/ref_tests/ holds the reference's own test files as fixtures: they are run by
# tests/test_reference_unit_tests.py / test_shim_host_logic.py through the chessEngine alias, not collected directly
collect_ignore_glob = ["golden/*"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # every torch "fp32 reference" in these tests is true fp32: no TF32 in cuDNN convolutions or cuBLAS matmuls
    import torch
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
