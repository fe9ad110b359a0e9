import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def art():
    """Shipped artefacts of the reference, repacked by tests/golden/make_golden.py."""
    return dict(np.load(os.path.join(GOLDEN, "shipped_artifacts.npz")))


@pytest.fixture(scope="session")
def tucker_golden():
    return dict(np.load(os.path.join(GOLDEN, "tucker_golden.npz")))


@pytest.fixture(scope="session")
def mlp_golden():
    return dict(np.load(os.path.join(GOLDEN, "mlp_golden.npz")))


@pytest.fixture(scope="session")
def prepost_golden():
    return dict(np.load(os.path.join(GOLDEN, "prepost_golden.npz")))


@pytest.fixture(scope="session")
def rows(art):
    """optimized_*[0:3,:] as the reference slices them (TD_Inference.py:56-57)."""
    return tuple(np.ascontiguousarray(art[f"optimized_{k}"][0:3, :]) for k in ("yaw", "pitch", "roll"))


@pytest.fixture(scope="session")
def X1k(art, rows):
    from nlml_hpe_b200 import synthetic
    return synthetic.make_features(1000, art["W"], *rows, U_id=art["U_id"], seed=1234)


@pytest.fixture(scope="session")
def state_dicts(art):
    from nlml_hpe_b200 import synthetic
    enc = synthetic.synthetic_encoder_state_dict(art["W"], art["optimized_yaw"], art["optimized_pitch"],
                                                 art["optimized_roll"], U_id=art["U_id"], seed=0)
    heads = [{k.split(".", 1)[1]: v for k, v in art.items() if k.startswith(f"{h}_network.")}
             for h in ("yaw", "pitch", "roll")]
    return enc, heads[0], heads[1], heads[2]


@pytest.fixture(scope="session")
def cuda_lib():
    """The in-tree CUDA library (built by nvcc if stale; no GPU needed to build)."""
    from nlml_hpe_b200 import _build, _lib
    _build.build()
    return _lib.load()


@pytest.fixture(scope="session")
def hostcheck():
    """Host build of the kernel arithmetic (tests/hostcheck), test infrastructure only."""
    import ctypes
    src = os.path.join(ROOT, "tests", "hostcheck", "hostcheck.cpp")
    out = os.path.join(ROOT, "tests", "hostcheck", "libhostcheck.so")
    hdr = os.path.join(ROOT, "nlml_hpe_b200", "csrc", "tucker_math.h")
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-mfma", "-ffp-contract=off", "-shared", "-fPIC", "-o", out, src])
    return ctypes.CDLL(out)
