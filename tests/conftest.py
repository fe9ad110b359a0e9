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
def powell_golden():
    """96 outputs of the reference's Test() = scipy Powell (tests/golden/make_golden_powell.py)."""
    return dict(np.load(os.path.join(GOLDEN, "powell_golden.npz")))


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


ENCODER_YAML = """yaw_bins:
  min_bin: -50
  max_bin:  51
  interval: 10

pitch_bins:
  min_bin: -40
  max_bin:  41
  interval: 10

roll_bins:
  min_bin: -30
  max_bin:  31
  interval: 10

input_size: 1404   # number of inputs of the encoder
batch_size: 256
"""

# the shipped file's defect is reproduced on purpose: `val_set_path = "..."` ('=' instead of ':', config_NLML_HPE_Test.yaml:29)
TEST_YAML = """yaw_intervals:
  - [-51, -33.33]
  - [33.33, 51]

val_set : "facescape" # facescape / biwi / AFLW2000
# val_set_path: "C:\\\\somewhere\\\\BIWI.npz"
val_set_path = "E:/Mahdi/Databases/some_valset_(+y+p+r)_rotation_convention(jpg)" 
"""


@pytest.fixture()
def ref_layout(tmp_path, art, state_dicts, monkeypatch):
    """A working directory laid out like a checkout of the reference: configs/*.yaml, outputs/features/*.npz,
    models/*.pth (shipped heads + the synthetic stand-in for the missing Encoder.pth).  The entry points read
    CWD-relative paths (TD_Inference.py:40,49; NLML_HPE_Model_Builder.py:174-216; NLML_HPE_Test.py:188,200,217)."""
    import torch
    (tmp_path / "configs").mkdir()
    (tmp_path / "configs" / "config_EncoderTrainer.yaml").write_text(ENCODER_YAML)
    (tmp_path / "configs" / "config_NLML_HPE_Test.yaml").write_text(TEST_YAML)
    feat = tmp_path / "outputs" / "features"
    feat.mkdir(parents=True)
    np.savez(feat / "Trained_data.npz", W=art["W"], CoreTensor=np.zeros((1,), np.float32),
             optimized_yaw=art["optimized_yaw"], optimized_pitch=art["optimized_pitch"], optimized_roll=art["optimized_roll"])
    np.savez(feat / "Factor_Matrices.npz", U_id=art["U_id"], U_yaw=art["U_yaw"], U_pitch=art["U_pitch"], U_roll=art["U_roll"])
    (tmp_path / "models").mkdir()
    for name, sd in zip(("Encoder", "yaw_network", "pitch_network", "roll_network"), state_dicts):
        torch.save({k: torch.as_tensor(np.asarray(v)) for k, v in sd.items()}, tmp_path / "models" / f"{name}.pth")
    monkeypatch.chdir(tmp_path)
    return tmp_path
