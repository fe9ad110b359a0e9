"""The reference's entry-point FUNCTIONS executed as entry points, in a working directory laid out like the reference
checkout (conftest.ref_layout): model_builder (NLML_HPE_Model_Builder.py:168-224), the torch.jit.load line of
NLML_HPE_Test.py:217-219 with the batch-1 loop of :262-273, NLML_HPE_Tester, TD_Inference.inference (TD_Inference.py:20-76)."""
import numpy as np
import pytest
import torch

import nlml_hpe_b200  # noqa: F401  (registers the TorchScript operator the scripted archive calls)
from nlml_hpe_b200 import NLML_HPE_Model_Builder as MB
from nlml_hpe_b200 import NLML_HPE_Test, TD_Inference, TD_Tester

DEG = 180.0 / np.pi


def test_model_builder_writes_a_torchscript_archive(ref_layout, state_dicts, capsys):
    """model_builder() reads the same four .pth / two .npz / one .yaml and writes models/combined_model_scripted.pth
    (:222-223); the file opens with torch.jit.load as the reference's callers do (NLML_HPE_Test.py:217) and holds the
    parameters under the reference's names."""
    MB.model_builder()
    assert "model is built" in capsys.readouterr().out                # :224
    model = torch.jit.load("models/combined_model_scripted.pth", map_location="cpu").eval()
    sd = model.state_dict()
    names = ("encoder.encoder", "yaw_network.model", "pitch_network.model", "roll_network.model")
    for prefix, ref in zip(names, state_dicts):
        for k, v in ref.items():
            key = f"{prefix}.{k.split('.', 1)[1]}" if prefix.startswith("encoder") else f"{prefix}.{k.split('.', 1)[1]}"
            assert torch.equal(sd[key], torch.as_tensor(np.asarray(v))), key
    assert len(sd) == 42
    # the archive also loads through the package's own loader (state_dict route)
    eager = MB.load_combined_model("models/combined_model_scripted.pth")
    assert sum(p.numel() for p in eager.parameters()) == 2136585 + 3 * 74753
    if not torch.cuda.is_available():
        with pytest.raises(Exception):                               # no CPU fallback: the operator fails loudly
            model(torch.zeros(2, 1404))


@pytest.mark.gpu
def test_reference_call_site_with_torch_jit_load(ref_layout, X1k, mlp_golden, cuda_lib):
    """The reference's own statements (NLML_HPE_Test.py:185, :217-219, :262, :271-273) on the archive model_builder wrote."""
    MB.model_builder()
    device = torch.device("cuda" if torch.cuda.is_available() else "cpu")                                   # :185
    model = torch.jit.load("models/combined_model_scripted.pth", map_location=device).to(device).eval()      # :217-219
    pred = []
    for i in range(64):
        x = torch.from_numpy(X1k[i])
        x = x.unsqueeze(0).float().to(device)                                                                # :262
        with torch.no_grad():
            yaw_pred, pitch_pred, roll_pred = model(x)                                                       # :271-272
        pred.append((round(np.degrees(yaw_pred.item()), 3), round(np.degrees(pitch_pred.item()), 3),
                     round(np.degrees(roll_pred.item()), 3)))                                                # :273
    ref = np.degrees(mlp_golden["angles_b1"].astype(np.float64))       # the reference's batch-1 loop on the same rows
    assert np.abs(np.array(pred) - ref).max() < 1e-3 + 5e-4             # 1e-3 degrees + the 3-decimal rounding
    # batched call through the same archive, CPU input (host pipeline)
    with torch.no_grad():
        y, p, r = model(torch.from_numpy(X1k).to(device))
    out = torch.cat([y, p, r], 1).cpu().numpy()
    assert np.abs(out - mlp_golden["angles_jit"]).max() * DEG < 1e-3


@pytest.mark.gpu
def test_NLML_HPE_Tester_driver(ref_layout, X1k, mlp_golden, capsys, cuda_lib):
    """NLML_HPE_Tester (the reference driver :182-449 minus image decoding): both yaml files incl. the '=' line, the
    scripted archive, the all-zero "no face" rows skipped (:257-260), degrees rounded to 3 decimals (:273), MAE lines."""
    MB.model_builder()
    X = X1k[:300].copy()
    X[[7, 123]] = 0.0
    true = np.degrees(mlp_golden["angles_jit"][:300].astype(np.float64)) + 0.25
    np.savez("features.npz", X=X, angles=true)
    pred = NLML_HPE_Test.NLML_HPE_Tester(["--features_npz", "features.npz"])
    out = capsys.readouterr().out
    assert "processed 298 samples, 2 without landmarks; val_set = facescape" in out
    assert "MAE  yaw/pitch/roll/mean" in out
    keep = np.ones(300, bool)
    keep[[7, 123]] = False
    ref = np.degrees(mlp_golden["angles_jit"][:300][keep].astype(np.float64))
    assert len(pred) == 298 and all(isinstance(t, tuple) and len(t) == 3 for t in pred)
    assert np.abs(np.array(pred) - ref).max() < 1e-3 + 5e-4
    mae = [float(v) for v in out.split("MAE  yaw/pitch/roll/mean =")[1].split("\n")[0].split()]
    assert all(abs(m - 0.25) < 2e-3 for m in mae)


@pytest.mark.gpu
def test_TD_Inference_cli(ref_layout, X1k, tucker_golden, capsys, cuda_lib, monkeypatch):
    """TD_Inference.inference: loads ./outputs/features/*.npz (:40-51), slices the cosine rows [0:3] (:56-57), calls
    Test, prints the three lines of :65-67."""
    np.save("x0.npy", X1k[0])
    monkeypatch.setattr(TD_Tester, "TEST_SOLVER", "sgd")   # the fixed-iteration block (:168-184): bit-level parity contract
    y, p, r = TD_Inference.inference(["--features_npy", "x0.npy"])
    out = capsys.readouterr().out.strip().splitlines()
    ref = np.degrees(tucker_golden["sgd3000_shipped_P"][0, :3].astype(np.float64))
    assert max(abs(y - ref[0]), abs(p - ref[1]), abs(r - ref[2])) < 1e-2
    assert out == [f"Estimated yaw in degree = {y:.2f}", f"Estimated pitch in degree = {p:.2f}",
                   f"Estimated roll in degree = {r:.2f}"]
    monkeypatch.setattr(TD_Tester, "TEST_SOLVER", "converged")   # the fast alternative: same basin as the reference's Powell
    y, p, r = TD_Inference.inference(["--features_npy", "x0.npy"])
    out = capsys.readouterr().out
    assert out.count("Estimated") == 3
    pw = tucker_golden["powell_shipped_deg"][0]
    assert max(abs(y - pw[0]), abs(p - pw[1]), abs(r - pw[2])) < 5.0       # see test_tucker_gpu.test_solve_vs_reference_powell_96
    monkeypatch.setattr(TD_Tester, "TEST_SOLVER", "powell")      # the default: the reference's own search, same printed lines
    y, p, r = TD_Inference.inference(["--features_npy", "x0.npy"])
    out = capsys.readouterr().out.strip().splitlines()
    assert max(abs(y - pw[0]), abs(p - pw[1]), abs(r - pw[2])) < 1e-6
    assert out[-3:] == [f"Estimated yaw in degree = {y:.2f}", f"Estimated pitch in degree = {p:.2f}",
                        f"Estimated roll in degree = {r:.2f}"]
    with pytest.raises(SystemExit):
        TD_Inference.inference([])                                          # argparse: one of the two inputs is required
