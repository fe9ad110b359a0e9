"""Config / artefact loading with the reference's conventions (CWD-relative paths).

load_config mirrors the per-script `load_config` helpers (NLML_HPE_Test.py:171-173,
NLML_HPE_Model_Builder.py:164-166, TD_main.py:47-49): yaml.safe_load of one file.  It also
tolerates the shipped defect in configs/config_NLML_HPE_Test.yaml:29 (`val_set_path = "..."`
uses '=' instead of ':' and makes yaml.safe_load raise ScannerError).
"""
from __future__ import annotations

import re

import numpy as np
import yaml

_ASSIGN = re.compile(r"^(\s*)([A-Za-z_][A-Za-z0-9_]*)\s*=\s*(.+?)\s*$")


def load_config(path):
    with open(path, "r") as f:
        text = f.read()
    try:
        return yaml.safe_load(text)
    except yaml.YAMLError:
        fixed = []
        for line in text.splitlines():
            m = _ASSIGN.match(line)
            fixed.append(f"{m.group(1)}{m.group(2)}: {m.group(3)}" if m and not line.lstrip().startswith("#") else line)
        return yaml.safe_load("\n".join(fixed))


def load_tucker_artifacts(trained_path="./outputs/features/Trained_data.npz",
                          factors_path="./outputs/features/Factor_Matrices.npz"):
    """The arrays TD_Inference.py:40-51 reads: W, optimized_{yaw,pitch,roll}, and u_id_shape."""
    td = np.load(trained_path)
    out = {k: td[k] for k in ("W", "optimized_yaw", "optimized_pitch", "optimized_roll")}
    fm = np.load(factors_path)
    out["u_id_shape"] = int(fm["U_id"][1].size)  # TD_Inference.py:51
    for k in ("U_id", "U_yaw", "U_pitch", "U_roll"):
        out[k] = fm[k]
    return out
