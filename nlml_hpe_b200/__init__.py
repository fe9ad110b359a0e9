"""nlml_hpe_b200 -- B200-native (sm_100a) inference hot path of NLML_HPE.

Two halves, both behind the reference's own Python entry points:
  TD_Tester / TD_Inference            batched fixed-iteration Tucker-fit pose inversion
  NLML_HPE_Model_Builder / _Test      Encoder + yaw/pitch/roll MLP-heads forward
Host code is Python; compute is hand-written CUDA reached through the C ABI in
include/nlml_hpe_b200.h (libnlml_hpe_b200.so, built in-tree).  No CPU fallback.
"""
__version__ = "0.1.0"

# importing the package registers the TorchScript operator nlml_hpe_b200::combined_forward, which the archive written by
# NLML_HPE_Model_Builder.model_builder() calls (so `import nlml_hpe_b200` + the reference's torch.jit.load line suffice)
from . import NLML_HPE_Model_Builder as _model_builder  # noqa: E402,F401
