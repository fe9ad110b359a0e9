"""Drop-in for the model classes of the reference's NLML_HPE_Model_Builder module, B200-backed.

Reference (/root/reference/NLML_HPE_Model_Builder.py):
  LandmarkEncoder(input_size, matrix_dims)                      :26-68
  AnglePredictionNetwork(input_size)                             :71-105
  CombinedAnglePredictionModel(encoder, yaw, pitch, roll)        :107-126
  model_builder()                                                :168-224

The classes here are torch.nn.Modules that HOLD the parameters under the reference's state_dict
keys (encoder.{0,2,4,6,8,10}.{weight,bias}; model.{0,2,4,6,8}.{weight,bias}) so `load_state_dict`
of the reference's .pth files works unchanged, but their forward does no torch math: it calls the
fused CUDA chain through the C ABI (include/nlml_hpe_b200.h).  Input must be a CUDA tensor (or a
CPU tensor / numpy array, which takes the pipelined host path); there is no CPU fallback.
"""
from __future__ import annotations

import ctypes
import warnings
from typing import List, Tuple

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .config import load_config
from .tucker import _device_index

ENCODER_HIDDEN = (1024, 512, 256, 128, 64)   # :34-49
HEAD_HIDDEN = (128, 256, 128, 64)            # :77-88


def _stack(widths, acts):
    """Sequential with Linear modules at even indices (the reference's key layout)."""
    mods = []
    for i, act in enumerate(acts):
        mods.append(nn.Linear(widths[i], widths[i + 1]))
        if act is not None:
            mods.append(act())
    return nn.Sequential(*mods)


def _linears(seq):
    return [m for m in seq if isinstance(m, nn.Linear)]


class _Plan:
    """Device-resident packed weights for one (encoder, 3 heads) set."""

    def __init__(self, linears, device_index):
        """linears: 21 nn.Linear modules, or 21 (weight, bias) tensor pairs, in the C ABI's order."""
        lib = _lib.load()
        assert len(linears) == 21
        pairs = [(l.weight, l.bias) if isinstance(l, nn.Linear) else l for l in linears]
        ws = [np.ascontiguousarray(w.detach().cpu().numpy(), dtype=np.float32) for w, _ in pairs]
        bs = [np.ascontiguousarray(b.detach().cpu().numpy(), dtype=np.float32) for _, b in pairs]
        wp = (ctypes.c_void_p * 21)(*[w.ctypes.data for w in ws])
        bp = (ctypes.c_void_p * 21)(*[b.ctypes.data for b in bs])
        outs = (ctypes.c_int * 21)(*[w.shape[0] for w in ws])
        ins = (ctypes.c_int * 21)(*[w.shape[1] for w in ws])
        h = ctypes.c_void_p()
        _lib.check(lib.nlml_mlp_plan_create(wp, bp, outs, ins, device_index, ctypes.byref(h)))
        self.lib, self.h, self.device_index = lib, h, device_index
        self.input_size, self.latent = ws[0].shape[1], ws[5].shape[0]

    def close(self):
        if self.h:
            self.lib.nlml_mlp_plan_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _check_x(x, input_size):
    if x.dtype != torch.float32 or x.dim() != 2 or x.shape[1] != input_size:
        raise ValueError(f"expected float32 [B,{input_size}], got {x.dtype} {tuple(x.shape)}")
    if x.shape[0] > 0 and x.stride(1) != 1:
        x = x.contiguous()
    return x


class LandmarkEncoder(nn.Module):
    def __init__(self, input_size, matrix_dims):
        super().__init__()
        self.input_size = input_size
        self.matrix_dims = matrix_dims
        self.total_output_size = sum(m * n for m, n in matrix_dims)
        self.encoder = _stack((input_size,) + ENCODER_HIDDEN + (self.total_output_size,),
                              (nn.ReLU, nn.ReLU, nn.ReLU, nn.ReLU, nn.Tanh, None))

    def forward(self, x):
        raise _lib.NlmlError("LandmarkEncoder alone is not on the B200 hot path; call it through "
                             "CombinedAnglePredictionModel (or .latent() of the combined model)")


class AnglePredictionNetwork(nn.Module):
    def __init__(self, input_size):
        super().__init__()
        self.input_size = input_size
        self.model = _stack((input_size,) + HEAD_HIDDEN + (1,), (nn.ReLU, nn.ReLU, nn.ReLU, nn.ReLU, None))
        for m in _linears(self.model):   # Xavier init as the reference (:94-102); overwritten by load_state_dict
            nn.init.xavier_uniform_(m.weight)
            nn.init.zeros_(m.bias)

    def forward(self, x):
        raise _lib.NlmlError("AnglePredictionNetwork alone is not on the B200 hot path; call it through "
                             "CombinedAnglePredictionModel")


class CombinedAnglePredictionModel(nn.Module):
    """forward(x[B,input_size]) -> (yaw[B,1], pitch[B,1], roll[B,1]) in radians (:115-126)."""

    def __init__(self, encoder, yaw_network, pitch_network, roll_network):
        super().__init__()
        self.encoder = encoder
        self.yaw_network = yaw_network
        self.pitch_network = pitch_network
        self.roll_network = roll_network
        self._plan = None

    def _all_linears(self):
        return (_linears(self.encoder.encoder) + _linears(self.yaw_network.model) +
                _linears(self.pitch_network.model) + _linears(self.roll_network.model))

    def invalidate(self):
        """Drop the device copy of the weights (call after changing parameters)."""
        if self._plan is not None:
            self._plan.close()
            self._plan = None

    def load_state_dict(self, *args, **kwargs):
        self.invalidate()
        return super().load_state_dict(*args, **kwargs)

    def _get_plan(self, device_index):
        if self._plan is None or self._plan.device_index != device_index:
            self.invalidate()
            self._plan = _Plan(self._all_linears(), device_index)
            _lib.check(self._plan.lib.nlml_mlp_set_path(self._plan.h, getattr(self, "_path", 0)))
        return self._plan

    @property
    def launches(self):
        return 0 if self._plan is None else int(self._plan.lib.nlml_mlp_launch_count(self._plan.h))

    def set_path(self, path, device=None):
        """'tensor_core' (default: tcgen05 chain for the wide layers) or 'fp32' (CUDA-core chain everywhere)."""
        self._path = {"tensor_core": 0, "fp32": 1}[path]
        if self._plan is not None:
            _lib.check(self._plan.lib.nlml_mlp_set_path(self._plan.h, self._path))

    def predict(self, x):
        """x CUDA float32 [B,input_size] -> CUDA [B,3] (yaw, pitch, roll) radians, asynchronous."""
        if not x.is_cuda:
            raise TypeError("predict() takes a CUDA tensor; use predict_host() for host arrays")
        plan = self._get_plan(x.device.index)
        x = _check_x(x, plan.input_size)
        out = torch.empty((x.shape[0], 3), dtype=torch.float32, device=x.device)
        ldx = x.stride(0) if x.shape[0] > 1 else x.shape[1]
        _lib.check(plan.lib.nlml_mlp_forward_f32(plan.h, x.data_ptr(), x.shape[0], ldx, out.data_ptr(),
                                                 torch.cuda.current_stream(x.device).cuda_stream))
        return out

    def predict_landmarks(self, landmarks):
        """RAW MediaPipe landmarks, CUDA float32 [B,468,3] or [B,1404] -> CUDA [B,3] radians.  The translation /
        scale normalisation of FeatureExtractor.Read_Landmarks_and_Normalizing_using_IPD
        (helpers/FeatureExtractor.py:30-66: subtract the nose tip, divide by the inter-pupillary distance) is
        fused into the first kernel's load stage, in float64 as the reference's Python floats."""
        if not landmarks.is_cuda:
            raise TypeError("predict_landmarks() takes a CUDA tensor")
        x = landmarks.reshape(landmarks.shape[0], -1)
        plan = self._get_plan(x.device.index)
        x = _check_x(x, plan.input_size)
        out = torch.empty((x.shape[0], 3), dtype=torch.float32, device=x.device)
        ldx = x.stride(0) if x.shape[0] > 1 else x.shape[1]
        _lib.check(plan.lib.nlml_mlp_forward_landmarks_f32(plan.h, x.data_ptr(), x.shape[0], ldx, out.data_ptr(),
                                                           torch.cuda.current_stream(x.device).cuda_stream))
        return out

    def predict_landmarks_host(self, landmarks, device=None):
        """predict_landmarks() for raw landmarks in HOST memory (numpy / CPU tensor, [B,468,3] or [B,1404]) -> numpy [B,3]
        radians; copies pipelined in the library."""
        if isinstance(landmarks, torch.Tensor):
            landmarks = landmarks.detach().numpy()
        x = np.ascontiguousarray(landmarks, dtype=np.float32)
        x = x.reshape(x.shape[0], -1)
        plan = self._get_plan(_device_index(device))
        if x.shape[1] != plan.input_size:
            raise ValueError(f"expected [B,{plan.input_size}] (or [B,{plan.input_size // 3},3]), got {landmarks.shape}")
        out = np.empty((x.shape[0], 3), dtype=np.float32)
        _lib.check(plan.lib.nlml_mlp_forward_landmarks_host_f32(plan.h, x.ctypes.data, x.shape[0], x.shape[1], out.ctypes.data))
        return out

    @staticmethod
    def to_degrees(angles, decimals=3, ema_alpha=None):
        """angles CUDA float32 [B,3] radians -> CUDA float64 [B,3]: round(np.degrees(t.item()), decimals) as the
        callers do (NLML_HPE_Test.py:273 decimals=3; generatePose_on_video.py:210 decimals=2) and, with ema_alpha,
        the exponential smoothing over consecutive rows (= frames) of generatePose_on_video.py:215-224 (alpha 0.4).
        Bit-identical to those Python statements."""
        if not (angles.is_cuda and angles.dtype == torch.float32 and angles.dim() == 2 and angles.shape[1] == 3):
            raise TypeError("to_degrees() takes a CUDA float32 [B,3] tensor")
        angles = angles.contiguous()
        out = torch.empty(angles.shape, dtype=torch.float64, device=angles.device)
        lib = _lib.load()
        with torch.cuda.device(angles.device):
            _lib.check(lib.nlml_pose_postprocess_f64(angles.data_ptr(), angles.shape[0], int(decimals),
                                                     float(ema_alpha) if ema_alpha else 0.0, out.data_ptr(),
                                                     torch.cuda.current_stream(angles.device).cuda_stream))
        return out

    def predict_host(self, x, device=None):
        """x numpy / CPU tensor float32 [B,input_size] -> numpy [B,3]; copies pipelined in the library."""
        if isinstance(x, torch.Tensor):
            x = x.detach().numpy()
        x = np.ascontiguousarray(x, dtype=np.float32)
        plan = self._get_plan(_device_index(device))
        if x.ndim != 2 or x.shape[1] != plan.input_size:
            raise ValueError(f"expected [B,{plan.input_size}], got {x.shape}")
        out = np.empty((x.shape[0], 3), dtype=np.float32)
        _lib.check(plan.lib.nlml_mlp_forward_host_f32(plan.h, x.ctypes.data, x.shape[0], x.shape[1], out.ctypes.data))
        return out

    def latent(self, x):
        """Encoder output [B, latent] (the tensor LandmarkEncoder.forward splits, :55-68)."""
        plan = self._get_plan(x.device.index)
        x = _check_x(x, plan.input_size)
        out = torch.empty((x.shape[0], plan.latent), dtype=torch.float32, device=x.device)
        ldx = x.stride(0) if x.shape[0] > 1 else x.shape[1]
        _lib.check(plan.lib.nlml_mlp_latent_f32(plan.h, x.data_ptr(), x.shape[0], ldx, out.data_ptr(),
                                                torch.cuda.current_stream(x.device).cuda_stream))
        return out

    def forward(self, input_landmarks):
        if input_landmarks.is_cuda:
            out = self.predict(input_landmarks)
        else:
            out = torch.from_numpy(self.predict_host(input_landmarks))
        return out[:, 0:1], out[:, 1:2], out[:, 2:3]


# ---------------------------------------------------------------------------------------------
# TorchScript form.  The reference SAVES torch.jit.script(combined_model) (:222-223) and its callers LOAD it with
# torch.jit.load("models/combined_model_scripted.pth", map_location=device).to(device).eval() (NLML_HPE_Test.py:217-219;
# generatePose_on_video.py:289).  To stay a drop-in at that boundary the archive written here is a real TorchScript
# module with the reference's parameter names whose forward is ONE custom operator, nlml_hpe_b200::combined_forward,
# implemented by the CUDA chain behind the C ABI.  The operator is registered when this package is imported, so a
# caller adds `import nlml_hpe_b200` and keeps its torch.jit.load line unchanged.
# ---------------------------------------------------------------------------------------------
_OPS = torch.library.Library("nlml_hpe_b200", "DEF")
_OPS.define("combined_forward(Tensor x, Tensor[] weights, Tensor[] biases) -> Tensor")
_OP_PLANS = {}   # (weights identity, device) -> _Plan


def _combined_forward_op(x, weights, biases):
    """x float32 [B,input_size] (CUDA: asynchronous device path; CPU: pipelined host path onto the current CUDA device)
    -> [B,3] (yaw, pitch, roll) radians on x's device.  21 weights / biases in the C ABI's order."""
    if len(weights) != 21 or len(biases) != 21:
        raise ValueError("combined_forward takes the 21 Linear layers of the encoder and the three heads")
    dev = x.device.index if x.is_cuda else _device_index(None)
    key = (tuple((w.data_ptr(), w._version) for w in weights) + tuple((b.data_ptr(), b._version) for b in biases), dev)
    plan = _OP_PLANS.get(key)
    if plan is None:
        if len(_OP_PLANS) >= 4:
            _OP_PLANS.pop(next(iter(_OP_PLANS))).close()
        plan = _OP_PLANS[key] = _Plan(list(zip(weights, biases)), dev)
    x = _check_x(x.detach(), plan.input_size)
    if x.is_cuda:
        out = torch.empty((x.shape[0], 3), dtype=torch.float32, device=x.device)
        ldx = x.stride(0) if x.shape[0] > 1 else x.shape[1]
        _lib.check(plan.lib.nlml_mlp_forward_f32(plan.h, x.data_ptr(), x.shape[0], ldx, out.data_ptr(),
                                                 torch.cuda.current_stream(x.device).cuda_stream))
        return out
    xh = np.ascontiguousarray(x.numpy())
    out = np.empty((xh.shape[0], 3), dtype=np.float32)
    _lib.check(plan.lib.nlml_mlp_forward_host_f32(plan.h, xh.ctypes.data, xh.shape[0], xh.shape[1], out.ctypes.data))
    return torch.from_numpy(out)


_OPS.impl("combined_forward", _combined_forward_op, "CompositeExplicitAutograd")


class _ParamStack(nn.Module):
    """Holds a Sequential under the attribute name the reference uses ('encoder' / 'model') so the scripted
    module's state_dict keys are the reference's; never evaluated with torch math."""

    def __init__(self, attr, seq):
        super().__init__()
        setattr(self, attr, seq)


class ScriptedCombinedAnglePredictionModel(nn.Module):
    """TorchScript-able twin of CombinedAnglePredictionModel: forward(x) -> (yaw[B,1], pitch[B,1], roll[B,1]) radians
    (:115-126) through torch.ops.nlml_hpe_b200.combined_forward."""

    def __init__(self, combined):
        super().__init__()
        self.encoder = _ParamStack("encoder", combined.encoder.encoder)
        self.yaw_network = _ParamStack("model", combined.yaw_network.model)
        self.pitch_network = _ParamStack("model", combined.pitch_network.model)
        self.roll_network = _ParamStack("model", combined.roll_network.model)

    def forward(self, input_landmarks: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        e = self.encoder.encoder
        y = self.yaw_network.model
        p = self.pitch_network.model
        r = self.roll_network.model
        ws: List[torch.Tensor] = [e[0].weight, e[2].weight, e[4].weight, e[6].weight, e[8].weight, e[10].weight,
                                  y[0].weight, y[2].weight, y[4].weight, y[6].weight, y[8].weight,
                                  p[0].weight, p[2].weight, p[4].weight, p[6].weight, p[8].weight,
                                  r[0].weight, r[2].weight, r[4].weight, r[6].weight, r[8].weight]
        bs: List[torch.Tensor] = [e[0].bias, e[2].bias, e[4].bias, e[6].bias, e[8].bias, e[10].bias,
                                  y[0].bias, y[2].bias, y[4].bias, y[6].bias, y[8].bias,
                                  p[0].bias, p[2].bias, p[4].bias, p[6].bias, p[8].bias,
                                  r[0].bias, r[2].bias, r[4].bias, r[6].bias, r[8].bias]
        out = torch.ops.nlml_hpe_b200.combined_forward(input_landmarks, ws, bs)
        return out[:, 0:1], out[:, 1:2], out[:, 2:3]


def script_combined_model(combined):
    """torch.jit.script form of a CombinedAnglePredictionModel (what model_builder saves, :222-223)."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return torch.jit.script(ScriptedCombinedAnglePredictionModel(combined).eval())


# ---------------------------------------------------------------------------------------------
# weight loading (the contract of model_builder(), :168-224)
# ---------------------------------------------------------------------------------------------
def build_combined_model(encoder_sd, yaw_sd, pitch_sd, roll_sd, input_size=None, matrix_dims=None):
    """Assemble the combined model from the four state_dicts the reference loads (:202, :214-216)."""
    to_t = lambda sd: {k: torch.as_tensor(np.asarray(v)) for k, v in sd.items()}  # noqa: E731
    encoder_sd, yaw_sd, pitch_sd, roll_sd = map(to_t, (encoder_sd, yaw_sd, pitch_sd, roll_sd))
    if input_size is None:
        input_size = encoder_sd["encoder.0.weight"].shape[1]
    heads_in = [sd["model.0.weight"].shape[1] for sd in (yaw_sd, pitch_sd, roll_sd)]
    if matrix_dims is None:
        matrix_dims = [(1, w) for w in heads_in]
    encoder = LandmarkEncoder(input_size, matrix_dims)
    encoder.load_state_dict(encoder_sd)
    nets = []
    for sd, w in zip((yaw_sd, pitch_sd, roll_sd), heads_in):
        net = AnglePredictionNetwork(w)
        net.load_state_dict(sd)
        nets.append(net)
    return CombinedAnglePredictionModel(encoder, *nets).eval()


def load_combined_model(path="models/combined_model_scripted.pth", map_location=None):
    """Replacement for `torch.jit.load("models/combined_model_scripted.pth")` (NLML_HPE_Test.py:217).

    Accepts the reference's TorchScript archive (its parameters are read out of the scripted module) or
    the plain checkpoint written by model_builder() below."""
    try:
        sd = torch.jit.load(path, map_location="cpu").state_dict()
    except Exception:
        sd = torch.load(path, map_location="cpu")
    split = {"encoder": {}, "yaw_network": {}, "pitch_network": {}, "roll_network": {}}
    for k, v in sd.items():
        top, rest = k.split(".", 1)
        split[top][rest] = v
    return build_combined_model(split["encoder"], split["yaw_network"], split["pitch_network"], split["roll_network"])


def model_builder(out_path="models/combined_model_scripted.pth"):
    """Same inputs and output as the reference's model_builder (:168-224): configs/config_EncoderTrainer.yaml,
    outputs/features/*.npz, models/{Encoder,yaw_network,pitch_network,roll_network}.pth -> a TorchScript archive at
    models/combined_model_scripted.pth (:222-223) that torch.jit.load opens (NLML_HPE_Test.py:217) once this package has
    been imported; its forward is the CUDA chain (nlml_hpe_b200::combined_forward).  Returns the eager model."""
    warnings.filterwarnings("ignore")
    config = load_config("configs/config_EncoderTrainer.yaml")
    input_size = config["input_size"]
    trained = np.load("outputs/features/Trained_data.npz")
    head_in = [trained[f"optimized_{k}"].shape[0] for k in ("yaw", "pitch", "roll")]
    factors = np.load("outputs/features/Factor_Matrices.npz")
    matrix_dims = [(1, factors[f"U_{k}"].shape[1]) for k in ("yaw", "pitch", "roll")]
    sds = [torch.load(f"models/{n}.pth", map_location="cpu") for n in ("Encoder", "yaw_network", "pitch_network", "roll_network")]
    for sd, w in zip(sds[1:], head_in):
        assert sd["model.0.weight"].shape[1] == w
    model = build_combined_model(*sds, input_size=input_size, matrix_dims=matrix_dims)
    script_combined_model(model).save(out_path)
    print("model is built")
    return model


if __name__ == "__main__":
    model_builder()
