"""Drop-in mirror of the reference's ``models/mobilenetv2.py`` (the distillation teacher / audio-tagging baseline), eval
forward only, on the B200 kernels of ``libuitk.so``.

Same constructor signature, same module tree and therefore the same ``state_dict`` keys / shapes / dtypes
(``features.N.conv...``, ``classifier.1``, the front-end buffers), same ``forward(x[B, L]) -> [B, outputdim]``: log-mel + batch-global
top-dB clamp (the shared front-end kernel), the MobileNetV2 feature extractor, mean over the mel axis, Linear + sigmoid per time
step, mean over time (mobilenetv2.py:164-178).  The sub-modules only hold parameters: ``forward`` packs them (eval BatchNorm
folded) and launches ``uitk_mnv2_forward``.  No CPU path, no PyTorch fallback; training mode and configurations other than the
default ``inverted_residual_setting`` / ``width_mult=1.0`` / ``last_channel=1280`` raise.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from .. import _native as N
from .uit import AmplitudeToDB, FrontEnd, MelSpectrogram, _Holder

__all__ = ["MobileNetV2"]

_DEFAULT_SETTING = [[1, 16, 1, 1], [6, 24, 2, 2], [6, 32, 3, 2], [6, 64, 4, 2], [6, 96, 3, 1], [6, 160, 3, 2], [6, 320, 1, 1]]   # t, c, n, s


class _Conv(_Holder):
    """Conv2d weight holder (bias=False everywhere in this network)."""

    def __init__(self, cin: int, cout: int, k: int, groups: int = 1):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(cout, cin // groups, k, k))
        nn.init.kaiming_uniform_(self.weight, a=5 ** 0.5)           # nn.Conv2d default


class _BN(_Holder):
    def __init__(self, n: int):
        super().__init__()
        self.weight, self.bias = nn.Parameter(torch.ones(n)), nn.Parameter(torch.zeros(n))
        self.register_buffer("running_mean", torch.zeros(n))
        self.register_buffer("running_var", torch.ones(n))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))


def _conv_bn_relu(cin: int, cout: int, k: int = 3, groups: int = 1) -> nn.Sequential:
    """mobilenetv2.py:8-28: Conv2d(bias=False), BatchNorm2d, ReLU6 -> keys ``0.weight``, ``1.*``."""
    return nn.Sequential(_Conv(cin, cout, k, groups), _BN(cout), nn.ReLU6(inplace=True))


class _InvertedResidual(_Holder):
    """mobilenetv2.py:30-64: [1x1 expand + BN + ReLU6] -> 3x3 depthwise + BN + ReLU6 -> 1x1 project + BN, residual if stride 1 and inp == oup."""

    def __init__(self, inp: int, oup: int, stride: int, expand_ratio: int):
        super().__init__()
        hidden = int(round(inp * expand_ratio))
        layers = [_conv_bn_relu(inp, hidden, 1)] if expand_ratio != 1 else []
        layers += [_conv_bn_relu(hidden, hidden, 3, groups=hidden), _Conv(hidden, oup, 1), _BN(oup)]
        self.conv = nn.Sequential(*layers)
        self.stride, self.use_res_connect = stride, stride == 1 and inp == oup


class MobileNetV2(nn.Module):
    def __init__(self, outputdim=527, width_mult=1.0, wavtransforms=None, spectransforms=None, inverted_residual_setting=None,
                 norm_layer=None, **kwargs):
        super().__init__()

        def need(cond, what):
            if not cond:
                raise NotImplementedError(f"uit_mobile_b200 implements the default MobileNetV2 configuration only: {what}")
        need(norm_layer is None, "custom norm_layer")
        need(float(width_mult) == 1.0, f"width_mult={width_mult}")
        need(inverted_residual_setting is None or [list(r) for r in inverted_residual_setting] == _DEFAULT_SETTING, "inverted_residual_setting")
        need(kwargs.get('last_channel', 1280) == 1280, "last_channel != 1280")
        n_mels, n_fft = kwargs.get('n_mels', 64), kwargs.get('n_fft', 512)
        hop_size, win_size, f_min = kwargs.get('hop_size', 160), kwargs.get('win_size', 512), kwargs.get('f_min', 0)
        need(n_mels == 64 and n_fft == 512 and hop_size == 160 and win_size == 512, "front-end other than n_mels=64, n_fft=win=512, hop=160")
        need(1 <= outputdim, f"outputdim={outputdim}")
        self.outputdim, self.last_channel = outputdim, 1280
        features = [_conv_bn_relu(1, 32, 3)]
        inp = 32
        for t, c, n, s in _DEFAULT_SETTING:
            for i in range(n):
                features.append(_InvertedResidual(inp, c, s if i == 0 else 1, t))
                inp = c
        features.append(_conv_bn_relu(inp, self.last_channel, 1))
        features.append(nn.AdaptiveAvgPool2d((1, None)))
        self.front_end = FrontEnd(MelSpectrogram(f_min, 8000, n_mels, n_fft), AmplitudeToDB(top_db=120))   # f_max defaults to sr / 2
        self.wavtransforms = wavtransforms if wavtransforms is not None else nn.Sequential()
        self.spectransforms = spectransforms if spectransforms is not None else nn.Sequential()
        self.features = nn.Sequential(*features)
        lin = _Holder()
        lin.weight = nn.Parameter(torch.empty(outputdim, self.last_channel))
        lin.bias = nn.Parameter(torch.zeros(outputdim))
        nn.init.kaiming_uniform_(lin.weight, a=5 ** 0.5)
        self.classifier = nn.Sequential(nn.Dropout(0.3), lin)
        self._packed = {}

    def _blob(self, device: torch.device) -> torch.Tensor:
        tensors = self.__dict__.get("_packed_tensors")
        if tensors is None:
            sd = dict(self.named_parameters())
            sd.update(dict(self.named_buffers()))
            tensors = [sd[n] for n in N.mnv2_tensor_names()]
            self.__dict__["_packed_tensors"] = tensors
        key = (str(device),) + tuple((t._version, t.data_ptr()) for t in tensors)
        if self._packed.get("key") != key:
            l = N.lib()
            host = [t.detach().to("cpu", torch.float32).contiguous() for t in tensors]
            ptrs = (C.c_void_p * len(host))(*[h.data_ptr() for h in host])
            nbytes = l.uitk_mnv2_blob_bytes(self.outputdim)
            blob = torch.zeros(nbytes, dtype=torch.uint8)
            N.check(l.uitk_pack_mnv2(self.outputdim, ptrs, blob.data_ptr(), nbytes), "uitk_pack_mnv2")
            self._packed = {"key": key, "blob": blob.to(device)}
        return self._packed["blob"]

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.training:
            raise NotImplementedError("uit_mobile_b200 implements inference only: call model.eval() "
                                      "(the training branches of mobilenetv2.py:168-172 are out of scope)")
        if not x.is_cuda:
            raise N.UitkError("MobileNetV2 runs on CUDA only (no CPU fallback); move the model and input to a B200")
        if x.dim() != 2:
            raise ValueError(f"expected a [B, L] waveform batch, got shape {tuple(x.shape)}")
        B = x.shape[0]
        out = torch.empty((B, self.outputdim), dtype=torch.float32, device=x.device)
        if B == 0:
            return out
        db = self.front_end(x)                                   # [B, 64, T] log-mel dB, batch-global top-dB clamp (Q2)
        T = db.shape[2]
        l = N.lib()
        with torch.cuda.device(x.device):
            ws = torch.empty(l.uitk_mnv2_workspace_bytes(B, T), dtype=torch.uint8, device=x.device)
            N.check(l.uitk_mnv2_forward(self.outputdim, self._blob(x.device).data_ptr(), db.data_ptr(), B, T, out.data_ptr(),
                                        ws.data_ptr(), ws.numel(), torch.cuda.current_stream(x.device).cuda_stream), "uitk_mnv2_forward")
        return out
