"""Drop-in mirror of the reference's ``models/uit.py`` for the batched-inference hot path on B200.

Same factory names (``uit_xs / uit_xxs / uit_xxxs``), same constructor kwargs, same ``state_dict`` keys / shapes /
dtypes, same ``forward(x[B, L]) -> [B, outputdim]`` sigmoid scores (reference: models/uit.py:252-493, 581-655).
The arithmetic is NOT here: ``forward`` calls the hand-written sm_100a kernels in ``libuitk.so`` through the C ABI
declared in ``include/uitk.h``.  The sub-modules below are parameter holders that keep the reference's
state_dict layout; calling them directly raises.  There is no CPU path and no PyTorch fallback: a model that is
not on a CUDA device, is in training mode, or is configured outside what the kernels implement raises.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Optional

import torch
import torch.nn as nn

from .. import _native as N

__all__ = ["UITBase", "uit_xs", "uit_xxs", "uit_xxxs", "PRETRAINED_CHECKPOINTS", "audio_transformer_h128_d4_m3_relu",
           "audio_transformer_h128_d4_m3", "audio_transformer_h128_d6_m3", "audio_transformer_h128_d6_m3_relu",
           "audio_transformer_h128_d3_m3_bneck_v2_relu", "AudioPatchEmbed", "BNeckAttention", "Attention", "Mlp", "Block"]


class _Holder(nn.Module):
    """Parameter container: arithmetic lives in the fused CUDA kernels."""

    def forward(self, *a, **k):  # pragma: no cover - guard
        raise RuntimeError(f"{type(self).__name__} only holds parameters; run the whole model (UITBase.forward), "
                           "the hot path is fused CUDA with no per-layer PyTorch fallback")


class _Linear(_Holder):
    def __init__(self, in_features: int, out_features: int):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.weight = nn.Parameter(torch.empty(out_features, in_features))
        self.bias = nn.Parameter(torch.zeros(out_features))
        nn.init.trunc_normal_(self.weight, std=.02)          # uit.py:369-373


class _LayerNorm(_Holder):
    def __init__(self, dim: int, eps: float):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(dim))
        self.bias = nn.Parameter(torch.zeros(dim))


class _BatchNorm(_Holder):
    """Eval-mode BatchNorm2d(64, momentum=0.01) state (uit.py:310-313)."""

    def __init__(self, n: int):
        super().__init__()
        self.eps, self.momentum = 1e-5, 0.01
        self.weight = nn.Parameter(torch.ones(n))
        self.bias = nn.Parameter(torch.zeros(n))
        self.register_buffer("running_mean", torch.zeros(n))
        self.register_buffer("running_var", torch.ones(n))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))


class _InitBN(nn.Sequential):
    """``init_bn`` of the reference (uit.py:310-313): Rearrange -> eval BatchNorm2d over the mel axis -> Rearrange.  Callable like
    the reference's ([B, 1, 64, T] -> same shape) through the CUDA kernel; inside ``UITBase.forward`` it is fused into the
    encoder's patch gather."""

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        owner = self.__dict__.get("_owner")
        if owner is None or owner() is None:
            raise RuntimeError("init_bn is bound to its UITBase model")
        return owner()._init_bn(x)


class _Conv(_Holder):
    def __init__(self, out_ch: int, k: int):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(out_ch, 1, k, k))
        self.bias = nn.Parameter(torch.empty(out_ch))
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))   # nn.Conv2d.reset_parameters
        bound = 1.0 / math.sqrt(k * k)
        nn.init.uniform_(self.bias, -bound, bound)


class AudioPatchEmbed(_Holder):
    """uit.py:43-74 (flatten=False, norm=Identity)."""

    def __init__(self, input_size, patch_size: int, patch_stride: int, embed_dim: int):
        super().__init__()
        self.input_size = tuple(input_size)
        self.patch_size = (patch_size, patch_size)
        self.grid_size = (input_size[0] // patch_stride, input_size[1] // patch_stride)
        self.num_patches = self.grid_size[0] * self.grid_size[1]
        self.flatten = False
        self.proj = _Conv(embed_dim, patch_size)
        self.norm = nn.Identity()


class BNeckAttention(_Holder):
    """uit.py:89-122."""

    def __init__(self, dim: int, num_heads: int):
        super().__init__()
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5          # 0.125: from the UN-bottlenecked head dim (Q3)
        self.inner_dim = dim // 4
        self.qkv = _Linear(dim, self.inner_dim * 3)
        self.proj = _Linear(self.inner_dim, dim)


class Attention(_Holder):
    """uit.py:124-178 (causal=False): qkv dim -> 3 dim, heads of dim // num_heads."""

    def __init__(self, dim: int, num_heads: int):
        super().__init__()
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.causal = False
        self.qkv = _Linear(dim, dim * 3)
        self.proj = _Linear(dim, dim)


class Mlp(_Holder):
    def __init__(self, dim: int, hidden: int, act_layer=nn.ReLU):
        super().__init__()
        self.fc1 = _Linear(dim, hidden)
        self.act = act_layer()
        self.fc2 = _Linear(hidden, dim)


class Block(_Holder):
    """uit.py:206-248 with LayerScale / DropPath = Identity."""

    def __init__(self, dim: int, num_heads: int, mlp_ratio: float, act_layer=nn.ReLU, attention_type: str = "BNeckAttention"):
        super().__init__()
        self.norm1 = _LayerNorm(dim, 1e-6)
        self.attn = {"BNeckAttention": BNeckAttention, "Attention": Attention}[attention_type](dim, num_heads)
        self.norm2 = _LayerNorm(dim, 1e-6)
        self.mlp = Mlp(dim, int(dim * mlp_ratio), act_layer)


class _Spectrogram(_Holder):
    def __init__(self, n_fft: int):
        super().__init__()
        self.register_buffer("window", torch.hann_window(n_fft, periodic=True))


def _melscale_fbanks_htk(n_freqs: int, f_min: float, f_max: float, n_mels: int, sample_rate: int) -> torch.Tensor:
    """HTK triangular filterbank, torchaudio.functional.melscale_fbanks(norm=None, mel_scale='htk')."""
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_min = 2595.0 * math.log10(1.0 + f_min / 700.0)
    m_max = 2595.0 * math.log10(1.0 + f_max / 700.0)
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.max(torch.zeros(1), torch.min(down, up))


class _MelScale(_Holder):
    def __init__(self, f_min, f_max, n_mels, n_fft):
        super().__init__()
        self.register_buffer("fb", _melscale_fbanks_htk(n_fft // 2 + 1, float(f_min), float(f_max), n_mels, 16000))


class MelSpectrogram(_Holder):
    """Holds ``spectrogram.window`` / ``mel_scale.fb`` (persistent buffers of the reference's front-end, Q9)."""

    def __init__(self, f_min, f_max, n_mels, n_fft):
        super().__init__()
        self.spectrogram = _Spectrogram(n_fft)
        self.mel_scale = _MelScale(f_min, f_max, n_mels, n_fft)


class AmplitudeToDB(_Holder):
    def __init__(self, top_db: float):
        super().__init__()
        self.top_db = top_db


class FrontEnd(nn.Sequential):
    """``front_end`` of the reference (uit.py:298-308): log-mel in dB with the batch-global top-dB clamp,
    [B, L] -> [B, 64, T], computed by the fused CUDA kernel."""

    def new_words(self, device) -> torch.Tensor:
        device = torch.device(device)
        cache = self.__dict__.setdefault("_words_init", {})
        if device not in cache:
            cache[device] = torch.tensor([0, 0x7F800000], dtype=torch.int32, device=device)
        return cache[device].clone()

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda:
            raise N.UitkError("UiT hot path runs on CUDA only (no CPU fallback); move the model and input to a B200")
        words = self.new_words(x.device)                                                # [max power bits, min power bits]
        db, _ = self.logmel_unclamped(x, max_pow=words[0:1], min_pow=words[1:2])
        return self.clamp_(db, words[0:1], words[1:2])

    def clamp_(self, db: torch.Tensor, max_pow: torch.Tensor, min_pow: Optional[torch.Tensor] = None) -> torch.Tensor:
        """In-place top-dB clamp (AmplitudeToDB(top_db=120), one cutoff for the whole batch: Q2).  With the batch's minimum
        power word the pass decides on the device whether any value lies under the cutoff and returns at once if none does
        (true for any realistic signal: the 4 bytes/value read-modify-write pass is then free)."""
        with torch.cuda.device(db.device):
            N.check(N.lib().uitk_clamp_db(db.data_ptr(), db.numel(), max_pow.data_ptr(), None if min_pow is None else min_pow.data_ptr(),
                                          float(self[1].top_db), torch.cuda.current_stream(db.device).cuda_stream), "uitk_clamp_db")
        return db

    def _blob(self, device: torch.device) -> torch.Tensor:
        win, fb = self[0].spectrogram.window, self[0].mel_scale.fb
        key = (str(device), win._version, fb._version, win.data_ptr(), fb.data_ptr())
        cache = self.__dict__.setdefault("_uitk_cache", {})
        if cache.get("key") != key:
            l = N.lib()
            w = win.detach().to("cpu", torch.float32).contiguous()
            f = fb.detach().to("cpu", torch.float32).contiguous()
            if tuple(w.shape) != (512,) or tuple(f.shape) != (257, 64):
                raise N.UitkError("the log-mel kernel implements n_fft=win=512, hop=160, 64 mels only")
            host = torch.zeros(l.uitk_frontend_blob_bytes(f.data_ptr()), dtype=torch.uint8)
            N.check(l.uitk_pack_frontend(w.data_ptr(), f.data_ptr(), host.data_ptr(), host.numel()), "uitk_pack_frontend")
            cache["key"], cache["blob"] = key, host.to(device)
        return cache["blob"]

    def logmel_unclamped(self, x: torch.Tensor, ld: Optional[int] = None, B: Optional[int] = None, L: Optional[int] = None,
                         out: Optional[torch.Tensor] = None, max_pow: Optional[torch.Tensor] = None,
                         min_pow: Optional[torch.Tensor] = None):
        """Launch K1.  Returns (dB [B,64,T] un-clamped, max-power word [1] int32).  ``ld/B/L`` describe strided
        views (sliding windows over one long stream) without materialising them.  ``out`` / ``max_pow`` let a
        pipelined caller write chunks of one batch into a shared buffer and keep ONE running maximum (Q2)."""
        if not x.is_cuda:
            raise N.UitkError("UiT hot path runs on CUDA only (no CPU fallback); move the model and input to a B200")
        if x.dtype not in (torch.float32, torch.int16):
            raise N.UitkError(f"expected a float32 waveform (or int16 PCM), got {x.dtype}")
        if B is None:
            if x.dim() != 2:
                raise ValueError(f"expected a [B, L] waveform batch, got shape {tuple(x.shape)}")
            if x.stride(1) != 1 or (x.shape[0] > 1 and x.stride(0) < 1):
                x = x.contiguous()
            B, L = x.shape
            ld = x.stride(0) if B > 1 else L
        l = N.lib()
        T = int(l.uitk_num_frames(L))
        if out is None:
            out = torch.empty((B, 64, T), dtype=torch.float32, device=x.device)
        elif tuple(out.shape) != (B, 64, T) or not out.is_contiguous() or out.dtype != torch.float32:
            raise ValueError("out must be a contiguous float32 [B, 64, T] tensor")
        if max_pow is None:
            max_pow = torch.zeros(1, dtype=torch.int32, device=x.device)
        with torch.cuda.device(x.device):
            fn = l.uitk_logmel_i16 if x.dtype == torch.int16 else l.uitk_logmel     # PCM: x = pcm / 32768 in-kernel
            N.check(fn(x.data_ptr(), B, L, ld, self._blob(x.device).data_ptr(), out.data_ptr(),
                       max_pow.data_ptr(), None if min_pow is None else min_pow.data_ptr(),
                       torch.cuda.current_stream(x.device).cuda_stream), "uitk_logmel")
        return out, max_pow


def _logmel_sliding(front_end, stream: torch.Tensor, window: int = 16000, hop: int = 1600,
                    max_pow: Optional[torch.Tensor] = None, min_pow: Optional[torch.Tensor] = None):
    """Log-mel of all sliding windows of ONE long stream.  Same result, bit for bit, as
    ``logmel_unclamped(stream, ld=hop, B=W, L=window)``; when ``hop`` is a multiple of 160 the interior STFT frames are
    computed once for the stream instead of once per overlapping window (10x less front-end work at hop 1600)."""
    if not stream.is_cuda or stream.dtype != torch.float32 or stream.dim() != 1 or not stream.is_contiguous():
        raise N.UitkError("expected a contiguous 1-D float32 CUDA tensor (the stream)")
    n = stream.numel()
    if n < window:
        raise ValueError("stream shorter than one window")
    W = (n - window) // hop + 1
    if hop % 160 != 0 or window % 160 != 0 or hop // 160 > window // 160 - 3:
        # no shared interior frames (hop not on the STFT grid), or windows so far apart that some stream frames belong to no
        # window (they would still raise the batch-global top-dB maximum): every window runs its own front-end, read in place
        return front_end.logmel_unclamped(stream, ld=hop, B=W, L=window, max_pow=max_pow, min_pow=min_pow)
    l = N.lib()
    T = int(l.uitk_num_frames(window))
    out = torch.empty((W, 64, T), dtype=torch.float32, device=stream.device)
    if max_pow is None:
        max_pow = torch.zeros(1, dtype=torch.int32, device=stream.device)
    ws = torch.empty(int(l.uitk_logmel_sliding_workspace_bytes(n)), dtype=torch.uint8, device=stream.device)
    with torch.cuda.device(stream.device):
        N.check(l.uitk_logmel_sliding(stream.data_ptr(), n, window, hop, front_end._blob(stream.device).data_ptr(), out.data_ptr(),
                                      max_pow.data_ptr(), None if min_pow is None else min_pow.data_ptr(), ws.data_ptr(), ws.numel(),
                                      torch.cuda.current_stream(stream.device).cuda_stream), "uitk_logmel_sliding")
    return out, max_pow


FrontEnd.logmel_sliding = _logmel_sliding


class UITBase(nn.Module):
    """uit.py:252-493, inference path only (eval mode, init_bn).  The UiT-XS/XXS/XXXS configuration (BNeckAttention, ReLU MLP,
    pooling='mean') runs on the tcgen05 megakernel when ``precision='bf16'``; the other variants of the class (full
    ``Attention``, GELU, pooling='token' | 'dm') run on the fp32 CUDA-core kernels whatever ``precision`` says."""

    def __init__(self, outputdim=527, patch_size=16, patch_stride=16, embed_dim=768, depth=12, num_heads=12,
                 mlp_ratio=4., qkv_bias=True, drop_rate=0., attn_drop_rate=0., drop_path_rate=0., init_bn: bool = True,
                 norm_layer=None, act_layer=None, init_values=None, target_length=1012, pooling='token',
                 wavtransforms=None, spectransforms=None, time_patch_out: Optional[float] = None,
                 freq_patch_out: Optional[float] = None, block_type='Block', attention_type='Attention',
                 eval_avg='mean', precision: str = "bf16", process_group=None, **kwargs):
        super().__init__()
        assert pooling in ('mean', 'token', 'dm')
        self.outputdim, self.pooling, self.embed_dim = outputdim, pooling, embed_dim
        self.patch_stride, self.patch_size = patch_stride, patch_size
        self.n_mels = kwargs.get('n_mels', 64)
        n_fft = kwargs.get('n_fft', 512)
        self.hop_size = kwargs.get('hop_size', 160)
        self.win_size = kwargs.get('win_size', 512)
        f_min, f_max = kwargs.get('f_min', 0), kwargs.get('f_max', 8000)
        self.center = kwargs.get('center', True)
        self.eval_avg = eval_avg
        self.time_patch_out, self.freq_patch_out = time_patch_out, freq_patch_out
        self.target_length = target_length
        self.depth, self.num_heads, self.mlp_ratio = depth, num_heads, mlp_ratio
        self.precision = precision
        self.process_group = process_group          # batch-sharded inference: all-reduce(max) of the top-dB scope
        self.peer_words = None                      # optional sharding.PeerWords: the same exchange through NVLink peer memory
        self.max_clips_per_launch = 16384

        # Everything the kernels do not implement is refused up front (no silent fallback).
        def need(cond, what):
            if not cond:
                raise NotImplementedError(f"uit_mobile_b200 implements the UiT-XS/XXS/XXXS inference path only: {what}")
        if attention_type not in ('BNeckAttention', 'Attention'):
            raise KeyError(attention_type)           # the reference looks the class up in globals() (uit.py:224): same error
        need(block_type == 'Block', f"block_type={block_type!r}")
        act_layer = act_layer or nn.GELU             # uit.py:338
        need(act_layer in (nn.ReLU, nn.GELU), "act_layer must be nn.ReLU or nn.GELU")
        self.attention_type, self.act_layer = attention_type, act_layer
        need(init_bn, "init_bn=False")
        need(embed_dim == 128 and num_heads == 2 and float(mlp_ratio) == 3.0, "embed_dim/num_heads/mlp_ratio != 128/2/3.0")
        need(patch_size == 16 and patch_stride == 16, "patch_size/stride != 16")
        need(self.n_mels == 64 and n_fft == 512 and self.hop_size == 160 and self.win_size == 512 and self.center,
             "front-end other than n_mels=64, n_fft=win=512, hop=160, center=True")
        need(init_values is None and drop_rate == 0. and attn_drop_rate == 0. and drop_path_rate == 0., "LayerScale/dropout/drop-path")
        need(qkv_bias, "qkv_bias=False")
        need(norm_layer is None, "custom norm_layer")
        need(16 <= target_length <= 111, f"target_length={target_length} (16..111 frames, i.e. at most 6 time patches)")
        need(1 <= outputdim <= 768, f"outputdim={outputdim}")
        need(eval_avg in ('mean', 'max'), f"eval_avg={eval_avg!r}")
        need(precision in N.PRECISIONS, f"precision={precision!r} (fp32 or bf16)")

        self.front_end = FrontEnd(MelSpectrogram(f_min, f_max, self.n_mels, n_fft), AmplitudeToDB(top_db=120))
        self.init_bn = _InitBN(nn.Identity(), _BatchNorm(self.n_mels), nn.Identity())
        import weakref
        self.init_bn.__dict__["_owner"] = weakref.ref(self)
        self.patch_embed = AudioPatchEmbed((self.n_mels, target_length), patch_size, patch_stride, embed_dim)
        self.spectransforms = nn.Sequential() if spectransforms is None else spectransforms
        self.wavtransforms = nn.Sequential() if wavtransforms is None else wavtransforms
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.token_pos_embed = nn.Parameter(torch.randn(1, embed_dim) * .02)
        self.time_pos_embed = nn.Parameter(torch.randn(1, embed_dim, 1, self.patch_embed.grid_size[1]) * .02)
        self.freq_pos_embed = nn.Parameter(torch.randn(1, embed_dim, self.patch_embed.grid_size[0], 1) * .02)
        self.pos_drop = nn.Identity()
        self.blocks = nn.Sequential(*[Block(embed_dim, num_heads, mlp_ratio, act_layer, attention_type) for _ in range(depth)])
        self.norm = _LayerNorm(embed_dim, 1e-6)
        self.outputlayer = nn.Sequential(_LayerNorm(embed_dim, 1e-5), _Linear(embed_dim, outputdim))
        nn.init.normal_(self.cls_token, std=1e-6)
        self._packed: Dict = {}

    # ---- reference API surface -----------------------------------------------------------------------------
    @torch.jit.ignore
    def no_weight_decay(self):
        return {'time_pos_embed', 'cls_token', 'freq_pos_embed', 'token_pos_embed'}

    def load_state_dict(self, state_dict, strict=True, **kwargs):
        """uit.py:416-450: slice / bilinearly resize the positional embeddings when shapes differ."""
        if 'time_pos_embed' in state_dict and self.time_pos_embed.shape != state_dict['time_pos_embed'].shape:
            state_dict = dict(state_dict)
            self.change_pos_embedding(state_dict)
        return super().load_state_dict(state_dict, strict=strict, **kwargs)

    def change_pos_embedding(self, state_dict):
        tt, tf = self.time_pos_embed.shape[-1], self.freq_pos_embed.shape[-2]
        pt, pf = state_dict['time_pos_embed'], state_dict['freq_pos_embed']
        if tt <= pt.shape[-1]:
            state_dict['time_pos_embed'] = pt[..., :tt]
        else:
            state_dict['time_pos_embed'] = torch.nn.functional.interpolate(pt, size=(1, tt), align_corners=False, mode='bilinear')
        if tf <= pf.shape[-2]:
            state_dict['freq_pos_embed'] = pf[:, :, :tf, :]
        else:
            state_dict['freq_pos_embed'] = torch.nn.functional.interpolate(pf, size=(tf, 1), align_corners=False, mode='bilinear')

    def train(self, mode: bool = True):
        # The module can be flipped like any nn.Module, but only the eval path is implemented (Q11).
        return super().train(mode)

    # ---- kernel plumbing -----------------------------------------------------------------------------------------
    def _cfg(self) -> N.EncoderCfg:
        return N.EncoderCfg(self.depth, self.outputdim, self.patch_embed.grid_size[1], N.PRECISIONS[self.precision],
                            N.ATTENTION[self.attention_type], N.ACT["relu" if self.act_layer is nn.ReLU else "gelu"],
                            N.POOLING[self.pooling])

    def _encoder_blob(self, device: torch.device) -> torch.Tensor:
        """Pack the state_dict into the kernel layout (lazily; re-packed when any tensor changed)."""
        tensors = self.__dict__.get("_packed_tensors")
        if tensors is None:
            sd = dict(self.named_parameters())
            sd.update(dict(self.named_buffers()))
            tensors = [sd[n] for n in N.encoder_tensor_names(self.depth)]
            self.__dict__["_packed_tensors"] = tensors
        key = (str(device), self.precision) + tuple((t._version, t.data_ptr()) for t in tensors)
        if self._packed.get("key") != key:
            l = N.lib()
            cfg = self._cfg()
            host = [t.detach().to("cpu", torch.float32).contiguous() for t in tensors]
            ptrs = (C.c_void_p * len(host))(*[h.data_ptr() for h in host])
            nbytes = l.uitk_encoder_blob_bytes(C.byref(cfg))
            if nbytes == 0:
                N.check(-1, "uitk_encoder_blob_bytes")
            blob = torch.zeros(nbytes, dtype=torch.uint8)
            N.check(l.uitk_pack_encoder(C.byref(cfg), ptrs, blob.data_ptr(), nbytes), "uitk_pack_encoder")
            self._packed = {"key": key, "blob": blob.to(device)}
        return self._packed["blob"]

    def _check_ready(self, t: torch.Tensor, what: str):
        if self.training:
            raise NotImplementedError("uit_mobile_b200 implements inference only: call model.eval() "
                                      "(training branches of uit.py:453-459 are out of scope)")
        if not t.is_cuda:
            raise N.UitkError(f"{what}: UiT hot path runs on CUDA only (no CPU fallback); move the model and input to a B200")

    def encode(self, db: torch.Tensor, max_pow: torch.Tensor, out: Optional[torch.Tensor] = None, *,
               fixup: Optional[tuple] = None, workspace_out: Optional[list] = None, _blob: Optional[torch.Tensor] = None) -> torch.Tensor:
        """init_bn + crops + forward_features + forward_head on un-clamped log-mel (uit.py:460-492).

        ``fixup=(max_used, min_pow)`` turns the call into the device-conditional exact re-run of a speculative encode
        (``uitk_encoder_fixup``); ``workspace_out`` (a list) receives the launch workspace (tests: debug taps)."""
        self._check_ready(db, "encode")
        if db.dtype != torch.float32 or db.dim() != 3 or db.shape[1] != 64 or not db.is_contiguous():
            raise ValueError(f"db must be a contiguous float32 [B, 64, T] CUDA tensor, got {db.dtype} {tuple(db.shape)}")
        words = (max_pow,) + (tuple(fixup) if fixup is not None else ())
        for w in words:
            if w.dtype != torch.int32 or w.device != db.device or w.numel() != 1:
                raise ValueError("max_pow / fixup words must be one-element int32 tensors on the device of db")
        l = N.lib()
        B, _, T = db.shape
        cfg = self._cfg()
        blob = _blob if _blob is not None else self._encoder_blob(db.device)
        if out is None:
            probs = torch.empty((B, self.outputdim), dtype=torch.float32, device=db.device)
        else:
            if tuple(out.shape) != (B, self.outputdim) or out.dtype != torch.float32 or not out.is_contiguous() or out.device != db.device:
                raise ValueError("out must be a contiguous float32 [B, outputdim] tensor on the device of db")
            probs = out
        # launches are cut at multiples of the kernel's 128-row tile (5 clip-crops of 24 token slots): the tensor-core attention
        # sums over a tile's keys, so a clip's rounding depends on its position inside the tile; tile-aligned cuts keep
        # chunked / sharded runs bit-identical to a single launch
        tile = self.tile_clips(T)
        step = max(tile, self.max_clips_per_launch // int(l.uitk_num_crops(T, self.target_length)) // tile * tile)
        stream = torch.cuda.current_stream(db.device).cuda_stream
        with torch.cuda.device(db.device):
            ws = None
            for b0 in range(0, B, step):
                nb = min(step, B - b0)
                need = l.uitk_encoder_workspace_bytes(C.byref(cfg), nb, T, self.target_length)
                if need == 0:
                    N.check(-1, "uitk_encoder_workspace_bytes")
                if ws is None or ws.numel() < need:
                    ws = torch.empty(need, dtype=torch.uint8, device=db.device)     # stream-ordered caching allocator: not kept
                if fixup is None:
                    N.check(l.uitk_encoder(C.byref(cfg), blob.data_ptr(), db[b0:b0 + nb].data_ptr(), nb, T, self.target_length,
                                           1 if self.eval_avg == 'max' else 0, max_pow.data_ptr(), probs[b0:b0 + nb].data_ptr(),
                                           ws.data_ptr(), ws.numel(), stream), "uitk_encoder")
                else:
                    N.check(l.uitk_encoder_fixup(C.byref(cfg), blob.data_ptr(), db[b0:b0 + nb].data_ptr(), nb, T, self.target_length,
                                                 1 if self.eval_avg == 'max' else 0, max_pow.data_ptr(), fixup[0].data_ptr(),
                                                 fixup[1].data_ptr(), probs[b0:b0 + nb].data_ptr(), ws.data_ptr(), ws.numel(), stream),
                            "uitk_encoder_fixup")
            if workspace_out is not None:
                workspace_out.append(ws)
        return probs

    def _init_bn(self, x: torch.Tensor) -> torch.Tensor:
        """init_bn (uit.py:310-313, 460-462) on [B, 1, 64, T] (or [B, 64, T]): eval BatchNorm over the mel axis."""
        self._check_ready(x, "init_bn")
        shape = x.shape
        if x.dtype != torch.float32 or x.dim() not in (3, 4) or shape[-2] != 64 or (x.dim() == 4 and shape[1] != 1):
            raise ValueError(f"init_bn expects a float32 [B, 1, 64, T] spectrogram, got {x.dtype} {tuple(shape)}")
        xc = x.reshape(shape[0], 64, shape[-1]).contiguous()
        out = torch.empty_like(xc)
        cfg = self._cfg()
        with torch.cuda.device(x.device):
            N.check(N.lib().uitk_init_bn(C.byref(cfg), self._encoder_blob(x.device).data_ptr(), xc.data_ptr(), xc.shape[0], xc.shape[2],
                                         out.data_ptr(), torch.cuda.current_stream(x.device).cuda_stream), "uitk_init_bn")
        return out.reshape(shape)

    def forward_features(self, x: torch.Tensor) -> torch.Tensor:
        """uit.py:379-396: normalised spectrogram crop [B, 1, 64, T] (T <= 16 * grid + 15) -> tokens [B, N, 128] after the final
        LayerNorm (N = 4 * time patches, + the cls token for pooling='token').  One launch of the same kernels ``forward`` uses
        (the tensor-core megakernel for the UiT configuration), stopped before the pooling."""
        self._check_ready(x, "forward_features")
        if x.dtype != torch.float32 or x.dim() not in (3, 4) or x.shape[-2] != 64 or (x.dim() == 4 and x.shape[1] != 1):
            raise ValueError(f"forward_features expects a float32 [B, 1, 64, T] spectrogram, got {x.dtype} {tuple(x.shape)}")
        B, T = x.shape[0], x.shape[-1]
        xc = x.reshape(B, 64, T).contiguous()
        l = N.lib()
        cfg = self._cfg()
        target = 16 * self.patch_embed.grid_size[1] + 15
        if not (16 <= T <= target):
            raise ValueError(f"forward_features takes 16..{target} frames (time_pos_embed has {self.patch_embed.grid_size[1]} entries), got {T}")
        n_tok = int(l.uitk_tokens_total(C.byref(cfg), T, target))
        out = torch.empty((B, n_tok, self.embed_dim), dtype=torch.float32, device=x.device)
        if B == 0:
            return out
        with torch.cuda.device(x.device):
            need = l.uitk_forward_features_workspace_bytes(C.byref(cfg), B, T)
            if need == 0:
                N.check(-1, "uitk_forward_features_workspace_bytes")
            ws = torch.empty(need, dtype=torch.uint8, device=x.device)
            N.check(l.uitk_forward_features(C.byref(cfg), self._encoder_blob(x.device).data_ptr(), xc.data_ptr(), B, T, out.data_ptr(),
                                            ws.data_ptr(), ws.numel(), torch.cuda.current_stream(x.device).cuda_stream),
                    "uitk_forward_features")
        return out

    def forward_head(self, x: torch.Tensor) -> torch.Tensor:
        """uit.py:398-412: tokens [B, N, 128] -> sigmoid scores [B, outputdim] (pooling 'mean' | 'token' | 'dm')."""
        self._check_ready(x, "forward_head")
        if x.dtype != torch.float32 or x.dim() != 3 or x.shape[2] != self.embed_dim:
            raise ValueError(f"forward_head expects float32 tokens [B, N, {self.embed_dim}], got {x.dtype} {tuple(x.shape)}")
        xc = x.contiguous()
        out = torch.empty((x.shape[0], self.outputdim), dtype=torch.float32, device=x.device)
        cfg = self._cfg()
        with torch.cuda.device(x.device):
            N.check(N.lib().uitk_forward_head(C.byref(cfg), self._encoder_blob(x.device).data_ptr(), xc.data_ptr(), xc.shape[0], xc.shape[1],
                                              out.data_ptr(), torch.cuda.current_stream(x.device).cuda_stream), "uitk_forward_head")
        return out

    def tile_clips(self, T: int) -> int:
        """Clips per 128-row encoder tile for T frames: chunk / shard boundaries that are multiples of this (in clips)
        reproduce a single launch bit for bit.  The tile holds 5 clip-crops of 24 token slots for every clip length."""
        crops = int(N.lib().uitk_num_crops(T, self.target_length))
        return 5 // math.gcd(5, crops)

    def _finish(self, db: torch.Tensor, words: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Encoder + head given the un-clamped log-mel and the [max, min] power words of THIS rank's clips (Q2)."""
        max_w, min_w = words[0:1], words[1:2]
        if self.process_group is None:
            return self.encode(db, max_w, out=out)
        dist = torch.distributed
        if not self._cfg().tensor_core:
            # fp32 / variant kernels: the batch-global cutoff first (blocking all-reduce of one word), then the encoder
            dist.all_reduce(max_w, op=dist.ReduceOp.MAX, group=self.process_group)
            return self.encode(db, max_w, out=out)
        # Sharded, tensor-core configuration: NO collective on the critical path.  Encode speculatively with the rank-local
        # maximum, learn the global maximum meanwhile, then launch the device-conditional exact re-run: its kernels return at
        # once unless this rank's cutoff was below the global one AND one of its values lies under the global cutoff
        # (uitk_encoder_fixup).
        blob = self._encoder_blob(db.device) if db.shape[0] else None
        empty = torch.empty((0, self.outputdim), dtype=torch.float32, device=db.device)
        pw = getattr(self, "peer_words", None)
        if pw is not None:
            # global maximum through NVLink peer memory (sharding.PeerWords): publish now, read the peers' words after the
            # speculative encode (they are long there by then) - no NCCL kernel, no collective launch
            epoch = pw.publish(max_w)
            probs = self.encode(db, max_w, out=out, _blob=blob) if db.shape[0] else empty
            g = pw.collect(epoch, torch.empty(1, dtype=torch.int32, device=db.device))
            if db.shape[0]:
                self.encode(db, g, out=probs, fixup=(max_w, min_w), _blob=blob)
            return probs
        # NCCL: the all-reduce(MAX) of [max, -1 - min] (non-negative floats order like their int32 bit patterns) runs on the
        # NCCL stream under the speculative encode
        g = torch.bitwise_xor(words, self._words_mask(db.device))      # [max, ~min] = [max, -1 - min]: ONE MAX all-reduce for both
        work = dist.all_reduce(g, op=dist.ReduceOp.MAX, group=self.process_group, async_op=True)
        probs = self.encode(db, max_w, out=out, _blob=blob) if db.shape[0] else empty
        work.wait()
        if db.shape[0]:
            self.encode(db, g[0:1], out=probs, fixup=(max_w, min_w), _blob=blob)
        return probs

    def _words_mask(self, device) -> torch.Tensor:
        cache = self.__dict__.setdefault("_words_mask_cache", {})
        device = torch.device(device)
        if device not in cache:
            cache[device] = torch.tensor([0, -1], dtype=torch.int32, device=device)
        return cache[device]

    def _new_words(self, device) -> torch.Tensor:
        """[max power bits = 0, min power bits = +inf] on ``device``: a device-side clone of a cached constant (building the
        tensor from a Python list would be a blocking pageable host-to-device copy in every forward)."""
        return self.front_end.new_words(device)

    def forward(self, x: torch.Tensor, mixup=None) -> torch.Tensor:
        if self.training:
            raise NotImplementedError("uit_mobile_b200 implements inference only: call model.eval() "
                                      "(training branches of uit.py:453-459 are out of scope)")
        if self.eval_avg not in ('mean', 'max'):
            raise ValueError(f'Unknown Eval average function ({self.eval_avg})')
        if x.dim() == 2 and x.shape[0] == 0:
            # an empty shard still joins the collective of its group (the other ranks would block in it otherwise)
            if self.process_group is not None:
                if not x.is_cuda:
                    raise N.UitkError("UiT hot path runs on CUDA only (no CPU fallback)")
                return self._finish(torch.empty((0, 64, 1), dtype=torch.float32, device=x.device), self._new_words(x.device))
            return torch.empty((0, self.outputdim), dtype=torch.float32, device=x.device)
        if not x.is_cuda:
            raise N.UitkError("UiT hot path runs on CUDA only (no CPU fallback); move the model and input to a B200")
        words = self._new_words(x.device)
        db, _ = self.front_end.logmel_unclamped(x, max_pow=words[0:1], min_pow=words[1:2])
        return self._finish(db, words)


def _forward_sliding(self, stream: torch.Tensor, hop: int = 1600, window: int = 16000) -> torch.Tensor:
    """Scores of all sliding windows of one long stream: exactly ``self(stream.unfold(0, window, hop))`` (one top-dB scope over all
    windows, Q2), with the stream's STFT shared between overlapping windows (``FrontEnd.logmel_sliding``)."""
    if self.training:
        raise NotImplementedError("uit_mobile_b200 implements inference only: call model.eval()")
    if not stream.is_cuda:
        raise N.UitkError("UiT hot path runs on CUDA only (no CPU fallback)")
    words = self._new_words(stream.device)
    db, _ = self.front_end.logmel_sliding(stream, window, hop, max_pow=words[0:1], min_pow=words[1:2])
    return self._finish(db, words)


UITBase.forward_sliding = _forward_sliding


def _factory(depth: int, kwargs, **fixed) -> UITBase:
    model_kwargs = dict(patch_size=16, embed_dim=128, depth=depth, num_heads=2, mlp_ratio=3.0, pooling='mean',
                        init_bn=True, drop_path_rate=0.0, **(fixed or dict(act_layer=nn.ReLU, attention_type='BNeckAttention')))
    model_kwargs = {**model_kwargs, **kwargs}
    return UITBase(**model_kwargs)


# The reference's other h128 factories (uit.py:496-578): full Attention, GELU unless "_relu".  They run on the fp32 CUDA-core kernels.
def audio_transformer_h128_d4_m3_relu(**kwargs):     # uit.py:513-528
    return _factory(4, kwargs, act_layer=nn.ReLU)


def audio_transformer_h128_d4_m3(**kwargs):          # uit.py:531-545
    return _factory(4, kwargs, act_layer=None)


def audio_transformer_h128_d6_m3(**kwargs):          # uit.py:547-561
    return _factory(6, kwargs, act_layer=None)


def audio_transformer_h128_d6_m3_relu(**kwargs):     # uit.py:563-578
    return _factory(6, kwargs, act_layer=nn.ReLU)


def audio_transformer_h128_d3_m3_bneck_v2_relu(**kwargs):   # uit.py:496-511: names a class the reference never defines (Q10)
    return _factory(3, kwargs, act_layer=nn.ReLU, attention_type='BNeckAttentionV2')      # -> KeyError, as upstream


def uit_xs(**kwargs):       # uit.py:581-597
    return _factory(12, kwargs)


def uit_xxs(**kwargs):      # uit.py:619-635
    return _factory(6, kwargs)


def uit_xxxs(**kwargs):     # uit.py:600-616
    return _factory(4, kwargs)


PRETRAINED_CHECKPOINTS = {   # uit.py:639-655
    'uit_xs': {'model': uit_xs, 'model_kwargs': dict(outputdim=537, target_length=102),
               'chkpt': 'https://zenodo.org/record/7690036/files/uit_xs_mAP3409.pt?download=1'},
    'uit_xxs': {'model': uit_xxs, 'model_kwargs': dict(outputdim=537, target_length=102),
                'chkpt': 'https://zenodo.org/record/7690036/files/uit_xxs_mAP3221.pt?download=1'},
    'uit_xxxs': {'model': uit_xxxs, 'model_kwargs': dict(outputdim=537, target_length=102),
                 'chkpt': 'https://zenodo.org/record/7690036/files/uit_xxxs_mAP3097.pt?download=1'},
}
