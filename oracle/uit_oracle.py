"""CPU oracle for the UiT batched-inference hot path.  TEST INFRASTRUCTURE ONLY.

This file is the *checker*: a functional, state_dict-driven restatement (torch CPU ops, fp32) of the
reference's eval-mode path ``raw waveform -> log-mel -> BatchNorm -> patch embed -> pre-norm ViT blocks ->
head -> 537 sigmoid scores``.  Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it.  Nothing under ``uit_mobile_b200/`` imports it and the
product path has no CPU fallback.

Parity pin: the reference ships no tests or golden vectors for this path and its known
answers (README.md:85-139) need network checkpoints, so the oracle is pinned against the *reference itself*
imported on CPU: ``tests/golden/generate_golden.py`` runs ``/root/reference/models/uit.py`` under a timm shim,
asserts this oracle reproduces it, and commits the vectors that ``tests/test_oracle_golden.py`` re-checks
without the reference.

Citations: ``uit.py`` = /root/reference/models/uit.py; ``TA:`` = torchaudio (pinned 0.13.0 by the reference's
requirements.txt:17, 2.11.0 installed) ``functional/functional.py`` and ``transforms/_transforms.py``.
The algorithm of the front-end lives in that third-party dependency; it is restated here with torch ops only
(no torchaudio import).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

# Fixed front-end configuration of the hot path (uit.py:287-308).
SAMPLE_RATE = 16000
N_FFT = 512
WIN = 512
HOP = 160
N_MELS = 64
N_FREQS = N_FFT // 2 + 1
TOP_DB = 120.0
AMIN = 1e-10
BN_EPS = 1e-5          # torch.nn.BatchNorm2d default (uit.py:311-313)
LN_EPS_BLOCK = 1e-6    # partial(nn.LayerNorm, eps=1e-6) (uit.py:337)
LN_EPS_HEAD = 1e-5     # plain nn.LayerNorm in outputlayer (uit.py:358-360)
PATCH = 16
EMBED = 128
HEADS = 2
DEPTH = {"uit_xs": 12, "uit_xxs": 6, "uit_xxxs": 4}   # uit.py:581-635


# ----------------------------------------------------------------------------------------------------------
# Front-end constants (buffers that live in the state_dict: Q9)
# ----------------------------------------------------------------------------------------------------------
def hann_window() -> Tensor:
    """``torch.hann_window(512)`` periodic, as built by ``Spectrogram.__init__`` (TA:_transforms.py:96-99)."""
    return torch.hann_window(WIN, periodic=True, dtype=torch.float32)


def _hz_to_mel_htk(f: float) -> float:
    return 2595.0 * math.log10(1.0 + f / 700.0)          # TA:functional.py:425-455


def melscale_fbanks_htk(f_min: float = 0.0, f_max: float = 8000.0, n_mels: int = N_MELS) -> Tensor:
    """HTK triangular filterbank ``fb[257, n_mels]`` (TA:functional.py:492-587, norm=None, mel_scale='htk')."""
    all_freqs = torch.linspace(0, SAMPLE_RATE // 2, N_FREQS)
    m_pts = torch.linspace(_hz_to_mel_htk(f_min), _hz_to_mel_htk(f_max), n_mels + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)     # TA:functional.py:458-489
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.max(torch.zeros(1), torch.min(down, up))


# ----------------------------------------------------------------------------------------------------------
# Front-end: MelSpectrogram + AmplitudeToDB (uit.py:298-308, 455)
# ----------------------------------------------------------------------------------------------------------
def num_frames(L: int) -> int:
    return 1 + L // HOP


def power_spectrogram(wav: Tensor, window: Tensor) -> Tensor:
    """[B, L] -> [B, 257, T]: reflect-pad 256, frame, window, 512-pt rDFT, ``abs().pow(2)``
    (TA:functional.py:123-144)."""
    spec = torch.stft(wav, n_fft=N_FFT, hop_length=HOP, win_length=WIN, window=window, center=True,
                      pad_mode="reflect", normalized=False, onesided=True, return_complex=True)
    return spec.abs().pow(2.0)


def mel_power(spec: Tensor, fb: Tensor) -> Tensor:
    """[B, 257, T] -> [B, 64, T] (TA:_transforms.py:407-419)."""
    return torch.matmul(spec.transpose(-1, -2), fb).transpose(-1, -2)


def power_to_db(mel: Tensor, cutoff_max_db: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    """10*log10(clamp(x, 1e-10)); then ONE top-dB cutoff for the whole 3-D batch (Q2,
    TA:functional.py:390-399).  ``cutoff_max_db`` lets a sharded caller supply the batch-global maximum.
    Returns (dB, max dB actually used)."""
    x_db = 10.0 * torch.log10(torch.clamp(mel, min=AMIN))
    gmax = x_db.amax() if cutoff_max_db is None else torch.as_tensor(cutoff_max_db, dtype=x_db.dtype)
    return torch.max(x_db, gmax - TOP_DB), gmax


def logmel(wav: Tensor, window: Tensor, fb: Tensor, cutoff_max_db: Optional[Tensor] = None) -> Tensor:
    return power_to_db(mel_power(power_spectrogram(wav, window), fb), cutoff_max_db)[0]


# ----------------------------------------------------------------------------------------------------------
# Encoder
# ----------------------------------------------------------------------------------------------------------
def init_bn(db: Tensor, sd: Dict[str, Tensor]) -> Tensor:
    """Eval-mode BatchNorm2d over the mel axis (uit.py:310-313, 460-462).  [B,64,T] -> [B,64,T]."""
    mean, var = sd["init_bn.1.running_mean"], sd["init_bn.1.running_var"]
    w, b = sd["init_bn.1.weight"], sd["init_bn.1.bias"]
    # F.batch_norm on [B, C=64, 1, T] is what the reference executes.
    return F.batch_norm(db.unsqueeze(2), mean, var, w, b, False, 0.0, BN_EPS).squeeze(2)


def crop_starts(T: int, target_length: int) -> List[int]:
    """Start frames of the eval crops (uit.py:468-481): split(target) with a short tail replaced by the
    last ``target`` frames.  T <= target -> single crop at 0 covering all T frames."""
    if T <= target_length:
        return [0]
    n = -(-T // target_length)
    return [min(c * target_length, T - target_length) for c in range(n)]


def patch_tokens(xn: Tensor, sd: Dict[str, Tensor]) -> Tensor:
    """AudioPatchEmbed conv + pos-embeds + flatten (uit.py:380-388).  [B,64,Tc] -> [B, 4*t, 128]."""
    y = F.conv2d(xn.unsqueeze(1), sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], stride=PATCH)
    t = y.shape[-1]
    y = y + sd["time_pos_embed"][:, :, :, :t]
    y = y + sd["freq_pos_embed"]
    return y.flatten(2).transpose(1, 2)          # 'b c f t -> b (f t) c'


def attention(x: Tensor, sd: Dict[str, Tensor], i: int) -> Tensor:
    """BNeckAttention (uit.py:89-122): inner dim 32, 2 heads x 16, scale = (128//2)**-0.5 = 0.125 (Q3); or the full
    ``Attention`` (uit.py:124-178, causal=False): inner dim 128, 2 heads x 64, same scale.  Which one is read off the qkv
    weight's shape ([96, 128] vs [384, 128])."""
    B, N, C = x.shape
    inner = sd[f"blocks.{i}.attn.qkv.weight"].shape[0] // 3
    qkv = F.linear(x, sd[f"blocks.{i}.attn.qkv.weight"], sd[f"blocks.{i}.attn.qkv.bias"])
    qkv = qkv.reshape(B, N, 3, HEADS, inner // HEADS).permute(2, 0, 3, 1, 4)
    q, k, v = qkv.unbind(0)
    a = (q @ k.transpose(-2, -1)) * ((C // HEADS) ** -0.5)
    a = a.softmax(dim=-1)
    o = (a @ v).transpose(1, 2).reshape(B, N, inner)
    return F.linear(o, sd[f"blocks.{i}.attn.proj.weight"], sd[f"blocks.{i}.attn.proj.bias"])


def mlp(x: Tensor, sd: Dict[str, Tensor], i: int, act: str = "relu") -> Tensor:
    """fc2(act(fc1(x))) (uit.py:197-203; act_layer=nn.ReLU for all UiT archs, Q7; nn.GELU (exact erf) is the UITBase
    default, uit.py:338)."""
    h = F.linear(x, sd[f"blocks.{i}.mlp.fc1.weight"], sd[f"blocks.{i}.mlp.fc1.bias"])
    h = F.relu(h) if act == "relu" else F.gelu(h)
    return F.linear(h, sd[f"blocks.{i}.mlp.fc2.weight"], sd[f"blocks.{i}.mlp.fc2.bias"])


def block(x: Tensor, sd: Dict[str, Tensor], i: int, act: str = "relu") -> Tensor:
    """Pre-norm residual block (uit.py:245-248); LayerScale/DropPath are Identity."""
    p = f"blocks.{i}."
    x = x + attention(F.layer_norm(x, (EMBED,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], LN_EPS_BLOCK), sd, i)
    x = x + mlp(F.layer_norm(x, (EMBED,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], LN_EPS_BLOCK), sd, i, act)
    return x


def depth_of(sd: Dict[str, Tensor]) -> int:
    return 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("blocks."))


def features(xn: Tensor, sd: Dict[str, Tensor], trace: Optional[list] = None, act: str = "relu", pooling: str = "mean") -> Tensor:
    """forward_features (uit.py:379-396).  pooling='token' prepends cls_token + token_pos_embed (uit.py:389-392); the
    UiT archs use pooling='mean' (no cls token, Q4)."""
    x = patch_tokens(xn, sd)
    if pooling == "token":
        cls = sd["cls_token"].expand(x.shape[0], -1, -1) + sd["token_pos_embed"]
        x = torch.cat((cls, x), dim=1)
    if trace is not None:
        trace.append(x.clone())
    for i in range(depth_of(sd)):
        x = block(x, sd, i, act)
        if trace is not None:
            trace.append(x.clone())
    return F.layer_norm(x, (EMBED,), sd["norm.weight"], sd["norm.bias"], LN_EPS_BLOCK)


def _outputlayer(x: Tensor, sd: Dict[str, Tensor]) -> Tensor:
    x = F.layer_norm(x, (EMBED,), sd["outputlayer.0.weight"], sd["outputlayer.0.bias"], LN_EPS_HEAD)
    return F.linear(x, sd["outputlayer.1.weight"], sd["outputlayer.1.bias"])


def head(x: Tensor, sd: Dict[str, Tensor], pooling: str = "mean") -> Tensor:
    """forward_head (uit.py:398-412).  'mean': mean over tokens, LN(1e-5), Linear, sigmoid (Q1); 'token': the cls row;
    'dm': unpack (f t), mean over frequency, head + sigmoid per time step, mean of the scores."""
    if pooling == "token":
        return _outputlayer(x[:, 0], sd).sigmoid()
    if pooling == "mean":
        return _outputlayer(x.mean(1), sd).sigmoid()
    if pooling == "dm":
        B, N, D = x.shape
        return _outputlayer(x.reshape(B, 4, N // 4, D).mean(1), sd).sigmoid().mean(1)
    raise ValueError(pooling)


def encode(db: Tensor, sd: Dict[str, Tensor], target_length: int = 102, eval_avg: str = "mean", act: str = "relu",
           pooling: str = "mean") -> Tensor:
    """BatchNorm + crop loop + features + head on an already computed log-mel (uit.py:460-492)."""
    xn = init_bn(db, sd)
    T = xn.shape[-1]
    starts = crop_starts(T, target_length)
    if T <= target_length:
        return head(features(xn, sd, None, act, pooling), sd, pooling)
    outs = [head(features(xn[..., s:s + target_length], sd, None, act, pooling), sd, pooling) for s in starts]
    y = torch.stack(outs, -1)
    if eval_avg == "mean":
        return y.mean(-1)
    if eval_avg == "max":
        return y.max(-1)[0]
    raise ValueError(f"Unknown Eval average function ({eval_avg})")


@torch.no_grad()
def forward(sd: Dict[str, Tensor], wav: Tensor, target_length: int = 102, eval_avg: str = "mean",
            cutoff_max_db: Optional[Tensor] = None, act: str = "relu", pooling: str = "mean") -> Tensor:
    """UITBase.forward, eval branch (uit.py:452-493).  wav [B, L] fp32 -> [B, outputdim] probabilities."""
    if wav.dim() != 2:
        raise ValueError("expected a [B, L] waveform batch")
    db = logmel(wav, sd["front_end.0.spectrogram.window"], sd["front_end.0.mel_scale.fb"], cutoff_max_db)
    return encode(db, sd, target_length, eval_avg, act, pooling)


@torch.no_grad()
def forward_trace(sd: Dict[str, Tensor], wav: Tensor, target_length: int = 102) -> Dict[str, Tensor]:
    """Single-crop trace of every stage (debugging aid for the kernels)."""
    db = logmel(wav, sd["front_end.0.spectrogram.window"], sd["front_end.0.mel_scale.fb"])
    xn = init_bn(db, sd)
    tr: list = []
    f = features(xn[..., :target_length], sd, tr)
    return {"db": db, "bn": xn, "tokens": tr[0], "blocks": torch.stack(tr[1:]), "features": f, "probs": head(f, sd)}
