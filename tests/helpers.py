"""Shared test helpers: seeded weights / inputs that need neither the reference nor a GPU.

Everything here is generated with numpy's PCG64 (stable across numpy versions), so the golden generator (run in
the build container against /root/reference) and the GPU box (no reference) see bit-identical weights/inputs.
"""
from __future__ import annotations

import os
from typing import Dict

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN_DIR = os.path.join(REPO, "tests", "golden")
DEPTH = {"uit_xs": 12, "uit_xxs": 6, "uit_xxxs": 4}
ARCHS = ("uit_xs", "uit_xxs", "uit_xxxs")
OUTPUTDIM = 537


def _t(a) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))


# UITBase variants beyond the three UiT archs: name -> (depth, attention_type, act, pooling)
VARIANTS = {
    "h128_d4_m3": (4, "Attention", "gelu", "mean"),                 # audio_transformer_h128_d4_m3 (uit.py:531-545)
    "h128_d4_m3_relu_token": (4, "Attention", "relu", "token"),     # audio_transformer_h128_d4_m3_relu(pooling='token')
    "uit_xxxs_dm": (4, "BNeckAttention", "relu", "dm"),             # uit_xxxs(pooling='dm')
    "uit_xxxs_gelu_token": (4, "BNeckAttention", "gelu", "token"),  # uit_xxxs(act_layer=nn.GELU, pooling='token')
}


def make_state_dict(arch: str, kind: str = "trained", seed: int = 1, outputdim: int = OUTPUTDIM,
                    grid_t: int = 6, attention: str = "BNeckAttention") -> Dict[str, torch.Tensor]:
    """A full UiT state_dict (same keys/shapes/dtypes as the reference).

    kind='init'    ~ the reference's default init statistics (flat outputs, BN = identity).
    kind='trained' ~ weights with trained-like scale so that no branch is trivially identity, and a
                     sparse-activation head (a handful of classes above 0.1 per clip, the rest near sigmoid(-6)):
                     literal top-5 comparisons are meaningful on it.
    """
    from oracle import uit_oracle as O     # only for the analytic window / mel filterbank buffers
    depth = DEPTH[arch] if arch in DEPTH else VARIANTS[arch][0]
    if arch in VARIANTS:
        attention = VARIANTS[arch][1]
    full = attention == "Attention"
    g = np.random.Generator(np.random.PCG64([seed, depth, 0 if kind == "init" else 1] + ([7] if full else [])))
    tr = kind == "trained"

    def lin(out_f, in_f, std):
        return _t(g.standard_normal((out_f, in_f)) * std)

    def vec(n, mean, std):
        return _t(mean + g.standard_normal(n) * std)

    D, H, I = 128, 384, 32
    sd: Dict[str, torch.Tensor] = {}
    if full:
        I = D                      # Attention: qkv 128 -> 384, proj 128 -> 128
    sd["cls_token"] = _t(g.standard_normal((1, 1, D)) * (0.3 if tr else 1e-6))
    sd["token_pos_embed"] = _t(g.standard_normal((1, D)) * 0.02)
    sd["time_pos_embed"] = _t(g.standard_normal((1, D, 1, grid_t)) * (0.3 if tr else 0.02))
    sd["freq_pos_embed"] = _t(g.standard_normal((1, D, 4, 1)) * (0.3 if tr else 0.02))
    sd["front_end.0.spectrogram.window"] = O.hann_window()
    sd["front_end.0.mel_scale.fb"] = O.melscale_fbanks_htk()
    sd["init_bn.1.weight"] = vec(64, 1.0, 0.1 if tr else 0.0)
    sd["init_bn.1.bias"] = vec(64, 0.0, 0.1 if tr else 0.0)
    sd["init_bn.1.running_mean"] = vec(64, 5.0 if tr else 0.0, 5.0 if tr else 0.0)
    sd["init_bn.1.running_var"] = _t(g.uniform(50.0, 200.0, 64)) if tr else torch.ones(64)
    sd["init_bn.1.num_batches_tracked"] = torch.tensor(1234 if tr else 0, dtype=torch.int64)
    sd["patch_embed.proj.weight"] = _t(g.standard_normal((D, 1, 16, 16)) * (1.0 / 16.0))
    sd["patch_embed.proj.bias"] = vec(D, 0.0, 0.05)
    for i in range(depth):
        p = f"blocks.{i}."
        sd[p + "norm1.weight"] = vec(D, 1.0, 0.1 if tr else 0.0)
        sd[p + "norm1.bias"] = vec(D, 0.0, 0.05 if tr else 0.0)
        sd[p + "attn.qkv.weight"] = lin(3 * I, D, 0.09 if tr else 0.02)
        sd[p + "attn.qkv.bias"] = vec(3 * I, 0.0, 0.05 if tr else 0.0)
        sd[p + "attn.proj.weight"] = lin(D, I, (0.06 if full else 0.12) if tr else 0.02)
        sd[p + "attn.proj.bias"] = vec(D, 0.0, 0.05 if tr else 0.0)
        sd[p + "norm2.weight"] = vec(D, 1.0, 0.1 if tr else 0.0)
        sd[p + "norm2.bias"] = vec(D, 0.0, 0.05 if tr else 0.0)
        sd[p + "mlp.fc1.weight"] = lin(H, D, 0.09 if tr else 0.02)
        sd[p + "mlp.fc1.bias"] = vec(H, 0.0, 0.05 if tr else 0.0)
        sd[p + "mlp.fc2.weight"] = lin(D, H, 0.04 if tr else 0.02)
        sd[p + "mlp.fc2.bias"] = vec(D, 0.0, 0.05 if tr else 0.0)
    sd["norm.weight"] = vec(D, 1.0, 0.1 if tr else 0.0)
    sd["norm.bias"] = vec(D, 0.0, 0.05 if tr else 0.0)
    sd["outputlayer.0.weight"] = vec(D, 1.0, 0.1 if tr else 0.0)
    sd["outputlayer.0.bias"] = vec(D, 0.0, 0.05 if tr else 0.0)
    if tr:
        # Sparse-activation head, like a trained tagger (the reference's own known answers, README.md:85-116, have a
        # handful of classes above 0.1: 0.4467 / 0.3263 / 0.1718 ...): ~40 "active" classes with large-norm rows around a
        # bias of -4, every other class pinned near sigmoid(-6).  Top-5 gaps are then >> the bf16 noise, so LITERAL top-5
        # equality is a meaningful check (round-1 verdict: the old std-0.25 / bias -1 head saturated ~50 classes).
        n_act = max(5, (40 * outputdim) // 537)
        scale = np.full((outputdim, 1), 0.04)
        bias = np.full(outputdim, -6.0)
        act = g.permutation(outputdim)[:n_act]
        scale[act] = 0.22
        bias[act] = -4.0
        sd["outputlayer.1.weight"] = _t(g.standard_normal((outputdim, D)) * scale)
        sd["outputlayer.1.bias"] = _t(bias + g.standard_normal(outputdim) * 0.3)
    else:
        sd["outputlayer.1.weight"] = lin(outputdim, D, 0.02)
        sd["outputlayer.1.bias"] = vec(outputdim, 0.0, 0.0)
    return sd


def make_mnv2_state_dict(kind: str = "trained", seed: int = 3, outputdim: int = OUTPUTDIM) -> Dict[str, torch.Tensor]:
    """A full MobileNetV2 state_dict (keys / shapes of the reference's module tree, mobilenetv2.py:118-161), seeded.
    He-scaled convolutions keep the activations O(1) through the 53 layers; 'trained' adds non-trivial BatchNorm statistics and
    the sparse-activation classifier of ``make_state_dict``."""
    from oracle import uit_oracle as O
    from oracle.mobilenetv2_oracle import SETTING
    g = np.random.Generator(np.random.PCG64([seed, 0 if kind == "init" else 1, 22]))
    tr = kind == "trained"
    sd: Dict[str, torch.Tensor] = {"front_end.0.spectrogram.window": O.hann_window(), "front_end.0.mel_scale.fb": O.melscale_fbanks_htk()}

    def unit(conv, bn, cout, cin_g, k, gain=2.0):
        fan_in = cin_g * k * k
        sd[conv + ".weight"] = _t(g.standard_normal((cout, cin_g, k, k)) * np.sqrt(gain / fan_in))
        sd[bn + ".weight"] = _t(1.0 + g.standard_normal(cout) * (0.1 if tr else 0.0))
        sd[bn + ".bias"] = _t(g.standard_normal(cout) * (0.1 if tr else 0.0))
        sd[bn + ".running_mean"] = _t(g.standard_normal(cout) * (0.2 if tr else 0.0))
        sd[bn + ".running_var"] = _t(g.uniform(0.5, 1.5, cout)) if tr else torch.ones(cout)
        sd[bn + ".num_batches_tracked"] = torch.tensor(77 if tr else 0, dtype=torch.int64)

    unit("features.0.0", "features.0.1", 32, 1, 3, gain=0.002)      # the input is log-mel dB (tens of dB): scale it to O(1)
    f, inp = 1, 32
    for t, c, n, s in SETTING:
        for i in range(n):
            p, hidden = f"features.{f}.conv", inp * t
            j = 0
            if t != 1:
                unit(f"{p}.0.0", f"{p}.0.1", hidden, inp, 1); j = 1
            unit(f"{p}.{j}.0", f"{p}.{j}.1", hidden, 1, 3)
            unit(f"{p}.{j + 1}", f"{p}.{j + 2}", c, hidden, 1, gain=1.0)
            inp, f = c, f + 1
    unit(f"features.{f}.0", f"features.{f}.1", 1280, inp, 1)
    if tr:
        n_act = max(5, (40 * outputdim) // 537)
        scale = np.full((outputdim, 1), 0.01)
        bias = np.full(outputdim, -6.0)
        act = g.permutation(outputdim)[:n_act]
        scale[act] = 0.05
        bias[act] = -4.0
        sd["classifier.1.weight"] = _t(g.standard_normal((outputdim, 1280)) * scale)
        sd["classifier.1.bias"] = _t(bias + g.standard_normal(outputdim) * 0.3)
    else:
        sd["classifier.1.weight"] = _t(g.standard_normal((outputdim, 1280)) * 0.01)
        sd["classifier.1.bias"] = torch.zeros(outputdim)
    return sd


def build_variant(models_pkg, name: str, **kw):
    """Construct VARIANTS[name] from a ``models`` package (the reference's or uit_mobile_b200's: same factories / kwargs)."""
    import torch.nn as nn
    depth, attention, act, pooling = VARIANTS[name]
    act_layer = nn.ReLU if act == "relu" else nn.GELU
    if attention == "Attention":
        fac = getattr(models_pkg, f"audio_transformer_h128_d{depth}_m3" + ("_relu" if act == "relu" else ""))
        return fac(outputdim=OUTPUTDIM, target_length=102, pooling=pooling, **kw)
    return models_pkg.uit_xxxs(outputdim=OUTPUTDIM, target_length=102, pooling=pooling, act_layer=act_layer, **kw)


def variant_inputs():
    return {"noise": noise_clips(8), "adversarial": adversarial_batch(), "short14336": noise_clips(3, 14336, seed=12),
            "len16160": noise_clips(2, 16160, seed=14), "long10s": noise_clips(2, 160000, seed=13)}


def noise_clips(B: int, L: int = 16000, seed: int = 1234, amp: float = 0.1) -> np.ndarray:
    """The synthetic workload of SURVEY §8d: amp*randn clamped to [-1, 1], float32."""
    g = np.random.Generator(np.random.PCG64([seed, L]))
    return np.clip(g.standard_normal((B, L), dtype=np.float32) * np.float32(amp), -1.0, 1.0)


def adversarial_batch(L: int = 16000) -> np.ndarray:
    """zeros / full-scale sine / single impulse / loud noise / very quiet noise in ONE batch: exercises the
    batch-global top-dB cutoff (Q2) and the amin clamp."""
    t = np.arange(L, dtype=np.float64)
    x = np.zeros((5, L), dtype=np.float32)
    x[1] = np.sin(2 * np.pi * 440.0 * t / 16000.0).astype(np.float32)
    x[2, L // 2] = 1.0
    x[3] = noise_clips(1, L, seed=7, amp=0.9)[0]
    x[4] = noise_clips(1, L, seed=8, amp=1e-4)[0]
    return x


def load_golden(name: str):
    return np.load(os.path.join(GOLDEN_DIR, name))


def samples_int16() -> np.ndarray:
    """The reference's 11 sample clips (int16), zero-padded to 16384, plus their true lengths."""
    z = load_golden("samples_int16.npz")
    return z["pcm"], z["length"], [str(s) for s in z["names"]]


def logits_of(p: np.ndarray) -> np.ndarray:
    """log(p / (1 - p)) in float64 (the model outputs sigmoid probabilities, Q1)."""
    p = np.clip(p.astype(np.float64), 1e-12, 1.0 - 1e-12)
    return np.log(p) - np.log1p(-p)


def topk_report(ref: np.ndarray, got: np.ndarray, k: int = 5, eps_logit: float = 0.05) -> Dict[str, float]:
    """LITERAL top-k check in logit space.  A clip is *decisive* when the reference's k-th and (k+1)-th logits are more
    than 2*eps_logit apart: the top-k SET of an implementation whose logits are within eps_logit must then be identical.
    Returns the max |d logit| (over classes whose reference probability is representable: 1e-6 < p < 1 - 1e-6), the
    fraction of decisive clips, the fraction of decisive clips whose literal top-k set matches (must be 1.0), the fraction
    of ALL clips whose literal top-k set matches, and the mean size of the strict `must` set (reference classes more
    than 2*eps_logit above the k-th): a vacuous check shows up as decisive == 0 / must == 0."""
    zr, zg = logits_of(ref), logits_of(got)
    ok = (ref > 1e-6) & (ref < 1 - 1e-6)
    dz = float(np.abs(zr - zg)[ok].max()) if ok.any() else 0.0
    decisive = match_dec = match_all = 0
    must_sizes = []
    for r, o in zip(zr, zg):
        srt = np.sort(r)
        kth, nxt = srt[-k], srt[-k - 1]
        top_r = set(np.argsort(-r, kind="stable")[:k])
        top_o = set(np.argsort(-o, kind="stable")[:k])
        same = top_r == top_o
        match_all += same
        must_sizes.append(int((r > kth + 2 * eps_logit).sum()))
        if kth - nxt > 2 * eps_logit:
            decisive += 1
            match_dec += same
    n = len(zr)
    return {"max_dlogit": dz, "decisive_frac": decisive / n, "decisive_match_frac": (match_dec / decisive) if decisive else 0.0,
            "literal_match_frac": match_all / n, "must_mean": float(np.mean(must_sizes)), "n": n}


def tie_aware_topk_equal(ref: np.ndarray, got: np.ndarray, k: int = 5, eps: float = 1e-3) -> bool:
    """top-k label sets agree up to substitutions among classes whose REFERENCE probability lies within
    2*eps of the reference k-th value (SURVEY §7.2 tolerance statement)."""
    for r, o in zip(ref, got):
        kth = np.sort(r)[-k]
        must = set(np.nonzero(r > kth + 2 * eps)[0])
        may = set(np.nonzero(r >= kth - 2 * eps)[0])
        top = set(np.argsort(-o, kind="stable")[:k])
        if not (must <= top and top <= may):
            return False
    return True
