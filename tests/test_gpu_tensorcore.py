"""GPU tests of the tcgen05 tensor-core encoder (precision='bf16'): UMMA plumbing self-test, per-block residual
stream against the reference trace, and scores against the reference goldens.

Stated tolerance (north_star: "logits within a stated bf16 tolerance with identical top-5"): bf16 GEMM operands,
fp32 accumulate / residual / LayerNorm / softmax.  ONE bound per weight set, in logits and in probabilities:
  'init'    (reference-like init statistics, flat 0.3-0.7 outputs):  |d logit| <= 5e-3, |d prob| <= 1e-3; top-5 compared
            tie-aware (literal top-5 equality is unattainable on flat outputs even in TF32: SURVEY 7.2);
  'trained' (large-magnitude blocks + sparse-activation head, helpers.make_state_dict): |d logit| <= 0.08 (measured 0.057 on
            UiT-XS, 0.030 XXS, 0.029 XXXS), |d prob| <= 1.5e-2 (measured 8.5e-3); LITERAL top-5 set equality on every clip whose reference 5th / 6th logits are more
            than 2*eps apart, with the assertion that the strict `must` set is non-empty on >= 90 % of the clips (the check
            cannot go vacuous)."""
import numpy as np
import pytest
import torch

from oracle import uit_oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
# measured on B200 (round 2): init 1.9e-3 logit / 4.7e-4 prob; trained 5.7e-2 logit (XS, 12 blocks; the worst clip is a 2400-sample one) / 8.6e-3 prob
TOL = {"init": 1e-3, "trained": 1.5e-2}
EPS_LOGIT = {"init": 5e-3, "trained": 0.08}


def pack_kmajor(w: torch.Tensor) -> torch.Tensor:
    """[N, K] fp32 -> bf16 K-major core-matrix layout [(K/8), N, 8] (csrc/tc_ptx.cuh)."""
    n, k = w.shape
    return w.to(torch.bfloat16).view(n, k // 8, 8).permute(1, 0, 2).contiguous()


@pytest.mark.parametrize("N,K,init", [(128, 128, False), (96, 128, False), (128, 32, True), (128, 256, False), (128, 128, True), (64, 64, False)])
def test_umma_selftest(N, K, init):
    from uit_mobile_b200 import _native as N_
    g = torch.Generator().manual_seed(N * 1000 + K)
    a = torch.randn(128, K, generator=g)
    b = torch.randn(N, K, generator=g)
    c0 = torch.randn(128, N, generator=g) if init else None
    ref = a.to(torch.bfloat16).double() @ b.to(torch.bfloat16).double().T + (c0.double() if init else 0)
    a_d, bp_d = a.to(DEV), pack_kmajor(b).to(DEV)
    c0_d = c0.to(DEV) if init else None
    out = torch.full((128, N), float("nan"), device=DEV)
    N_.check(N_.lib().uitk_selftest_umma(a_d.data_ptr(), bp_d.data_ptr(), c0_d.data_ptr() if init else None, out.data_ptr(), N, K,
                                         torch.cuda.current_stream().cuda_stream), "uitk_selftest_umma")
    torch.cuda.synchronize()
    err = (out.cpu().double() - ref).abs().max().item()
    assert err <= 1e-3 * max(1.0, ref.abs().max().item()), err


@pytest.mark.parametrize("N,K,init", [(16, 128, False), (128, 128, False), (64, 64, True), (16, 16, False)])
def test_umma_selftest_a_in_tmem(N, K, init):
    """tcgen05.mma with the A operand in tensor memory (the layout the encoder's P V product relies on)."""
    from uit_mobile_b200 import _native as N_
    g = torch.Generator().manual_seed(N * 1000 + K + 7)
    a = torch.randn(128, K, generator=g)
    b = torch.randn(N, K, generator=g)
    c0 = torch.randn(128, N, generator=g) if init else None
    ref = a.to(torch.bfloat16).double() @ b.to(torch.bfloat16).double().T + (c0.double() if init else 0)
    a_d, bp_d = a.to(DEV), pack_kmajor(b).to(DEV)
    c0_d = c0.to(DEV) if init else None
    out = torch.full((128, N), float("nan"), device=DEV)
    N_.check(N_.lib().uitk_selftest_umma_ts(a_d.data_ptr(), bp_d.data_ptr(), c0_d.data_ptr() if init else None, out.data_ptr(), N, K,
                                            torch.cuda.current_stream().cuda_stream), "uitk_selftest_umma_ts")
    torch.cuda.synchronize()
    err = (out.cpu().double() - ref).abs().max().item()
    assert err <= 1e-3 * max(1.0, ref.abs().max().item()), err


def _model(arch, kind, precision, depth=None, **kw):
    import uit_mobile_b200 as U
    sd = H.make_state_dict(arch, kind)
    if depth is None:
        m = getattr(U.models, arch)(outputdim=537, target_length=102, precision=precision, **kw)
    else:
        sd = {k: v for k, v in sd.items() if not k.startswith("blocks.") or int(k.split(".")[1]) < depth}
        m = U.models.UITBase(outputdim=537, target_length=102, patch_size=16, embed_dim=128, depth=depth, num_heads=2,
                             mlp_ratio=3.0, pooling="mean", init_bn=True, act_layer=torch.nn.ReLU,
                             attention_type="BNeckAttention", precision=precision)
    m.load_state_dict(sd, strict=True)
    return m.to(DEV).eval()


@pytest.mark.parametrize("depth", [1, 2, 4])
def test_bf16_residual_stream_vs_reference_trace(depth):
    from uit_mobile_b200 import _native as N_
    z = H.load_golden("trace_xxxs.npz")
    m = _model("uit_xxxs", "trained", "bf16", depth)
    x = torch.from_numpy(H.noise_clips(32)[:2]).to(DEV)
    N_.lib().uitk_debug_taps(1)
    try:
        db, mp = m.front_end.logmel_unclamped(x)
        ws = []
        m.encode(db, mp, workspace_out=ws)
        torch.cuda.synchronize()
    finally:
        N_.lib().uitk_debug_taps(0)
    tok = ws[0][: 2 * 24 * 128 * 4].view(torch.float32).view(2, 24, 128).cpu().numpy()
    ref = z["blocks"][depth - 1]
    err = np.abs(tok - ref).max()
    print(f"depth {depth}: max|d| {err:.4f}  ref max {np.abs(ref).max():.2f}")
    assert err <= 0.02 * np.abs(ref).max() + 0.02


def _golden_inputs():
    pcm, length, _ = H.samples_int16()
    x16 = np.zeros((len(pcm), 16000), np.float32)
    for i in range(len(pcm)):
        n = min(16000, int(length[i]))
        x16[i, :n] = pcm[i, :n].astype(np.float32) / 32768.0
    return {"samples16k": x16, "noise": H.noise_clips(32), "adversarial": H.adversarial_batch(), "short2400": H.noise_clips(3, 2400, seed=11),
            "short14336": H.noise_clips(3, 14336, seed=12), "len16160": H.noise_clips(2, 16160, seed=14),
            "long10s": H.noise_clips(2, 160000, seed=13)}


@pytest.mark.parametrize("arch", H.ARCHS)
@pytest.mark.parametrize("kind", ["init", "trained"])
def test_bf16_scores_vs_reference_golden(arch, kind):
    """The shipped default path against the REAL reference's outputs on every golden input set, BASELINE config 1
    (samples/*.wav as a 1 s batch and at native length) included."""
    g = H.load_golden("probs.npz")
    m = _model(arch, kind, "bf16")
    pcm, length, _ = H.samples_int16()
    refs, gots = [], []
    for name, x in _golden_inputs().items():
        y = m(torch.from_numpy(x).to(DEV)).cpu().numpy()
        ref = g[f"{arch}/{kind}/{name}"]
        assert y.shape == ref.shape and np.isfinite(y).all(), name
        refs.append(ref); gots.append(y)
    nat = np.stack([m(torch.from_numpy(pcm[i, : length[i]].astype(np.float32)[None] / 32768.0).to(DEV)).cpu().numpy()[0]
                    for i in range(len(pcm))])
    refs.append(g[f"{arch}/{kind}/samples_native"]); gots.append(nat)
    ref, got = np.concatenate(refs), np.concatenate(gots)
    worst = float(np.abs(got - ref).max())
    rep = H.topk_report(ref, got, 5, eps_logit=EPS_LOGIT[kind])
    print(f"{arch}/{kind}: max|d prob| = {worst:.2e}  {rep}")
    assert worst <= TOL[kind], worst
    assert rep["max_dlogit"] <= EPS_LOGIT[kind], rep
    if kind == "trained":
        assert rep["decisive_match_frac"] == 1.0, rep          # literal top-5 wherever the reference is decisive
        must_nonempty = np.mean([(H.logits_of(r) > np.sort(H.logits_of(r))[-5] + 2 * EPS_LOGIT[kind]).any() for r in ref])
        assert must_nonempty >= 0.9 and rep["decisive_frac"] >= 0.5, (must_nonempty, rep)   # never vacuous
    else:
        assert H.tie_aware_topk_equal(ref, got, 5, eps=TOL[kind])


def test_bf16_matches_oracle_on_large_seeded_batch_with_ragged_tail():
    """517 clips = 103 full tiles of 5 clips + a ragged tile of 2; several tiles per CTA on a 148-SM part is
    covered by the 4096-clip bench parity check."""
    x = H.noise_clips(517, seed=77)
    sd = H.make_state_dict("uit_xs", "init")
    ref = O.forward(sd, torch.from_numpy(x)).numpy()
    y = _model("uit_xs", "init", "bf16")(torch.from_numpy(x).to(DEV)).cpu().numpy()
    assert np.abs(y - ref).max() <= TOL["init"]
    assert H.tie_aware_topk_equal(ref, y, 5, eps=TOL["init"])


def test_bf16_persistent_many_tiles_per_cta_equals_small_batches():
    """2000 clips = 400 tiles on <=148 CTAs (3 tiles per CTA): must equal the same clips run 5 at a time."""
    m = _model("uit_xxxs", "trained", "bf16")
    x = torch.from_numpy(H.noise_clips(2000, seed=3)).to(DEV)
    db, mp = m.front_end.logmel_unclamped(x)
    full = m.encode(db, mp)
    part = torch.cat([m.encode(db[i:i + 5].contiguous(), mp) for i in range(0, 50, 5)])
    assert torch.equal(full[:50], part)


@pytest.mark.parametrize("L", [2400, 4000, 8160, 14336, 16000, 16160, 16384, 33000])
def test_bf16_vs_fp32_device_paths_over_shapes(L):
    """Shape sweep on the device: the tensor-core path against the exact fp32 CUDA path for every token count (4..24 tokens,
    1..3 crops) and batch sizes that leave ragged tiles.  Covers the CUDA-core attention fallback (tokens != 24) too."""
    mb, mf = _model("uit_xxs", "init", "bf16"), _model("uit_xxs", "init", "fp32")
    for B in (1, 4, 6, 37):
        x = torch.from_numpy(H.noise_clips(B, L, seed=100 + B)).to(DEV)
        yb, yf = mb(x), mf(x)
        assert yb.shape == yf.shape == (B, 537) and torch.isfinite(yb).all()
        assert (yb - yf).abs().max().item() <= TOL["init"], (L, B)


def test_bf16_full_size_config3_periodicity():
    """BASELINE config 3 size on one GPU (UiT-XXS, 65 520 x 1 s clips in ONE forward): a batch made of 16 copies of a 4095-clip
    base (a multiple of the 5-clip tile, so every copy sits at the same in-tile positions and the batch maximum is the same)
    must give 16 bit-identical copies of the base scores - size-independent property at the full size, through every
    launch-splitting path of encode()."""
    m = _model("uit_xxs", "trained", "bf16")
    g = torch.Generator(device=DEV).manual_seed(11)
    base = (0.1 * torch.randn(4095, 16000, generator=g, device=DEV)).clamp_(-1, 1)
    base[7] *= 1e-4                                             # a quiet clip: large dynamic range inside the batch
    with torch.no_grad():
        y_base = m(base)
        y_big = m(base.repeat(16, 1))
    assert y_big.shape == (65520, 537) and torch.isfinite(y_big).all()
    assert torch.equal(y_big.view(16, 4095, 537), y_base.unsqueeze(0).expand(16, -1, -1))
