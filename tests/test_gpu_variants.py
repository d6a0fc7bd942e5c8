"""GPU tests of the round-2 boundary rows: forward_features / forward_head as real entry points (uit.py:379-412), the UITBase
variants behind the same factories (full Attention, GELU, pooling 'token' | 'dm'), the device-conditional exact
re-run that takes the max all-reduce off the sharded critical path, and short clips on the tensor-core megakernel."""
import numpy as np
import pytest
import torch

from oracle import uit_oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
FP32_TOL = {"init": 5e-5, "trained": 2e-4}


def _uit(arch, kind, precision, **kw):
    import uit_mobile_b200 as U
    m = getattr(U.models, arch)(outputdim=537, target_length=102, precision=precision, **kw)
    m.load_state_dict(H.make_state_dict(arch, kind), strict=True)
    return m.to(DEV).eval()


@pytest.mark.parametrize("name", list(H.VARIANTS))
@pytest.mark.parametrize("kind", ["init", "trained"])
def test_variants_vs_reference_golden(name, kind):
    """Each variant through the reference's own factory on CPU (golden) vs the fp32 CUDA kernels, incl. 20-token clips and the
    10-crop branch.  precision='bf16' is accepted and runs the same fp32 kernels (only the UiT configuration has a megakernel)."""
    import uit_mobile_b200 as U
    g = H.load_golden("probs_variants.npz")
    for precision in ("fp32", "bf16"):
        m = H.build_variant(U.models, name, precision=precision)
        m.load_state_dict(H.make_state_dict(name, kind), strict=True)
        m = m.to(DEV).eval()
        worst = 0.0
        for k, x in H.variant_inputs().items():
            y = m(torch.from_numpy(x).to(DEV)).cpu().numpy()
            ref = g[f"{name}/{kind}/{k}"]
            assert y.shape == ref.shape and np.isfinite(y).all()
            worst = max(worst, float(np.abs(y - ref).max()))
        print(f"{name}/{kind}/{precision}: max|d prob| = {worst:.2e}")
        assert worst <= FP32_TOL[kind], (name, kind, worst)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_forward_features_and_head_vs_reference_trace(precision):
    """forward_features(bn) and forward_head(features) against the reference's own method outputs (trace_xxxs.npz)."""
    z = H.load_golden("trace_xxxs.npz")
    m = _uit("uit_xxxs", "trained", precision)
    bn = torch.from_numpy(z["bn"][:, None, :, :102].copy()).to(DEV)          # [2, 1, 64, 101] as the reference passes it
    feat = m.forward_features(bn)
    assert feat.shape == (2, 24, 128)
    tol_f = 2e-4 if precision == "fp32" else 0.05 * float(np.abs(z["features"]).max())
    assert np.abs(feat.cpu().numpy() - z["features"]).max() <= tol_f
    probs = m.forward_head(torch.from_numpy(z["features"]).to(DEV))          # the head alone is fp32 in both precisions
    assert np.abs(probs.cpu().numpy() - z["probs"]).max() <= 2e-5
    # init_bn as a callable module: eval BatchNorm over the mel axis
    db = torch.from_numpy(z["db"]).to(DEV)
    got_bn = m.init_bn(db.unsqueeze(1)).squeeze(1).cpu().numpy()
    np.testing.assert_allclose(got_bn, z["bn"], atol=2e-5, rtol=0)
    # composition == forward, within the precision's tolerance (forward fuses the clamp, BatchNorm, pooling and head)
    x = torch.from_numpy(H.noise_clips(32)[:2]).to(DEV)
    composed = m.forward_head(m.forward_features(m.init_bn(m.front_end(x).unsqueeze(1))))
    assert (composed - m(x)).abs().max().item() <= (2e-5 if precision == "fp32" else 2e-3)


@pytest.mark.parametrize("L", [2400, 8160, 14336, 14592])
def test_forward_features_short_clips_and_token_pooling(L):
    """Token counts 4 / 12 / 20 (+ cls for pooling='token') against the oracle."""
    import uit_mobile_b200 as U
    x = torch.from_numpy(H.noise_clips(3, L, seed=L))
    for name, precision in (("uit_xxxs_gelu_token", "fp32"), ("uit_xxxs", "fp32"), ("uit_xxxs", "bf16")):
        sd = H.make_state_dict(name, "init")
        if name in H.VARIANTS:
            m, (_, _, act, pooling) = H.build_variant(U.models, name, precision=precision), H.VARIANTS[name]
        else:
            m, act, pooling = U.models.uit_xxxs(outputdim=537, target_length=102, precision=precision), "relu", "mean"
        m.load_state_dict(sd, strict=True)
        m = m.to(DEV).eval()
        xn = O.init_bn(O.logmel(x, sd["front_end.0.spectrogram.window"], sd["front_end.0.mel_scale.fb"]), sd)
        ref = O.features(xn, sd, None, act, pooling)
        got = m.forward_features(xn.unsqueeze(1).to(DEV))
        assert got.shape == ref.shape
        assert (got.cpu() - ref).abs().max().item() <= (1e-4 if precision == "fp32" else 0.03)
        assert (m.forward_head(ref.to(DEV)).cpu() - O.head(ref, sd, pooling)).abs().max().item() <= 2e-5


def test_conditional_fixup_is_exact_and_skips_when_not_needed():
    """uitk_encoder_fixup: a speculative encode with a too-low maximum (what a rank does while the all-reduce is in flight) followed by
    the device-conditional re-run equals the encode with the true maximum; when no value lies under the true cutoff, or the maximum
    did not change, the re-run's kernels return at once and leave the scores untouched."""
    m = _uit("uit_xxs", "trained", "bf16")
    x = torch.from_numpy(np.concatenate([H.noise_clips(7, seed=41, amp=1e-3), H.adversarial_batch()[:1]])).to(DEV)   # quiet clips + digital silence
    words = m._new_words(DEV)
    db, _ = m.front_end.logmel_unclamped(x, max_pow=words[0:1], min_pow=words[1:2])
    true_max = torch.tensor([np.float32(1e4).view(np.int32)], dtype=torch.int32, device=DEV)        # another rank holds a loud clip (+40 dB)
    want = m.encode(db, true_max)
    spec = m.encode(db, words[0:1])
    assert not torch.equal(spec, want)                      # the silent clip is clamped differently under the global cutoff
    m.encode(db, true_max, out=spec, fixup=(words[0:1], words[1:2]))
    assert torch.equal(spec, want)
    # not needed (1): same maximum
    sentinel = torch.full_like(want, -7.0)
    m.encode(db, words[0:1], out=sentinel, fixup=(words[0:1], words[1:2]))
    assert (sentinel == -7.0).all()
    # not needed (2): higher global maximum but nothing of this rank lies under the true cutoff
    xq = torch.from_numpy(H.noise_clips(6, seed=43, amp=1e-2)).to(DEV)
    wq = m._new_words(DEV)
    dbq, _ = m.front_end.logmel_unclamped(xq, max_pow=wq[0:1], min_pow=wq[1:2])
    sentinel = torch.full((6, 537), -7.0, device=DEV)
    m.encode(dbq, true_max, out=sentinel, fixup=(wq[0:1], wq[1:2]))
    assert (sentinel == -7.0).all()
    assert torch.equal(m.encode(dbq, true_max), m.encode(dbq, wq[0:1]))


def test_sharded_forward_world1_nccl_and_empty_shard():
    """model.process_group set (single-rank NCCL group): the speculative encode + async all-reduce + conditional re-run path gives
    exactly the unsharded scores, an EMPTY shard joins the collective instead of hanging its peers, and the fp32
    configuration takes the blocking all-reduce."""
    import torch.distributed as dist
    if not dist.is_initialized():
        import os, socket
        s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device(DEV))
    try:
        x = torch.from_numpy(np.concatenate([H.noise_clips(5, seed=41, amp=1e-3), H.adversarial_batch()])).to(DEV)
        for precision in ("bf16", "fp32"):
            m = _uit("uit_xxxs", "trained", precision)
            want = m(x)
            m.process_group = dist.group.WORLD
            assert torch.equal(m(x), want)
            assert tuple(m(torch.zeros(0, 16000, device=DEV)).shape) == (0, 537)
            stream = torch.from_numpy(H.noise_clips(1, 16000 * 4, seed=3)[0]).to(DEV)
            got = m.forward_sliding(stream, hop=1600)
            m.process_group = None
            assert torch.equal(got, m.forward_sliding(stream, hop=1600))
        # the same with the maximum word exchanged through peer memory (sharding.PeerWords) instead of the NCCL all-reduce:
        # ten steps (the epoch ring wraps), an empty shard in between
        from uit_mobile_b200 import sharding
        m = _uit("uit_xxxs", "trained", "bf16")
        want = m(x)
        m.process_group = dist.group.WORLD
        m.peer_words = sharding.PeerWords(dist.group.WORLD, torch.device(DEV))
        for i in range(10):
            assert torch.equal(m(x), want)
            if i == 4:
                assert tuple(m(torch.zeros(0, 16000, device=DEV)).shape) == (0, 537)
        assert m.peer_words.epoch == 11
        torch.cuda.synchronize()
    finally:
        dist.destroy_process_group()


def test_sliding_hop_equal_to_window_scopes_the_cutoff_to_window_frames():
    """with hop == window some stream frames straddle two windows and belong to none; they must not raise the
    batch-global top-dB maximum.  forward_sliding falls back to the per-window front-end there."""
    m = _uit("uit_xxxs", "trained", "bf16")
    stream = torch.from_numpy(H.noise_clips(1, 16000 * 3, seed=9, amp=1e-3)[0]).to(DEV)
    stream[16000 - 40: 16000 + 40] = 0.9             # a loud click exactly on a window boundary
    stream[100:4000] = 0.0                           # digital silence inside window 0: the cutoff matters
    want = m(stream.unfold(0, 16000, 16000).contiguous())
    assert torch.equal(m.forward_sliding(stream, hop=16000), want)
    db_a, mp_a = m.front_end.logmel_unclamped(stream, ld=16000, B=3, L=16000)
    db_b, mp_b = m.front_end.logmel_sliding(stream, 16000, 16000)
    assert torch.equal(db_a, db_b) and torch.equal(mp_a, mp_b)


def test_encode_validates_its_arguments():
    m = _uit("uit_xxxs", "trained", "bf16")
    x = torch.from_numpy(H.noise_clips(4, seed=3)).to(DEV)
    db, mp = m.front_end.logmel_unclamped(x)
    with pytest.raises(ValueError):
        m.encode(db[..., :96], mp)                   # non-contiguous view
    with pytest.raises(ValueError):
        m.encode(db.double(), mp)
    with pytest.raises(ValueError):
        m.encode(db, mp.float())
    from uit_mobile_b200 import _native as N
    with pytest.raises(N.UitkError):
        m.encode(db.cpu(), mp)
    m.train()
    try:
        with pytest.raises(NotImplementedError):
            m.encode(db, mp)
    finally:
        m.eval()


@pytest.mark.parametrize("kind", ["init", "trained"])
def test_mobilenetv2_vs_reference_golden(kind):
    """models.MobileNetV2 (the reference's teacher, mobilenetv2.py:66-178) on the fp32 CUDA kernels against the reference's own
    outputs (tests/golden/mobilenetv2.npz): 1 s / short / 1.01 s / 10 s clips and the adversarial batch (batch-global top-dB
    clamp).  53 fp32 convolution layers with folded BatchNorm: |d prob| <= 5e-5 (measured ~1e-6 on noise; the last feature map agrees
    to 3.5e-5 of a 0..6 range).  The adversarial batch gets 1e-3: its silent / clamped clips are CONSTANT inputs at the top-dB cutoff, so
    the front-end's ~1e-4 dB agreement on the batch maximum (test_logmel_q2_batch_global_cutoff) shifts every pixel of those clips
    the same way.  Literal (tie-aware) top-5 on the sparse 'trained' head."""
    import uit_mobile_b200 as U
    m = U.models.MobileNetV2(outputdim=H.OUTPUTDIM)
    m.load_state_dict(H.make_mnv2_state_dict(kind), strict=True)
    m = m.to(DEV).eval()
    g = H.load_golden("mobilenetv2.npz")
    inputs = {"noise": H.noise_clips(6), "adversarial": H.adversarial_batch(), "short2400": H.noise_clips(3, 2400, seed=11),
              "len16160": H.noise_clips(2, 16160, seed=14), "long10s": H.noise_clips(2, 160000, seed=13)}
    for name, x in inputs.items():
        with torch.no_grad():
            y = m(torch.from_numpy(x).to(DEV)).cpu().numpy()
        ref = g[f"{kind}/{name}"]
        assert y.shape == ref.shape
        err = float(np.abs(y - ref).max())
        print(f"  mobilenetv2/{kind}/{name}: max|d prob| = {err:.2e}")
        assert err <= (1e-3 if name == "adversarial" else 5e-5), (name, err)
        if kind == "trained" and name in ("noise", "len16160", "long10s"):
            assert H.tie_aware_topk_equal(ref, y, 5, eps=3e-4), name
    # a batch larger than the kernel's 256-clip pass, ragged: chunks must not interact (the top-dB scope aside, which is the batch's)
    xb = torch.from_numpy(H.noise_clips(300, 4000, seed=9)).to(DEV)
    with torch.no_grad():
        full = m(xb)
        db = m.front_end(xb)
    assert full.shape == (300, H.OUTPUTDIM) and torch.isfinite(full).all()
    assert tuple(m(torch.zeros(0, 16000, device=DEV)).shape) == (0, H.OUTPUTDIM)
