"""CPU-only tests of the boundary: the C-ABI library loads and exports every symbol include/uitk.h declares,
the host-side packers produce the documented layouts, and the Python mirror keeps the reference's module
contract.  No kernel is launched here."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from oracle import uit_oracle as O
from tests import helpers as H


@pytest.fixture(scope="module")
def lib():
    from uit_mobile_b200 import build
    from uit_mobile_b200 import _native as N
    build.build()
    return N.lib()


def test_every_declared_symbol_is_exported(lib):
    from uit_mobile_b200 import _native as N
    hdr = open(os.path.join(H.REPO, "include", "uitk.h")).read()
    declared = set(re.findall(r"UITK_API\s+[\w\s\*]+?\b(uitk_\w+)\s*\(", hdr))
    assert declared == set(N.SIGNATURES), declared ^ set(N.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name)
    assert lib.uitk_version() == 230


def test_geometry_helpers_match_oracle(lib):
    for L in (2400, 14336, 16000, 16160, 16384, 160000):
        T = O.num_frames(L)
        assert lib.uitk_num_frames(L) == T
        assert lib.uitk_num_crops(T, 102) == len(O.crop_starts(T, 102))
        tc = min(T, 102)
        assert lib.uitk_tokens_per_crop(T, 102) == 4 * ((tc - 16) // 16 + 1)


def test_frontend_pack_layout(lib):
    win, fb = O.hann_window(), O.melscale_fbanks_htk()
    n = lib.uitk_frontend_blob_bytes(fb.data_ptr())
    blob = np.zeros(n, np.uint8)
    assert lib.uitk_pack_frontend(win.data_ptr(), fb.data_ptr(), blob.ctypes.data, n) == 0
    i32, f32 = blob.view(np.int32), blob.view(np.float32)
    assert i32[0] == 0x55464535
    np.testing.assert_array_equal(f32[4:516], win.numpy())
    tw256 = f32[516:516 + 512].reshape(256, 2)
    j = np.arange(256)
    np.testing.assert_allclose(tw256[:, 0] + 1j * tw256[:, 1], np.exp(-2j * np.pi * ((j >> 4) * (j & 15)) / 256), atol=1e-7)
    base = 516 + 1024
    # tensor-core mel projection: per (mel octet, 8-bin group) block the mma.m16n8k8 B fragment (fp32; split in the kernel); every warp has a
    # list of blocks {group | fin << 8 | role << 9 | octet << 11 | aux << 14, block index}, runs of one octet end with fin
    nblk, blk = i32[base:base + 8], i32[base + 8:base + 8 + 8 * 41 * 2].reshape(8, 41, 2)
    frag = f32[base + 8 + 8 * 41 * 2:].reshape(-1, 32, 2)
    assert i32[1] + 1 == frag.shape[0] and (nblk >= 1).all() and (nblk <= 40).all() and (frag[-1] == 0).all()
    dense = np.zeros((8 * 33, 64), np.float64)
    used = np.zeros(frag.shape[0] - 1, int)
    runs = {o: [] for o in range(8)}                 # octet -> [(role, aux, blocks of the run)]
    rank = {1: 0, 0: 1, 2: 2}                        # producers, whole octets, owners: an owner never waits for a warp that waits
    for w in range(8):
        assert (blk[w, nblk[w]] == blk[w, nblk[w] - 1]).all()        # the kernel reads one entry ahead
        order, run = [], []
        for x, y in blk[w, :nblk[w]]:
            run.append((int(x) & 0xff, int(y)))
            if x & 0x100:
                role, o, aux = (int(x) >> 9) & 3, (int(x) >> 11) & 7, int(x) >> 14
                order.append(rank[role])
                runs[o].append((role, aux, run))
                run = []
            else:
                assert int(x) >> 8 == 0
        assert not run and order == sorted(order)
    slots = []
    for o, rs in runs.items():                       # an octet is whole, or one owner + its 1..2 producers
        prods = [(a & 7, a >> 3) for r, a, _ in rs if r == 1]
        if prods:
            own = [a for r, a, _ in rs if r == 2]
            assert len(own) == 1 and len(prods) <= 2 and len(rs) == len(prods) + 1
            assert all(n == len(prods) for _, n in prods) and own[0] & 3 == len(prods)
            assert [(own[0] >> 2) & 7, (own[0] >> 5) & 7][:len(prods)] == [sl for sl, _ in prods]
            slots += [sl for sl, _ in prods]
        else:
            assert [r for r, _, _ in rs] == [0]
        for _, _, run in rs:
            for g, b in run:
                used[b] += 1
                for lane in range(32):
                    tig, gid = lane & 3, lane >> 2
                    k0 = 8 * g + 2 * tig
                    assert dense[k0, 8 * o + gid] == 0 and dense[k0 + 1, 8 * o + gid] == 0
                    dense[k0, 8 * o + gid], dense[k0 + 1, 8 * o + gid] = frag[b, lane]
    assert (used == 1).all() and sorted(slots) == list(range(len(slots))) and len(slots) <= 8
    assert nblk.max() <= 1.5 * nblk.mean()                           # the blocks are spread evenly over the warps
    want = 0.25 * fb.numpy().astype(np.float64)                      # weights carry the 1/4 of the kernel's 4|X|^2
    np.testing.assert_array_equal(dense[:257], want)
    assert (dense[257:] == 0).all() and frag.shape[0] <= 49          # HTK/64: 40 blocks of the 264 possible
    # error path: blob too small
    assert lib.uitk_pack_frontend(win.data_ptr(), fb.data_ptr(), blob.ctypes.data, 16) == -5
    assert b"needs" in lib.uitk_last_error()


def test_encoder_pack_tensor_order_and_size(lib):
    from uit_mobile_b200 import _native as N
    names = N.encoder_tensor_names(4)
    sd = H.make_state_dict("uit_xxxs")
    assert len(names) == 16 + 12 * 4 and all(n in sd for n in names)
    dead = set(sd) - set(names)
    assert dead == {"front_end.0.spectrogram.window", "front_end.0.mel_scale.fb", "init_bn.1.num_batches_tracked"}
    cfg = N.EncoderCfg(4, 537, 6, 0)
    assert lib.uitk_encoder_blob_bytes(C.byref(cfg)) > 4 * 568089 * 0.9
    bad = N.EncoderCfg(4, 5000, 6, 0)
    assert lib.uitk_encoder_blob_bytes(C.byref(bad)) == 0 and b"outputdim" in lib.uitk_last_error()
    bad = N.EncoderCfg(4, 537, 6, 0, attention=2)
    assert lib.uitk_encoder_blob_bytes(C.byref(bad)) == 0 and b"attention" in lib.uitk_last_error()
    # variants: the full Attention blob is larger (qkv 384 x 128, proj 128 x 128); only the UiT configuration carries a bf16 section
    full = N.EncoderCfg(4, 537, 6, 1, attention=1)
    assert lib.uitk_encoder_blob_bytes(C.byref(full)) > lib.uitk_encoder_blob_bytes(C.byref(cfg)) + 4 * 4 * (288 * 128 + 96 * 128)
    assert N.EncoderCfg(4, 537, 6, 1).tensor_core and not full.tensor_core and not N.EncoderCfg(4, 537, 6, 1, pooling=1).tensor_core
    for pooling, extra in ((0, 0), (1, 1), (2, 0)):
        c = N.EncoderCfg(4, 537, 6, 0, pooling=pooling)
        assert lib.uitk_tokens_total(C.byref(c), 101, 102) == 24 + extra and lib.uitk_tokens_total(C.byref(c), 90, 102) == 20 + extra


def _bf16_to_f32(u16: np.ndarray) -> np.ndarray:
    return (u16.astype(np.uint32) << 16).view(np.float32)


def test_encoder_pack_bf16_section_layout(lib):
    """The tensor-core blob: weights as K-major core-matrix tiles in the kernel's consumption order, every Linear bias as a
    [N x 8] bf16 bias tile whose hi + mid + lo columns reproduce the fp32 value (csrc/pack.cu, csrc/encoder_tc.cu)."""
    from uit_mobile_b200 import _native as N
    depth = 4
    names = N.encoder_tensor_names(depth)
    sd = H.make_state_dict("uit_xxxs", "trained")
    host = [sd[n].detach().to(torch.float32).contiguous() for n in names]
    ptrs = (C.c_void_p * len(host))(*[h.data_ptr() for h in host])
    cfg32, cfg16 = N.EncoderCfg(depth, 537, 6, 0), N.EncoderCfg(depth, 537, 6, 1)
    n32, n16 = lib.uitk_encoder_blob_bytes(C.byref(cfg32)), lib.uitk_encoder_blob_bytes(C.byref(cfg16))
    patch_bytes, qkv_half, qkv_bias, proj, tile, btile, fc1_bias = 4 * 16384, 96 * 64 * 2, 96 * 16, 128 * 32 * 2, 16384, 2048, 1024
    block_bytes = 2 * qkv_half + qkv_bias + proj + btile + 6 * (tile + fc1_bias) + 6 * tile + btile
    assert n16 - n32 == patch_bytes + depth * block_bytes
    blob = np.zeros(n16, np.uint8)
    assert lib.uitk_pack_encoder(C.byref(cfg16), ptrs, blob.ctypes.data, n16) == 0, lib.uitk_last_error()
    sec = blob[n32:].view(np.uint16)                          # bf16 section (the fp32 section is padded to 1 KB, as n32 is)

    def kmajor(off_bytes, n, k):                              # [(k/8), n, 8] -> [n, k]
        t = sec[off_bytes // 2: off_bytes // 2 + n * k].reshape(k // 8, n, 8)
        return _bf16_to_f32(t.transpose(1, 0, 2).reshape(n, k))

    def bias_tile(off_bytes, n):
        t = _bf16_to_f32(sec[off_bytes // 2: off_bytes // 2 + n * 8].reshape(n, 8))
        assert (t[:, 3:] == 0).all()
        return t[:, 0].astype(np.float64) + t[:, 1] + t[:, 2]

    bf = lambda w: _bf16_to_f32((w.numpy().view(np.uint32) + 0x7FFF + ((w.numpy().view(np.uint32) >> 16) & 1) >> 16).astype(np.uint16))
    # patch weight: 4 K-quarters of [128 x 64]
    pw = sd["patch_embed.proj.weight"].reshape(128, 256)
    for c in range(4):
        np.testing.assert_array_equal(kmajor(c * 16384, 128, 64), bf(pw[:, c * 64:(c + 1) * 64].contiguous()))
    # block 1 (second block): LayerNorm affine folded into qkv / fc1, biases as tiles
    i = 1
    off = patch_bytes + i * block_bytes
    g1, b1 = sd[f"blocks.{i}.norm1.weight"].double(), sd[f"blocks.{i}.norm1.bias"].double()
    wq, bq = sd[f"blocks.{i}.attn.qkv.weight"].double(), sd[f"blocks.{i}.attn.qkv.bias"].double()
    np.testing.assert_allclose(kmajor(off, 96, 64), (wq * g1)[:, :64].float().numpy(), rtol=2 ** -8, atol=1e-30)
    np.testing.assert_allclose(bias_tile(off + qkv_half, 96), (bq + wq @ b1).numpy(), rtol=1e-6, atol=1e-7)
    off += 2 * qkv_half + qkv_bias
    np.testing.assert_array_equal(kmajor(off, 128, 32), bf(sd[f"blocks.{i}.attn.proj.weight"]))
    np.testing.assert_allclose(bias_tile(off + proj, 128), sd[f"blocks.{i}.attn.proj.bias"].double().numpy(), rtol=1e-6, atol=1e-7)
    off += proj + btile
    # MLP slots in issue order: fc1[0] fc1[1] | fc1[2] fc2[0] | fc1[3] fc2[1] | fc1[4] fc2[2] | fc1[5] fc2[3] | fc2[4] | fc2[5]
    g2, b2 = sd[f"blocks.{i}.norm2.weight"].double(), sd[f"blocks.{i}.norm2.bias"].double()
    w1, bb1 = sd[f"blocks.{i}.mlp.fc1.weight"].double(), sd[f"blocks.{i}.mlp.fc1.bias"].double()
    w2, bb2 = sd[f"blocks.{i}.mlp.fc2.weight"], sd[f"blocks.{i}.mlp.fc2.bias"].double()
    order = [("fc1", 0), ("fc1", 1)] + [x for c in range(1, 6) for x in ((("fc1", c + 1),) if c < 5 else ()) + (("fc2", c - 1),)] + [("fc2", 5)]
    assert [o for o in order if o[0] == "fc1"] == [("fc1", c) for c in range(6)] and len(order) == 12
    for kind, c in order:
        if kind == "fc1":
            np.testing.assert_allclose(kmajor(off, 64, 128), (w1 * g2)[64 * c:64 * c + 64].float().numpy(), rtol=2 ** -8, atol=1e-30)
            np.testing.assert_allclose(bias_tile(off + tile, 64), (bb1 + w1 @ b2)[64 * c:64 * c + 64].numpy(), rtol=1e-6, atol=1e-7)
            off += tile + fc1_bias
        else:
            np.testing.assert_array_equal(kmajor(off, 128, 64), bf(w2[:, 64 * c:64 * c + 64].contiguous()))
            off += tile
            if c == 0:
                np.testing.assert_allclose(bias_tile(off, 128), bb2.numpy(), rtol=1e-6, atol=1e-7)
                off += btile
    assert off == patch_bytes + (i + 1) * block_bytes


def test_launch_entry_points_validate_before_touching_cuda(lib):
    assert lib.uitk_logmel(None, 1, 16000, 16000, None, None, None, None, None) == -1
    one = C.c_float(0)
    p = C.addressof(one)
    assert lib.uitk_logmel(p, 1, 100, 100, p, p, p, None, None) == -1 and b"reflect" in lib.uitk_last_error()
    # sliding windows: hop / window must be multiples of the STFT hop, the stream at least one window long
    buf = np.zeros(64, np.float32)
    p = (buf.ctypes.data + 15) // 16 * 16                       # 16-byte aligned dummy pointer (never dereferenced)
    assert lib.uitk_logmel_sliding(p, 160000, 16000, 1000, p, p, p, None, p, 1 << 20, None) == -1 and b"multiple" in lib.uitk_last_error()
    assert lib.uitk_logmel_sliding(p, 160000, 16001, 1600, p, p, p, None, p, 1 << 20, None) == -1
    assert lib.uitk_logmel_sliding(p, 8000, 16000, 1600, p, p, p, None, p, 1 << 20, None) == -1 and b"shorter" in lib.uitk_last_error()
    assert lib.uitk_logmel_sliding_workspace_bytes(160000) == 64 * 1001 * 4
    assert lib.uitk_logmel_sliding(p, 160000, 16000, 1600, p, p, p, None, p, 16, None) < 0 and b"workspace" in lib.uitk_last_error()


def test_module_contract():
    import uit_mobile_b200 as U
    assert set(U.models.PRETRAINED_CHECKPOINTS) == {"uit_xs", "uit_xxs", "uit_xxxs"}
    for arch in H.ARCHS:
        m = getattr(U.models, arch)(**U.models.PRETRAINED_CHECKPOINTS[arch]["model_kwargs"])
        want = [l.rstrip("\n").split("\t")[1:] for l in open(H.GOLDEN_DIR + "/state_dict_layout.txt") if l.startswith(arch + "\t")]
        got = [[k, str(tuple(v.shape)), str(v.dtype)] for k, v in m.state_dict().items()]
        assert got == want
        m.load_state_dict(H.make_state_dict(arch), strict=True)
        assert m.target_length == 102 and m.hop_size == 160 and m.n_mels == 64 and m.outputdim == 537
        assert m.patch_embed.grid_size == (4, 6) and m.pooling == "mean" and m.eval_avg == "mean"
        assert m.no_weight_decay() == {"time_pos_embed", "cls_token", "freq_pos_embed", "token_pos_embed"}
        assert sum(p.numel() for p in m.parameters()) == {"uit_xs": 1495577, "uit_xxs": 799961, "uit_xxxs": 568089}[arch]


def test_tile_clips_and_fused_stage_methods(lib):
    import uit_mobile_b200 as U
    m = U.models.uit_xxxs(outputdim=537, target_length=102)
    from uit_mobile_b200 import _native as N
    assert m.tile_clips(101) == 5            # 1 s clip: 24 token slots -> 5 clips per 128-row tile
    assert m.tile_clips(1001) == 1           # 10 s clip: 10 crops -> a clip's crops fill whole tiles
    assert m.tile_clips(16) == 5             # 2400 samples: still 24 slots per clip (4 live), 5 clips per tile
    m.eval()
    with pytest.raises(N.UitkError):         # forward_features / forward_head are real entry points now: CUDA only, no fallback
        m.forward_features(torch.zeros(1, 1, 64, 102))
    with pytest.raises(N.UitkError):
        m.forward_head(torch.zeros(1, 24, 128))
    with pytest.raises(N.UitkError):
        m.init_bn(torch.zeros(1, 1, 64, 102))


def test_variant_factories_keep_the_reference_state_dict():
    """full Attention / GELU / pooling 'token' | 'dm' behind the same factories; shapes as the reference builds them
    (tests/golden/generate_golden.py loads the same state_dicts into the reference with strict=True)."""
    import uit_mobile_b200 as U
    for name, (depth, attention, act, pooling) in H.VARIANTS.items():
        m = H.build_variant(U.models, name)
        sd = H.make_state_dict(name)
        assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == [(k, tuple(v.shape)) for k, v in sd.items()], name
        m.load_state_dict(sd, strict=True)
        assert m.pooling == pooling and m.attention_type == attention and len(m.blocks) == depth
        assert isinstance(m.blocks[0].mlp.act, torch.nn.ReLU if act == "relu" else torch.nn.GELU)
        assert not m._cfg().tensor_core
    assert U.models.uit_xs(target_length=102)._cfg().tensor_core and not U.models.uit_xs(target_length=102, precision="fp32")._cfg().tensor_core


def test_pos_embed_resize_on_load():
    import uit_mobile_b200 as U
    sd = H.make_state_dict("uit_xxxs", grid_t=6)
    m = U.models.uit_xxxs(outputdim=537, target_length=64)          # 4 time patches: slice
    m.load_state_dict(sd, strict=True)
    assert torch.equal(m.time_pos_embed, sd["time_pos_embed"][..., :4])
    sd3 = H.make_state_dict("uit_xxxs", grid_t=3)
    m = U.models.uit_xxxs(outputdim=537, target_length=102)         # 3 -> 6: bilinear, like uit.py:433-438
    m.load_state_dict(sd3, strict=True)
    want = torch.nn.functional.interpolate(sd3["time_pos_embed"], size=(1, 6), align_corners=False, mode="bilinear")
    assert torch.equal(m.time_pos_embed, want)


def test_unsupported_configurations_raise():
    import uit_mobile_b200 as U
    with pytest.raises(NotImplementedError):
        U.models.UITBase()                                            # reference defaults: 768-dim token-pooled ViT
    with pytest.raises(KeyError):
        U.models.audio_transformer_h128_d3_m3_bneck_v2_relu()         # names an attention class the reference never defines (Q10)
    with pytest.raises(NotImplementedError):
        U.models.uit_xs(embed_dim=192)
    with pytest.raises(NotImplementedError):
        U.models.uit_xs(n_mels=80)
    with pytest.raises(NotImplementedError):
        U.models.uit_xs(target_length=1012)


def test_no_cpu_fallback():
    import uit_mobile_b200 as U
    from uit_mobile_b200 import _native as N
    m = U.models.uit_xxxs(outputdim=537, target_length=102).eval()
    with pytest.raises(N.UitkError):
        m(torch.zeros(1, 16000))
    with pytest.raises(RuntimeError):
        m.blocks(torch.zeros(1, 24, 128))
    src = open(os.path.join(H.REPO, "uit_mobile_b200", "models", "uit.py")).read()
    assert "oracle" not in src.replace("no CPU path", "")


def test_mobilenetv2_module_contract(lib):
    from uit_mobile_b200 import _native as N
    """models.MobileNetV2: the reference's state_dict keys / shapes / dtypes in the reference's order (state_dict_layout_mnv2.txt is
    the reference's own list), strict loading of a full state_dict, the kernel's tensor order, and loud refusals."""
    import uit_mobile_b200 as U
    m = U.models.MobileNetV2(outputdim=537)
    want = [l.rstrip("\n").split("\t")[1:] for l in open(os.path.join(H.GOLDEN_DIR, "state_dict_layout_mnv2.txt"))]
    got = [[k, str(tuple(v.shape)), str(v.dtype)] for k, v in m.state_dict().items()]
    assert got == want
    m.load_state_dict(H.make_mnv2_state_dict("trained"), strict=True)
    names = N.mnv2_tensor_names()
    assert len(names) == lib.uitk_mnv2_num_tensors() == 52 * 5 + 2 and set(names) <= set(m.state_dict())
    assert names[0] == "features.0.0.weight" and names[-2:] == ["classifier.1.weight", "classifier.1.bias"]
    assert lib.uitk_mnv2_blob_bytes(537) > 4 * (1280 * 537 + 320 * 1280) and lib.uitk_mnv2_workspace_bytes(4, 101) > 0
    with pytest.raises(N.UitkError):
        m.eval()(torch.zeros(2, 16000))                  # CUDA only, no fallback
    with pytest.raises(NotImplementedError):
        m.train()(torch.zeros(2, 16000))
    for kw in (dict(width_mult=0.5), dict(last_channel=640), dict(inverted_residual_setting=[[1, 16, 1, 1]]), dict(n_mels=80)):
        with pytest.raises(NotImplementedError):
            U.models.MobileNetV2(**kw)
