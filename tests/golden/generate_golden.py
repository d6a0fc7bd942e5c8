"""Generate the committed golden vectors from the REAL reference (build container only).

    python tests/golden/generate_golden.py

Imports /root/reference/models (uit.py) under an in-memory ``timm`` shim (timm is not installed; uit.py:8-9
only needs to_2tuple / DropPath / trunc_normal_), loads the seeded state_dicts of ``tests/helpers.py`` with
``strict=True`` (which also pins the state_dict key/shape contract), runs the reference's own eval forward on
CPU and stores inputs that cannot be regenerated (the sample wavs, as int16) and the reference outputs.  While
doing so it asserts that ``oracle/uit_oracle.py`` reproduces the reference, i.e. it is the script that pins
the oracle.  /root/reference does not exist on the GPU box; the tests only read the .npz files.
"""
from __future__ import annotations

import glob
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
REFERENCE = "/root/reference"


def install_timm_shim():
    t = types.ModuleType("timm"); m = types.ModuleType("timm.models")
    l = types.ModuleType("timm.models.layers"); h = types.ModuleType("timm.models.layers.helpers")
    h.to_2tuple = lambda x: tuple(x) if isinstance(x, (tuple, list)) else (x, x)

    class DropPath(nn.Module):
        def __init__(self, p=0.0):
            super().__init__()

        def forward(self, x):
            return x
    l.DropPath = DropPath; l.trunc_normal_ = nn.init.trunc_normal_; l.helpers = h
    for n, mod in (("timm", t), ("timm.models", m), ("timm.models.layers", l), ("timm.models.layers.helpers", h)):
        sys.modules[n] = mod


def import_reference():
    install_timm_shim()
    sys.path.insert(0, REFERENCE)
    import models as ref_models          # /root/reference/models/__init__.py
    return ref_models


def read_samples():
    from scipy.io import wavfile
    names, pcm, length = [], [], []
    for f in sorted(glob.glob(os.path.join(REFERENCE, "samples", "*.wav"))):
        sr, x = wavfile.read(f)
        assert sr == 16000 and x.dtype == np.int16 and x.ndim == 1
        buf = np.zeros(16384, np.int16); buf[:len(x)] = x
        names.append(os.path.basename(f)); pcm.append(buf); length.append(len(x))
    return names, np.stack(pcm), np.array(length, np.int64)


def main():
    from tests import helpers as H
    from oracle import uit_oracle as O
    from oracle import logmel_f64 as O64
    ref_models = import_reference()
    torch.manual_seed(0)
    torch.set_num_threads(8)

    names, pcm, length = read_samples()
    np.savez_compressed(os.path.join(HERE, "samples_int16.npz"), pcm=pcm, length=length, names=np.array(names))

    def crop16k(i):
        x = pcm[i, :length[i]].astype(np.float32) / 32768.0
        out = np.zeros(16000, np.float32); n = min(16000, len(x)); out[:n] = x[:n]
        return out
    inputs = {
        "samples16k": np.stack([crop16k(i) for i in range(len(names))]),
        "noise": H.noise_clips(32),
        "adversarial": H.adversarial_batch(),
        "short2400": H.noise_clips(3, 2400, seed=11),
        "short14336": H.noise_clips(3, 14336, seed=12),
        "len16160": H.noise_clips(2, 16160, seed=14),          # evaluate.py:56-64 zero-pad length, T=102
        "long10s": H.noise_clips(2, 160000, seed=13),
    }
    native = [pcm[i, :length[i]].astype(np.float32)[None] / 32768.0 for i in range(len(names))]

    # ---- buffers: our analytic window / fb must equal the reference module's (Q9)
    m0 = ref_models.uit_xxxs(outputdim=537, target_length=102)
    sd0 = m0.state_dict()
    assert torch.equal(sd0["front_end.0.spectrogram.window"], O.hann_window())
    assert torch.equal(sd0["front_end.0.mel_scale.fb"], O.melscale_fbanks_htk())
    ref_keys = {a: [(k, tuple(v.shape), str(v.dtype)) for k, v in
                    getattr(ref_models, a)(outputdim=537, target_length=102).state_dict().items()] for a in H.ARCHS}

    # ---- front-end goldens (weights-independent)
    fe = {}
    worst = 0.0
    for k, x in inputs.items():
        with torch.no_grad():
            ref_db = m0.front_end(torch.from_numpy(x))
        odb = O.logmel(torch.from_numpy(x), sd0["front_end.0.spectrogram.window"], sd0["front_end.0.mel_scale.fb"])
        assert torch.equal(ref_db, odb), f"oracle logmel != reference on {k}: {(ref_db - odb).abs().max()}"
        d64 = O64.logmel(x, sd0["front_end.0.spectrogram.window"].numpy(), sd0["front_end.0.mel_scale.fb"].numpy())
        worst = max(worst, float(np.abs(d64 - ref_db.numpy()).max()))
        fe[k] = ref_db.numpy()
    print(f"reference(fp32) vs float64 restatement: max |d dB| = {worst:.3e}")
    np.savez_compressed(os.path.join(HERE, "logmel.npz"), **{k: v for k, v in fe.items() if k != "long10s"},
                        long10s=fe["long10s"][:1])

    # ---- end-to-end goldens per arch / weight kind
    out = {}
    for arch in H.ARCHS:
        for kind in ("init", "trained"):
            sd = H.make_state_dict(arch, kind)
            model = getattr(ref_models, arch)(outputdim=537, target_length=102)
            assert [(k, tuple(v.shape), str(v.dtype)) for k, v in sd.items()] == ref_keys[arch], "state_dict layout drifted"
            model.load_state_dict(sd, strict=True)
            model.eval()
            for k, x in inputs.items():
                with torch.no_grad():
                    r = model(torch.from_numpy(x))
                o = O.forward(sd, torch.from_numpy(x))
                err = float((r - o).abs().max())
                assert err <= 2e-6, f"oracle != reference: {arch}/{kind}/{k}: {err}"
                out[f"{arch}/{kind}/{k}"] = r.numpy()
            nat = []
            for x in native:
                with torch.no_grad():
                    r = model(torch.from_numpy(x))
                o = O.forward(sd, torch.from_numpy(x))
                assert float((r - o).abs().max()) <= 2e-6
                nat.append(r.numpy()[0])
            out[f"{arch}/{kind}/samples_native"] = np.stack(nat)
            if kind == "trained":
                model.eval_avg = "max"
                with torch.no_grad():
                    r = model(torch.from_numpy(inputs["long10s"]))
                o = O.forward(sd, torch.from_numpy(inputs["long10s"]), eval_avg="max")
                assert float((r - o).abs().max()) <= 2e-6
                out[f"{arch}/{kind}/long10s_max"] = r.numpy()
            print(arch, kind, "ok")
    np.savez_compressed(os.path.join(HERE, "probs.npz"), **out)

    # ---- UITBase variants: full Attention, GELU, pooling 'token' / 'dm', through the reference's own factories
    vout = {}
    for name, (depth, attention, act, pooling) in H.VARIANTS.items():
        for kind in ("init", "trained"):
            sd = H.make_state_dict(name, kind)
            model = H.build_variant(ref_models, name)
            assert [(k, tuple(v.shape)) for k, v in sd.items()] == [(k, tuple(v.shape)) for k, v in model.state_dict().items()], name
            model.load_state_dict(sd, strict=True)
            model.eval()
            for k, x in H.variant_inputs().items():
                with torch.no_grad():
                    r = model(torch.from_numpy(x))
                o = O.forward(sd, torch.from_numpy(x), act=act, pooling=pooling)
                err = float((r - o).abs().max())
                assert err <= 2e-6, f"oracle != reference: {name}/{kind}/{k}: {err}"
                vout[f"{name}/{kind}/{k}"] = r.numpy()
            print(name, kind, "ok")
    np.savez_compressed(os.path.join(HERE, "probs_variants.npz"), **vout)

    # ---- per-stage trace (kernel bring-up aid): reference module pieces, xxxs/trained, 2 noise clips
    sd = H.make_state_dict("uit_xxxs", "trained")
    model = ref_models.uit_xxxs(outputdim=537, target_length=102); model.load_state_dict(sd); model.eval()
    x = torch.from_numpy(inputs["noise"][:2])
    with torch.no_grad():
        db = model.front_end(x)
        bn = model.init_bn(db.unsqueeze(1))
        tok = model.patch_embed(bn)
        tok = tok + model.time_pos_embed[:, :, :, :tok.shape[-1]] + model.freq_pos_embed
        tok = tok.flatten(2).transpose(1, 2)
        acts, h = [], tok
        for blk in model.blocks:
            h = blk(h); acts.append(h)
        feat = model.norm(h)
        probs = model.forward_head(feat)
    tr = O.forward_trace(sd, x)
    for a, b in ((tr["db"], db), (tr["bn"], bn.squeeze(1)), (tr["tokens"], tok), (tr["blocks"], torch.stack(acts)),
                 (tr["features"], feat), (tr["probs"], probs)):
        assert float((a - b).abs().max()) <= 2e-5, float((a - b).abs().max())
    np.savez_compressed(os.path.join(HERE, "trace_xxxs.npz"), db=db.numpy(), bn=bn.squeeze(1).numpy(), tokens=tok.numpy(),
                        blocks=torch.stack(acts).numpy(), features=feat.numpy(), probs=probs.numpy())
    with open(os.path.join(HERE, "state_dict_layout.txt"), "w") as f:
        for a in H.ARCHS:
            for k, s, d in ref_keys[a]:
                f.write(f"{a}\t{k}\t{s}\t{d}\n")
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
