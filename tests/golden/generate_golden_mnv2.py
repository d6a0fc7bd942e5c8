"""Golden vectors of the reference's MobileNetV2 (build container only; /root/reference is not on the GPU box).

    python tests/golden/generate_golden_mnv2.py

Imports /root/reference/models (same timm shim as generate_golden.py), loads the seeded state_dicts of ``tests/helpers.py``
with ``strict=True`` (pins the state_dict key / shape contract), runs the reference's own eval forward on CPU, asserts that
``oracle/mobilenetv2_oracle.py`` reproduces it (<= 2e-6 on the scores, per-block trace <= 1e-4 relative) and writes
``mobilenetv2.npz`` + ``state_dict_layout_mnv2.txt``.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, HERE)


def main():
    from generate_golden import import_reference
    from tests import helpers as H
    from oracle import mobilenetv2_oracle as M
    ref_models = import_reference()
    torch.manual_seed(0)
    torch.set_num_threads(8)
    inputs = {"noise": H.noise_clips(6), "adversarial": H.adversarial_batch(), "short2400": H.noise_clips(3, 2400, seed=11),
              "len16160": H.noise_clips(2, 16160, seed=14), "long10s": H.noise_clips(2, 160000, seed=13)}
    out = {}
    layout = None
    for kind in ("init", "trained"):
        sd = H.make_mnv2_state_dict(kind)
        model = ref_models.MobileNetV2(outputdim=H.OUTPUTDIM)
        ref_layout = [(k, tuple(v.shape), str(v.dtype)) for k, v in model.state_dict().items()]
        assert [(k, tuple(v.shape), str(v.dtype)) for k, v in sd.items()] == ref_layout, "state_dict layout drifted"
        layout = ref_layout
        model.load_state_dict(sd, strict=True)
        model.eval()
        for k, x in inputs.items():
            with torch.no_grad():
                r = model(torch.from_numpy(x))
            o = M.forward(sd, torch.from_numpy(x))
            err = float((r - o).abs().max())
            assert err <= 2e-6, f"oracle != reference: {kind}/{k}: {err}"
            out[f"{kind}/{k}"] = r.numpy()
            print(kind, k, "ok; scores", float(r.min()), float(r.max()), "classes > 0.1:", int((r > 0.1).sum(1).float().mean()))
        # per-block trace of the reference's own modules (kernel bring-up aid), 2 noise clips
        if kind == "trained":
            x = torch.from_numpy(inputs["noise"][:2])
            with torch.no_grad():
                h = model.front_end(x).unsqueeze(1)
                acts = []
                for i, layer in enumerate(model.features[:-1]):
                    h = layer(h)
                    if 1 <= i <= 17:
                        acts.append(h)
            tr = []
            hf = M.features(M.O.logmel(x, sd["front_end.0.spectrogram.window"], sd["front_end.0.mel_scale.fb"]).unsqueeze(1), sd, tr)
            for a, b in zip(tr, acts):
                assert float((a - b).abs().max()) <= 1e-4 * max(1.0, float(b.abs().max()))
            assert float((hf - h).abs().max()) <= 1e-4 * max(1.0, float(h.abs().max()))
            out["trace/block_absmax"] = np.array([float(a.abs().max()) for a in acts], np.float32)
            out["trace/last_conv"] = h.numpy()
    np.savez_compressed(os.path.join(HERE, "mobilenetv2.npz"), **out)
    with open(os.path.join(HERE, "state_dict_layout_mnv2.txt"), "w") as f:
        for k, s, d in layout:
            f.write(f"MobileNetV2\t{k}\t{s}\t{d}\n")
    print("written")


if __name__ == "__main__":
    main()
