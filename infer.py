#!/usr/bin/env python
"""Command-line tagger, the B200 counterpart of the reference's `inference.py` (inference.py:11-64).

    python infer.py [-m uit_xs | path/to/checkpoint.pt] [-k 3] [--labels merged_class_label_indices.csv] a.wav b.wav ...

Reads 16 kHz mono 16-bit wav files, feeds the whole file (NOT cropped to 1 s: files longer than 102 frames take the
multi-crop branch, like the reference) as int16 PCM straight to the GPU and prints the top-k of the 537 scores.
Indices above 526 are the GSC keywords and are printed as "Keyword: ..." (inference.py:60-61).  Checkpoints are the
reference's own format: {'config': {'model': name, 'model_args': {...}}, 'model': state_dict} (inference.py:42-48) or the
zenodo files in PRETRAINED_CHECKPOINTS (need network).  There is no CPU fallback.
"""
from __future__ import annotations

import argparse
import sys


def main() -> int:
    ap = argparse.ArgumentParser(description="UiT audio tagging + keyword spotting on B200")
    ap.add_argument("wavs", nargs="+")
    ap.add_argument("-m", "--model", default="uit_xs", help="uit_xs / uit_xxs / uit_xxxs (pretrained, needs network) or a checkpoint path")
    ap.add_argument("-k", "--topk", type=int, default=3)
    ap.add_argument("--labels", default=None, help="csv with columns index,mid,display_name (datasets/merged_class_label_indices.csv)")
    ap.add_argument("--random-init", action="store_true", help="skip checkpoint loading (smoke runs without network)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    args = ap.parse_args()

    import numpy as np
    import torch
    from scipy.io import wavfile
    import uit_mobile_b200 as U

    names = None
    if args.labels:
        import pandas as pd
        names = pd.read_csv(args.labels).set_index("index")["display_name"].to_dict()

    if args.model in U.models.PRETRAINED_CHECKPOINTS:
        entry = U.models.PRETRAINED_CHECKPOINTS[args.model]
        if args.random_init:
            torch.manual_seed(0)           # reproducible smoke weights (tests compare the printout with the oracle)
        model = entry["model"](precision=args.precision, **entry["model_kwargs"])
        if not args.random_init:
            sd = torch.hub.load_state_dict_from_url(entry["chkpt"], map_location="cpu")
            model.load_state_dict(sd, strict=True)
    else:
        dump = torch.load(args.model, map_location="cpu")
        cfg = dump["config"]
        # inference.py:42-47: class count from the checkpoint's config (default 537), model_args required
        model = getattr(U.models, cfg["model"])(outputdim=cfg.get("num_classes", 537), precision=args.precision, **cfg["model_args"])
        model.load_state_dict(dump["model"], strict=True)
    model = model.to("cuda:0").eval()

    for path in args.wavs:
        sr, pcm = wavfile.read(path)
        if sr != 16000:
            raise SystemExit(f"{path}: model is trained on 16 kHz audio, got {sr} Hz")
        if pcm.ndim != 1 or pcm.dtype != np.int16:
            raise SystemExit(f"{path}: expected mono 16-bit PCM")
        with torch.no_grad():
            scores = model(torch.from_numpy(pcm.copy()).unsqueeze(0).to("cuda:0")).squeeze(0).cpu()
        print(f"===== {path} =====")
        for p, i in zip(*scores.topk(args.topk)):
            i = int(i)
            label = names[i] if names else f"class {i}"
            print(f"{'Keyword: ' if i > 526 else ''}{label:<30} {float(p):.4f}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
