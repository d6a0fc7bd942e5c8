"""Float64 first-principles restatement of the log-mel front-end.  TEST INFRASTRUCTURE ONLY.

Independent of ``torch.stft``: explicit reflect padding, framing, windowing and a 512-point real DFT in
float64 (numpy).  Used by the tests to arbitrate fp32 rounding noise between the reference's pocketfft path,
the fp32 oracle (``uit_oracle.py``) and the CUDA kernel: all three must sit within a few fp32 ulps (in the
power domain) of this one.  Follows TA:functional/functional.py:123-144 (stft, power), TA:transforms/
_transforms.py:407-419 (mel matmul) and TA:functional/functional.py:390-399 (dB + top-dB).
"""
from __future__ import annotations

import numpy as np

N_FFT, HOP, N_MELS, TOP_DB, AMIN = 512, 160, 64, 120.0, 1e-10


def reflect_index(i: np.ndarray, L: int) -> np.ndarray:
    """Index map of ``F.pad(mode='reflect')`` (no edge repeat): -1 -> 1, L -> L-2."""
    i = np.abs(i)
    return np.where(i >= L, 2 * (L - 1) - i, i)


def frames(wav: np.ndarray) -> np.ndarray:
    """[B, L] -> [B, T, 512] with frame t = samples [160t-256, 160t+256) under reflect padding."""
    B, L = wav.shape
    if L <= N_FFT // 2:
        raise ValueError("reflect padding needs L > 256")
    T = 1 + L // HOP
    idx = (np.arange(T)[:, None] * HOP - N_FFT // 2) + np.arange(N_FFT)[None, :]
    return wav[:, reflect_index(idx, L)]


def mel_power(wav: np.ndarray, window: np.ndarray, fb: np.ndarray) -> np.ndarray:
    """[B, L] -> mel power [B, 64, T] in float64."""
    fr = frames(wav.astype(np.float64)) * window.astype(np.float64)[None, None, :]
    spec = np.fft.rfft(fr, axis=-1)                       # float64 pocketfft
    power = spec.real ** 2 + spec.imag ** 2               # [B, T, 257]
    return np.einsum("btk,km->bmt", power, fb.astype(np.float64))


def logmel(wav: np.ndarray, window: np.ndarray, fb: np.ndarray, cutoff_max_db=None) -> np.ndarray:
    mel = mel_power(wav, window, fb)
    db = 10.0 * np.log10(np.maximum(mel, AMIN))
    gmax = db.max() if cutoff_max_db is None else float(cutoff_max_db)
    return np.maximum(db, gmax - TOP_DB)
