"""What else could compute the log-mel front-end of the bench (4096 x 1 s clips = 413 696 frames) on this B200?  CUDA-event times of
  (a) torch.stft alone (cuFFT, batched 512-point R2C) and the whole reference front-end as eager PyTorch ops,
  (b) the DFT as ONE dense GEMM on the tensor cores through cuBLAS: frames[413696, 512] x DFT[512, 512] (257 cos + 255 sin columns),
      bf16 hi/lo operands = 3 GEMMs (2^-16 relative; fp32-grade hi/mid/lo needs 6), the frames ALREADY framed, windowed and split,
  (c) the same in TF32 (torch.backends.cuda.matmul.allow_tf32, 3 GEMMs for hi/lo),
  (d) this repo's fused kernel (framing, window, FFT, power, mel, dB, max/min words - everything).
(b)/(c) are contraction-only lower bounds of the "DFT-matrix contraction on tcgen05" that BASELINE.json's north_star sketches; the
hand-written two-stage (16 x 32) variant is bounded in scripts/microbench/tc_micro.cu.  Output kept in profiles/."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import uit_mobile_b200 as U

dev = "cuda:0"
torch.manual_seed(0)
model = U.models.uit_xs(outputdim=537, target_length=102).to(dev).eval()
B, L = 4096, 16000
x = (0.1 * torch.randn(B, L, device=dev)).clamp_(-1, 1)
win = model.front_end[0].spectrogram.window
fb = model.front_end[0].mel_scale.fb


def ms(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


with torch.no_grad():
    t_stft = ms(lambda: torch.stft(x, 512, 160, 512, win, center=True, pad_mode="reflect", return_complex=True))

    def eager():
        s = torch.stft(x, 512, 160, 512, win, center=True, pad_mode="reflect", return_complex=True).abs().pow(2.0)
        mel = torch.matmul(s.transpose(-1, -2), fb).transpose(-1, -2)
        db = 10.0 * torch.log10(torch.clamp(mel, min=1e-10))
        return torch.max(db, db.amax() - 120.0)
    t_eager = ms(eager)
    F = B * 101
    frames_hi = torch.randn(F, 512, device=dev, dtype=torch.bfloat16)
    frames_lo = torch.randn(F, 512, device=dev, dtype=torch.bfloat16)
    dft_hi = torch.randn(512, 512, device=dev, dtype=torch.bfloat16)
    dft_lo = torch.randn(512, 512, device=dev, dtype=torch.bfloat16)
    acc = torch.empty(F, 512, device=dev, dtype=torch.bfloat16)

    def bf16_3():
        torch.mm(frames_hi, dft_hi, out=acc); torch.mm(frames_lo, dft_hi, out=acc); torch.mm(frames_hi, dft_lo, out=acc)
    t_bf16 = ms(bf16_3)
    f32 = torch.randn(F, 512, device=dev)
    d32 = torch.randn(512, 512, device=dev)
    o32 = torch.empty(F, 512, device=dev)
    torch.backends.cuda.matmul.allow_tf32 = True

    def tf32_3():
        torch.mm(f32, d32, out=o32); torch.mm(f32, d32, out=o32); torch.mm(f32, d32, out=o32)
    t_tf32 = ms(tf32_3)
    torch.backends.cuda.matmul.allow_tf32 = False
    dbo = torch.empty(B, 64, 101, device=dev)
    t_ours = ms(lambda: model.front_end.logmel_unclamped(x, out=dbo), iters=20)

gb = B * (4 * L + 4 * 64 * 101) / 1e6
print(f"front-end alternatives, {B} x 1 s clips ({F} frames), one B200:")
print(f"  (a) torch.stft alone (cuFFT R2C, no power / mel / dB)            {t_stft:7.3f} ms")
print(f"      reference front-end as eager PyTorch (stft, pow, mel matmul, log10, clamp) {t_eager:7.3f} ms  = {gb / t_eager:6.0f} GB/s algorithmic")
print(f"  (b) dense DFT as cuBLAS GEMMs, bf16 hi/lo (3 x [{F} x 512] x [512 x 512]), contraction only   {t_bf16:7.3f} ms  ({3 * 2 * F * 512 * 512 / t_bf16 / 1e9:6.0f} TFLOP/s)")
print(f"  (c) dense DFT as cuBLAS GEMMs, TF32 hi/lo (3 GEMMs, fp32 operands), contraction only          {t_tf32:7.3f} ms  ({3 * 2 * F * 512 * 512 / t_tf32 / 1e9:6.0f} TFLOP/s)")
print(f"  (d) uit_mobile_b200 logmel_kernel (everything, fp32-grade)        {t_ours:7.3f} ms  = {gb / t_ours:6.0f} GB/s algorithmic")
