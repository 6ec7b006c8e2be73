"""Numerics of split-precision operand formats for the gather-GEMM (CPU, numpy; products accumulated exactly so that only the
operand representation shows): 3xTF32 (shipped), bf16x3, fp16x3 with the residual scaled by 2^11 (FSFB_GEMM_F16=1) and unscaled.

    python tools/split_precision_numerics.py

Error is max |result - float64 reference| / max |reference| per case.  Finding: the scaled fp16 split keeps
the 22 mantissa bits of the tf32 split (~8e-8) at twice the tensor rate; bf16x3 loses 6 bits (~5e-6); the unscaled fp16 residual
falls into fp16's subnormal range for small activations (2e-5 at |a| ~ 5e-3)."""
import numpy as np

rng = np.random.default_rng(0)


def tf32_trunc(x):
    return (x.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def tf32_rna(x):
    u = x.view(np.uint32).astype(np.uint64) + 0x1000
    return (u.astype(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def bf16_rn(x):
    u = x.view(np.uint32).astype(np.uint64) + 0x8000
    return (u.astype(np.uint32) & np.uint32(0xFFFF0000)).view(np.float32)


def mm(x, y):
    return x.astype(np.float64) @ y.astype(np.float64)


def run(K, scale_a=1.0, wide=False):
    M, N = 256, 128
    a = np.maximum((rng.standard_normal((M, K)) * scale_a).astype(np.float32), 0)
    if wide:
        a = a * np.exp(rng.uniform(-8, 4, (M, K))).astype(np.float32)
    b = (rng.standard_normal((K, N)) / np.sqrt(K)).astype(np.float32)
    ref = mm(a, b)
    out = {}
    ah, bh = tf32_trunc(a), tf32_rna(b)                     # the kernel truncates A (one AND), the prepack rounds W
    al, bl = tf32_trunc(a - ah), tf32_rna(b - bh)
    out["3xtf32"] = mm(ah, bh) + mm(ah, bl) + mm(al, bh)
    ah, bh = bf16_rn(a), bf16_rn(b)
    al, bl = bf16_rn(a - ah), bf16_rn(b - bh)
    out["bf16x3"] = mm(ah, bh) + mm(ah, bl) + mm(al, bh)
    S = np.float32(2048)
    ah, bh = a.astype(np.float16).astype(np.float32), b.astype(np.float16).astype(np.float32)
    al, bl = ((a - ah) * S).astype(np.float16).astype(np.float32), ((b - bh) * S).astype(np.float16).astype(np.float32)
    out["fp16x3 scaled"] = mm(ah, bh) + (mm(ah, bl) + mm(al, bh)) / 2048
    al, bl = (a - ah).astype(np.float16).astype(np.float32), (b - bh).astype(np.float16).astype(np.float32)
    out["fp16x3 unscaled"] = mm(ah, bh) + mm(ah, bl) + mm(al, bh)
    out["fp32 matmul"] = (a @ b).astype(np.float64)
    sc = np.abs(ref).max()
    print(f"K={K:6d} max|a|={np.abs(a).max():9.3g}  " + "  ".join(f"{k} {np.abs(v - ref).max() / sc:.1e}" for k, v in out.items()))


if __name__ == "__main__":
    run(128)
    run(3456)
    run(13824)
    run(3456, wide=True)
    run(3456, scale_a=1e-3)
    run(3456, scale_a=300.0)
