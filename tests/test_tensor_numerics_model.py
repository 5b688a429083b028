"""CPU model of the operand rounding in the tensor-core band DFT (csrc/kernels_tc.cu): the exact float64 product
sum_n A[n] x[n] against (a) the 3xTF32 split Ahi*xhi + Alo*xhi + Ahi*xlo (SYLDET_KERNEL_TENSOR_TF32) and (b) the production
variant for the sample shape: two-term fp16 splits of both operands, a1 = fp16(A), a2 = fp16(A - a1), h1 = fp16(x),
h2 = fp16((x - h1) 2^11), as one K-concatenated pass [a1 | a1 2^-11 | a2] . [h1 ; h2 ; h1].
Sums are taken in float64, so only the operand roundings are modelled (the tensor core adds float32 accumulation on top,
which the GPU parity tests cover). The model pins the validity range of (b): float32-level down to rms ~1e-5, degrading as
~1e-11 / rms below (where the kernel's range guard hands the launch to (a)), while (a) is amplitude-invariant."""
import numpy as np
import pytest


def _tf32_trunc(x):     # what the tensor core does to an fp32 B operand: the low 13 mantissa bits are ignored
    return (np.asarray(x, np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def _tf32_round(x):     # host-side split of the DFT matrix (engine.cu tf32_round): nearest, ties away
    u = np.asarray(x, np.float32).view(np.uint32)
    return ((u + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


def _operands(n_rows=64, n=256, seed=0):
    rng = np.random.default_rng(seed)
    m = np.arange(n)
    win = (0.54 - 0.46 * np.cos(2 * np.pi * m / n)).astype(np.float32).astype(np.float64)
    k = rng.integers(12, 41, size=n_rows)[:, None]
    a = win[None, :] * np.cos(-2 * np.pi * ((k * m[None, :]) % n) / n)          # rows of the windowed DFT matrix (real part)
    x = (rng.standard_normal((n_rows, n)) * 1.3e-2).astype(np.float32)          # the benchmark audio's rms
    return a, x


def _errors(a, x):
    exact = (a * x.astype(np.float64)).sum(axis=1)
    a_hi = _tf32_round(a.astype(np.float32)).astype(np.float64)
    x_hi = _tf32_trunc(x).astype(np.float64)
    x_lo = x.astype(np.float64) - x_hi
    # (a) 3xTF32: Alo rounded to tf32, xlo exact in fp32 (<= 13 significant bits, truncated to tf32 by the tensor core)
    a_lo32 = _tf32_round((a - a_hi).astype(np.float32)).astype(np.float64)
    tf32 = (a_hi * x_hi + a_lo32 * x_hi + a_hi * _tf32_trunc(x_lo.astype(np.float32)).astype(np.float64)).sum(axis=1)
    # (b) two-term fp16 splits (plan_tc put16 / the splitters' convert())
    a1 = a.astype(np.float32).astype(np.float16)
    a1s = (a1.astype(np.float32) * np.float32(2.0 ** -11)).astype(np.float16).astype(np.float64)
    a2 = (a - a1.astype(np.float64)).astype(np.float32).astype(np.float16).astype(np.float64)
    with np.errstate(over="ignore", invalid="ignore"):
        h1 = x.astype(np.float16)
        h2 = ((x - h1.astype(np.float32)) * np.float32(2048.0)).astype(np.float16).astype(np.float64)
        f16 = (a1.astype(np.float64) * h1.astype(np.float64) + a1s * h2 + a2 * h1.astype(np.float64)).sum(axis=1)
    scale = np.sqrt((a * a).sum(axis=1) * (x.astype(np.float64) ** 2).sum(axis=1))   # |A| |x|: what the error terms scale with
    return np.abs(tf32 - exact) / scale, np.abs(f16 - exact) / scale


@pytest.mark.parametrize("exp2", [12, 0, -6, -10])
def test_fp16_correction_pass_is_float32_level_in_range(exp2):
    a, x = _operands()
    e_tf32, e_f16 = _errors(a, (x * np.float32(2.0 ** exp2)).astype(np.float32))
    assert e_tf32.max() < 2e-7 and e_f16.max() < 2e-7, (e_tf32.max(), e_f16.max())   # float32 epsilon is 6e-8


def test_fp16_correction_pass_degrades_below_the_fp16_normal_range_but_tf32_does_not():
    a, x = _operands(seed=1)
    worst = {}
    for exp2 in (-14, -18, -22):
        e_tf32, e_f16 = _errors(a, (x * np.float32(2.0 ** exp2)).astype(np.float32))
        assert e_tf32.max() < 2e-7                        # amplitude-invariant: SYLDET_KERNEL_TENSOR_TF32
        worst[exp2] = e_f16.max()
    assert worst[-14] < 2e-5 and worst[-18] > worst[-14] and worst[-22] > 1e-6, worst   # ~1e-11 / rms


def test_samples_beyond_the_fp16_range_are_not_supported_by_the_fp16_pass():
    a, x = _operands(seed=2)
    x = x.copy()
    x[:, 7] = 1.0e5                                       # > 65504: fp16(x) = inf (documented limit, include/syldet.h)
    e_tf32, e_f16 = _errors(a, x)
    assert e_tf32.max() < 2e-7 and not np.isfinite(e_f16).all()
