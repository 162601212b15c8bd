"""The error bounds K2's certificate relies on (batched.cu batch_delta_kernel), checked on the host by
emulating the operand formats of tcgen05.mma.kind::tf32 with numpy:
  * the tensor core reads the upper 19 bits of an fp32 operand (truncation),
  * queries are pre-rounded to tf32 (round to nearest, ties away: rna_tf32),
  * single pass:  S1 = trunc19(V) . rna(Q)                         bound 2^-9  max(1, dim/1024) |q||v|
  * 3xTF32:       S3 = Vh.Ql + Vh.Qh + Vl.Qh  (Vh = trunc19(V), Vl = rna(V - Vh), Qh = rna(Q), Ql = rna(Q - Qh))
                                                                    bound 2^-15 max(1, dim/640) |q||v|
Products and sums are taken in float64 here; the accumulator's own fp32 truncation (at most dim/8 additions of
2^-23 relative each) is accounted for separately and must fit into the headroom the test measures.
The bf16 rung (kind::f16 on operands this library rounds itself, nearest even at 8 bits) is derived the same way at the end:
                                                                    bound 1.01 * 2^-7 |q||v| + accumulation.
No GPU and no oracle involved: this is arithmetic about the formats only."""
import numpy as np
import pytest


def trunc19(x):
    return (x.astype(np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def rna_tf32(x):
    b = x.astype(np.float32).view(np.uint32).astype(np.uint64)
    return ((b + 0x1000) & 0xFFFFE000).astype(np.uint32).view(np.float32)


def emulate(v, q):
    v, q = v.astype(np.float32), q.astype(np.float32)
    vh, qh = trunc19(v), rna_tf32(q)
    vl, ql = rna_tf32(v - vh), rna_tf32(q - qh)
    # the hardware truncates vl / ql / qh again to 19 bits: they already are tf32 values, so nothing changes
    assert np.array_equal(trunc19(vl), vl) and np.array_equal(trunc19(qh), qh) and np.array_equal(trunc19(ql), ql)
    d = lambda a, b: float(np.dot(a.astype(np.float64), b.astype(np.float64)))
    exact = d(v, q)
    return exact, d(vh, qh), d(vh, ql) + d(vh, qh) + d(vl, qh)


def adversarial(dim, rng, kind):
    """Vectors whose low 13 mantissa bits maximise the format error, aligned so that the errors add up."""
    if kind == "aligned_worst":
        # v just below the next 19-bit value (truncation loses ~2^-10), q just below a rounding midpoint (loses ~2^-11)
        mant_v = rng.integers(0, 1 << 10, dim).astype(np.uint32) << 13 | np.uint32(0x1FFF)
        mant_q = rng.integers(0, 1 << 10, dim).astype(np.uint32) << 13 | np.uint32(0x0FFF)
        exp = np.uint32(127) << 23
        return (exp | mant_v).view(np.float32), (exp | mant_q).view(np.float32)
    if kind == "aligned_small_mantissa":
        # mantissa 1.0...: the relative truncation error is largest
        v = (np.full(dim, 0x3F800000, np.uint32) | np.uint32(0x1FFF)).view(np.float32)
        q = (np.full(dim, 0x3F800000, np.uint32) | np.uint32(0x0FFF)).view(np.float32)
        return v, q
    if kind == "mixed_exponents":
        v = (rng.standard_normal(dim) * np.exp2(rng.integers(-8, 8, dim))).astype(np.float32)
        return v, np.abs(v).astype(np.float32) * np.sign(v)
    v = rng.standard_normal(dim).astype(np.float32)
    return v, rng.standard_normal(dim).astype(np.float32)


@pytest.mark.parametrize("kind", ["aligned_worst", "aligned_small_mantissa", "mixed_exponents", "random"])
@pytest.mark.parametrize("dim", [8, 33, 640, 768, 1024, 1536, 4096])
def test_selection_error_bounds(kind, dim):
    rng = np.random.default_rng(dim * 7 + len(kind))
    worst1 = worst3 = 0.0
    for _ in range(20):
        v, q = adversarial(dim, rng, kind)
        exact, s1, s3 = emulate(v, q)
        scale = float(np.linalg.norm(v.astype(np.float64)) * np.linalg.norm(q.astype(np.float64)))
        worst1 = max(worst1, abs(s1 - exact) / scale)
        worst3 = max(worst3, abs(s3 - exact) / scale)
    # accumulation in the tensor core: at most dim/8 truncating fp32 additions of partial sums bounded by |q||v|
    acc = (dim / 8 + 1) * 2.0 ** -23
    kappa1 = 2.0 ** -9 * max(1.0, dim / 1024)   # batch_delta_kernel's constants
    kappa3 = 2.0 ** -15 * max(1.0, dim / 640)
    assert worst1 + acc <= kappa1, f"single pass: format error {worst1:.3e} + accumulation {acc:.3e} exceeds {kappa1:.3e}"
    assert worst3 < 2.0 ** -21 and worst3 + 3 * acc <= kappa3, f"3xTF32: format error {worst3:.3e} + accumulation {3 * acc:.3e} exceeds {kappa3:.3e}"


def test_single_pass_bound_is_not_vacuous():
    """The aligned worst case really comes close to the bound (three quarters of it): a tighter constant would be wrong."""
    rng = np.random.default_rng(1)
    v, q = adversarial(768, rng, "aligned_small_mantissa")
    exact, s1, _ = emulate(v, q)
    rel = abs(s1 - exact) / float(np.linalg.norm(v.astype(np.float64)) * np.linalg.norm(q.astype(np.float64)))
    assert 0.7 * 2.0 ** -9 < rel < 2.0 ** -9


# ---- bf16 rung: both operands rounded to nearest even at 8 significant bits (convert_bf16_kernel) ------------------------
def rne_bf16(x):
    b = x.astype(np.float32).view(np.uint32).astype(np.uint64)
    r = b + 0x7FFF + ((b >> 16) & 1)
    return (r & 0xFFFF0000).astype(np.uint32).view(np.float32)


def adversarial_bf16(dim, rng, kind):
    if kind == "aligned_worst":
        # both operands just below a rounding midpoint: each loses almost 2^-8 relative, all errors with the same sign
        mant = rng.integers(0, 1 << 7, dim).astype(np.uint32) << 16 | np.uint32(0x7FFF)
        exp = np.uint32(127) << 23
        return (exp | mant).view(np.float32), (exp | mant).view(np.float32)
    if kind == "aligned_small_mantissa":
        v = (np.full(dim, 0x3F800000, np.uint32) | np.uint32(0x7FFF)).view(np.float32)
        return v, v.copy()
    if kind == "rounds_up":
        v = (np.full(dim, 0x3F800000, np.uint32) | np.uint32(0x8001)).view(np.float32)
        return v, v.copy()
    return adversarial(dim, rng, kind)


@pytest.mark.parametrize("kind", ["aligned_worst", "aligned_small_mantissa", "rounds_up", "mixed_exponents", "random"])
@pytest.mark.parametrize("dim", [8, 33, 640, 768, 1024, 1536, 4096])
def test_bf16_selection_error_bound(kind, dim):
    rng = np.random.default_rng(dim * 11 + len(kind))
    worst = 0.0
    for _ in range(20):
        v, q = adversarial_bf16(dim, rng, kind)
        vb, qb = rne_bf16(v), rne_bf16(q)
        exact = float(np.dot(v.astype(np.float64), q.astype(np.float64)))
        approx = float(np.dot(vb.astype(np.float64), qb.astype(np.float64)))
        scale = float(np.linalg.norm(v.astype(np.float64)) * np.linalg.norm(q.astype(np.float64)))
        worst = max(worst, abs(approx - exact) / scale)
    acc = (dim / 16 + 1) * 2.0 ** -23                           # truncating fp32 additions in the tensor core, K = 16 per step
    kappa = 1.01 * 2.0 ** -7 + (dim / 16 + 1) * 2.0 ** -21      # batch_delta_kernel's constant (4x the accumulation model)
    assert worst <= 2.0 ** -7 + 2.0 ** -16, f"operand rounding error {worst:.3e} exceeds the format bound"
    assert worst + 4 * acc <= kappa


def test_bf16_bound_is_not_vacuous():
    v, q = adversarial_bf16(768, np.random.default_rng(1), "aligned_small_mantissa")
    vb, qb = rne_bf16(v), rne_bf16(q)
    exact = float(np.dot(v.astype(np.float64), q.astype(np.float64)))
    approx = float(np.dot(vb.astype(np.float64), qb.astype(np.float64)))
    rel = abs(approx - exact) / float(np.linalg.norm(v.astype(np.float64)) * np.linalg.norm(q.astype(np.float64)))
    assert 0.95 * 2.0 ** -7 < rel < 2.0 ** -7
