"""The bf16 rounding the parity contract of OTTERS_VECTORS_FMT_BF16 stores rests on (oracle.round_bf16): checked against a
scalar restatement of IEEE round-to-nearest-even written with exact rational arithmetic, against every tie / boundary case,
and against the product's own helper (otters_b200.round_to_bf16) and the bench's (bench_workloads.round_bf16_np).  The device
side (cvt.rn.bf16x2.f32 in convert_bf16_kernel) is held to the same values by tests/test_gpu_bf16_store.py."""
import struct
from fractions import Fraction

import numpy as np

from helpers import ob, ora
from bench_workloads import round_bf16_np


def f32_bits(x):
    return struct.unpack("<I", struct.pack("<f", x))[0]


def bits_f32(b):
    return struct.unpack("<f", struct.pack("<I", b & 0xFFFFFFFF))[0]


def scalar_rne(x: float) -> float:
    """Nearest bf16 (8 significant bits) of a finite fp32 value by exact comparison of the two neighbours; ties to even."""
    b = f32_bits(x)
    lo = b & 0xFFFF0000                 # truncation towards zero: the neighbour with the smaller magnitude
    hi = lo + 0x10000                   # the next bf16 away from zero (may be inf)
    vx, vlo = Fraction(x), Fraction(bits_f32(lo))
    fhi = bits_f32(hi)
    if fhi in (float("inf"), float("-inf")):   # distance to "the value inf would have": 2^128
        vhi = Fraction(2) ** 128 * (1 if x > 0 else -1)
    else:
        vhi = Fraction(fhi)
    dlo, dhi = abs(vx - vlo), abs(vhi - vx)
    if dlo < dhi:
        return bits_f32(lo)
    if dhi < dlo:
        return fhi
    return bits_f32(lo) if ((lo >> 16) & 1) == 0 else fhi


def test_oracle_rounding_matches_the_scalar_restatement():
    rng = np.random.default_rng(0)
    x = np.concatenate([
        rng.standard_normal(2000).astype(np.float32),
        (rng.standard_normal(500) * 1e30).astype(np.float32),
        (rng.standard_normal(500) * 1e-30).astype(np.float32),
        np.array([0.0, -0.0, 1.0, -1.0, 1.00390625, 1.01171875, 1.0039062, 1.0039063, 3.3e38, -3.3e38, 3.4e38, 1e-45, 1.1754944e-38,
                  65280.0, 65408.0, 65407.996], np.float32),
        (np.arange(0x3F800000, 0x3F800000 + 0x30000, 0x1000, dtype=np.uint32)).view(np.float32),   # every 1/16 of three bf16 steps
    ])
    got = ora.round_bf16(x)
    want = np.array([scalar_rne(float(v)) for v in x], np.float32)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert (got.view(np.uint32) & 0xFFFF == 0).all()


def test_ties_go_to_the_even_neighbour_and_overflow_goes_to_infinity():
    t = np.array([bits_f32(0x3F808000), bits_f32(0x3F818000), bits_f32(0x3F807FFF), bits_f32(0x3F808001)], np.float32)
    assert [hex(v) for v in ora.round_bf16(t).view(np.uint32)] == ["0x3f800000", "0x3f820000", "0x3f800000", "0x3f810000"]
    big = np.array([bits_f32(0x7F7F7FFF), bits_f32(0x7F7F8000), bits_f32(0xFF7FFFFF)], np.float32)
    assert [hex(v) for v in ora.round_bf16(big).view(np.uint32)] == ["0x7f7f0000", "0x7f800000", "0xff800000"]
    assert np.isnan(ora.round_bf16(np.array([np.nan, -np.nan], np.float32))).all()
    assert np.array_equal(ora.round_bf16(np.array([np.inf, -np.inf], np.float32)), np.array([np.inf, -np.inf], np.float32))


def test_product_and_bench_helpers_agree_with_the_oracle():
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.standard_normal((300, 17)).astype(np.float32).ravel(),
                        np.arange(0x3F800000, 0x3F800000 + 0x40000, 0x800, dtype=np.uint32).view(np.float32),
                        np.array([0.0, -0.0, np.inf, -np.inf, 3.4e38, -3.4e38, 1e-45], np.float32)])
    want = ora.round_bf16(x).view(np.uint32)
    assert np.array_equal(ob.round_to_bf16(x).view(np.uint32), want)
    assert np.array_equal(round_bf16_np(x).view(np.uint32), want)
    assert ob.round_to_bf16(x.reshape(-1, 1)).shape == (len(x), 1)
    assert np.isnan(ob.round_to_bf16(np.array([np.nan], np.float32))).all()


def test_rounding_is_idempotent_and_widening_is_exact():
    x = np.random.default_rng(2).standard_normal(1000).astype(np.float32)
    r = ora.round_bf16(x)
    assert np.array_equal(ora.round_bf16(r).view(np.uint32), r.view(np.uint32))
    # what the kernels do with a stored bf16: append sixteen zero bits
    stored = (r.view(np.uint32) >> 16).astype(np.uint16)
    assert np.array_equal((stored.astype(np.uint32) << 16).view(np.float32), r)
    assert (np.abs(r - x) <= np.abs(x) * 2.0 ** -8).all()
