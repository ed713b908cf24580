"""The CPU oracle against the reference's golden vectors and known answers (no GPU)."""
import numpy as np
import pytest

from oracle import bitinfo_oracle as bo
from oracle import bitround_oracle as bro
from oracle import keepbits_oracle as ko

RH2_INFO = [
    0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.11799129e-01, 8.19114977e-01, 4.41578500e-01, 3.25470303e-01,
    4.35195738e-01, 2.81462993e-01, 2.10719742e-01, 1.46638224e-01, 9.24534031e-02, 4.41090879e-02,
    1.13842504e-02, 8.20088050e-04, 2.62239097e-06, 7.11284508e-07, 1.18183485e-06, 9.49338973e-09,
    1.80859255e-07, 7.72662891e-07, 1.37865391e-05, 2.11117224e-06, 2.01353088e-07, 3.20755770e-02,
    6.06721012e-03, 2.25987148e-04, 1.71530452e-06, 5.13067595e-03,
]  # reference: tests/test_get_keepbits.py:83-116


def test_exponent_doctests():
    # reference: xbitinfo/_py_bitinfo.py:12-15, 28-31
    assert bo.exponent_bias("f4") == 127
    assert bo.exponent_bias("f8") == 1023
    assert np.binary_repr(bo.exponent_mask(np.float32), width=32) == "01111111100000000000000000000000"
    assert np.binary_repr(bo.exponent_mask(np.float16), width=16) == "0111110000000000"


def test_signed_exponent_doctests():
    # reference: xbitinfo/_py_bitinfo.py:57-66
    a = np.array(0.03125, dtype="float32")
    assert np.binary_repr(int(bo.signed_exponent(a)), width=32) == "01000010100000000000000000000000"
    a = np.array(0.03125, dtype="float64")
    assert np.binary_repr(int(bo.signed_exponent(a)), width=64) == "0100000001010000" + "0" * 48


@pytest.mark.parametrize("dt", ["float16", "float32", "float64"])
def test_signed_exponent_matches_reference_output(golden, dt):
    x = golden[f"sexp_{dt}_in"].view(dt)
    ref = golden[f"sexp_{dt}_out"]
    assert np.array_equal(bo.signed_exponent(x), ref)
    assert np.array_equal(bo.signed_exponent_closed_form(x), ref)


def test_signed_exponent_is_a_bijection_on_float16():
    x = np.arange(1 << 16, dtype="uint16").view("float16")
    assert len(np.unique(bo.signed_exponent(x))) == 1 << 16


def test_counts_and_information_match_reference_output(golden):
    for name in golden["case_names"]:
        x = golden[f"case_{name}_x"]
        axis = int(golden[f"case_{name}_axis"])
        for route in ("unpackbits", "planes"):
            counts, npairs, mi = bo.bitinformation(x, axis, route=route)
            assert np.array_equal(counts, golden[f"case_{name}_counts"]), (name, route)
            ref = golden[f"case_{name}_mi"]
            assert np.array_equal(mi.view("u8"), ref.view("u8")) or (np.isnan(mi).all() and np.isnan(ref).all())
        scalar = np.array([bo.mutual_information_scalar(*counts[k].reshape(-1).tolist()) for k in range(len(counts))])
        both_nan = np.isnan(scalar) & np.isnan(ref)
        assert ((scalar.view("u8") == ref.view("u8")) | both_nan).all(), name


def test_mutual_information_rounding_pins(golden):
    a, b, nm = golden["mi_pairs_a"], golden["mi_pairs_b"], golden["mi_pairs_n_and_mi"]
    for i in range(len(a)):
        n = int(nm[i, 0])
        counts = bo.paircount_planes(a[i, :n].astype("u1"), b[i, :n].astype("u1"))
        mi = bo.mutual_information_from_counts(counts, n)
        assert np.array_equal(mi.view("u8"), nm[i, 1:].view("u8"))


def test_count_table_properties():
    rng = np.random.default_rng(1)
    x = rng.integers(0, 2**32, size=(6, 50, 70), dtype="uint32")
    for axis in range(3):
        a, b = bo.adjacent_views(x, axis)
        c1, c2 = bo.paircount_unpackbits(a, b), bo.paircount_planes(a, b)
        assert np.array_equal(c1, c2)
        assert (c1.sum(axis=(1, 2)) == a.size).all()


def test_masked_pairs_spec():
    x = np.array([[1.0, np.nan, 2.0, 3.0, 4.0], [5.0, 6.0, -999.0, 7.0, 8.0]], dtype="float32")
    _, n_nan, _ = bo.bitinformation(x, 1, masked_value=float("nan"))
    assert n_nan == 2 + 4
    _, n_val, _ = bo.bitinformation(x, 1, masked_value=-999.0)
    assert n_val == 4 + 2
    _, n_none, _ = bo.bitinformation(x, 1, masked_value=None)
    assert n_none == 8


def test_free_entropy_known_answers():
    # SURVEY.md Appendix A.6 (scipy probe of the BitInformation.jl formulas)
    assert bo.binom_confidence_z(0.99) == pytest.approx(2.5758293035489004, rel=1e-14)
    assert bo.binom_confidence_z(0.5) == pytest.approx(0.6744897501960817, rel=1e-14)
    assert bo.free_entropy_threshold(692634, 0.99) == pytest.approx(6.9099608923650635e-06, rel=1e-9)
    assert bo.free_entropy_threshold(9088666440, 0.99) == pytest.approx(5.265972102819205e-10, rel=1e-6)


def test_keepbits_golden_vector():
    # reference: tests/test_get_keepbits.py:117-138
    assert ko.keepbits_from_info(RH2_INFO, "float32", 0.99) == 19
    assert ko.keepbits_from_info(RH2_INFO, "float32", 0.99, "Gradient", threshold=0.7, tolerance=0.001) == 7
    assert ko.keepbits_from_info(RH2_INFO, "float32", 1.0) == 23
    assert ko.keepbits_from_info(np.zeros(64), "float64", 1.0) == 52


def test_bit_partitioning_table():
    # reference: xbitinfo/xbitinfo.py:46-67
    assert ko.bit_partitioning("float16") == (16, 1, 5, 10)
    assert ko.bit_partitioning("float32") == (32, 1, 8, 23)
    assert ko.bit_partitioning("float64") == (64, 1, 11, 52)
    assert ko.bit_partitioning("int16") == (16, 1, 0, 15)
    assert ko.bit_partitioning("uint32") == (32, 0, 0, 32)


def test_bitround_against_independent_rounding():
    allf16 = np.arange(1 << 16, dtype="uint16").view("float16")
    fin = allf16[np.isfinite(allf16)]
    for k in range(1, 10):
        assert np.array_equal(bro.bitround(fin, k).view("u2"), bro.bitround_float_rne(fin, k).view("u2"))
    rng = np.random.default_rng(0)
    x = (rng.standard_normal(50000) * np.exp(rng.uniform(-80, 80, 50000))).astype("float32")
    x = x[np.isfinite(x)]
    for k in (1, 7, 15, 22):
        assert np.array_equal(bro.bitround(x, k).view("u4"), bro.bitround_float_rne(x, k).view("u4"))


def test_bitround_contract():
    x = np.linspace(-3, 3, 101).astype("float32")
    assert np.array_equal(bro.bitround(x, 23), x)
    with pytest.raises(ValueError):
        bro.bitround(x, 24)
    with pytest.raises(ValueError):
        bro.bitround(x, -1)
    with pytest.raises(TypeError):
        bro.bitround(np.arange(4), 3)
    y = bro.bitround(x, 7)
    assert not np.shares_memory(x, y)
    np.testing.assert_allclose(y, x, rtol=2 ** -8, atol=0)
    assert (y.view("u4") & ((1 << 16) - 1) == 0).all()


def test_shuffle_oracle_layouts_and_round_trips():
    """The byte / bit shuffle restatements: documented layouts on hand-made vectors,
    inverses, and the relation between the two (a bit plane of the bit shuffle is the packed
    bit of the matching byte plane of the byte shuffle)."""
    from oracle import shuffle_oracle as so

    x = np.array([0x04030201, 0x08070605, 0x0C0B0A09, 0x100F0E0D, 0, 0, 0, 0x80000000], dtype="<u4")
    b = so.byte_shuffle(x)
    assert b.size == 32 and list(b[:4]) == [1, 5, 9, 13] and list(b[24:28]) == [4, 8, 12, 16] and b[31] == 0x80
    s = so.bit_shuffle(x)
    assert s.size == 32 and s[0] == 0b00001111 and s[31] == 0b10000000  # bit 0 of elements 0..3; sign of element 7
    rng = np.random.default_rng(0)
    for dt in ("float16", "float32", "float64", "int16", "int64"):
        y = rng.integers(0, 255, size=8 * 37 * np.dtype(dt).itemsize, dtype="u1").view(dt)
        assert np.array_equal(so.byte_unshuffle(so.byte_shuffle(y), dt).view("u1"), y.view("u1"))
        assert np.array_equal(so.bit_unshuffle(so.bit_shuffle(y), dt).view("u1"), y.view("u1"))
        planes = so.byte_shuffle(y).reshape(y.dtype.itemsize, -1)
        bits = so.bit_shuffle(y).reshape(8 * y.dtype.itemsize, -1)
        for j in range(y.dtype.itemsize):
            for k in range(8):
                assert np.array_equal(bits[8 * j + k], np.packbits((planes[j] >> k) & 1, bitorder="little"))


# --------------------------------------------------------------------------------------
# oracle/fast_count.c: the multi-threaded C restatement used for full-size bit-exact checks
# --------------------------------------------------------------------------------------
def test_fast_count_reproduces_the_golden_cases(golden):
    from oracle import fast_count as fc

    for name in golden["case_names"]:
        x = golden[f"case_{name}_x"]
        axis = int(golden[f"case_{name}_axis"])
        assert np.array_equal(fc.count_pairs(x, axis), golden[f"case_{name}_counts"]), name


@pytest.mark.parametrize("dtype", ["float32", "float64", "float16", "int8", "uint8", "int16", "uint32", "int64"])
def test_fast_count_equals_the_numpy_oracle(dtype):
    from oracle import fast_count as fc

    rng = np.random.default_rng(23)
    for shape, axis in (((7, 33, 20), 0), ((7, 33, 20), 1), ((7, 33, 20), 2), ((1000,), 0), ((130,), 0), ((5, 1), 1),
                        ((3, 2), 1), ((2, 64, 3), 1), ((0, 4), 1)):
        if np.dtype(dtype).kind == "f":
            x = (3.0 + np.cumsum(rng.standard_normal(shape) * 0.3, axis=axis)).astype(dtype)
            x[rng.random(shape) < 0.1] = np.nan
            if x.size:
                x.flat[0] = np.inf
            masks = [None, float("nan"), -999.0, float(x.flat[1]) if x.size > 1 else 1.0]
        else:
            info = np.iinfo(dtype)
            x = rng.integers(max(info.min, -40), min(info.max, 40), size=shape, dtype=dtype, endpoint=True)
            masks = [None, int(x.flat[0]) if x.size else 0]
        for mv in masks:
            want = bo.bitinformation(x, axis, masked_value=mv)[0]
            for threads in (0, 1, 3):
                assert np.array_equal(fc.count_pairs(x, axis, mv, threads=threads), want), (shape, axis, mv, threads)
