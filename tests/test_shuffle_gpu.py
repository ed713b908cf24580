"""K4 (byte / bit shuffle on the device) against the numpy restatement and through round trips."""
import zlib

import numpy as np
import pytest
import torch

from oracle import bitround_oracle as bro
from oracle import shuffle_oracle as so
from xbitinfo_b200 import _native as nat
from xbitinfo_b200 import shuffle as xs
from xbitinfo_b200.device import bitround_

from conftest import smooth_field

pytestmark = pytest.mark.gpu

DTYPES = ["float16", "float32", "float64", "int16", "int32", "int64"]
# tile multiples, ragged ends (multiples of 8 / 16 only), tiny, and sizes that defeat the vector paths
SIZES = [8, 16, 64, 1000, 1024, 2048, 4096 + 64, 1 << 16, (1 << 16) + 8, 100_000, 262_144 + 1032]


def _data(dtype, n, seed=0):
    rng = np.random.default_rng(seed)
    if np.dtype(dtype).kind == "f":
        return smooth_field(rng, (n,), dtype, axis=0, scale=0.3, base=3.0)
    info = np.iinfo(dtype)
    return rng.integers(info.min, info.max, size=n, dtype=dtype, endpoint=True)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n", SIZES)
def test_byte_shuffle_equals_oracle_and_round_trips(dtype, n):
    x = _data(dtype, n)
    t = torch.from_numpy(x).cuda()
    before = nat.kernel_launches()
    s = xs.byte_shuffle(t)
    assert nat.kernel_launches() > before
    assert np.array_equal(s.cpu().numpy(), so.byte_shuffle(x))
    back = xs.byte_unshuffle(s, t.dtype)
    assert torch.equal(back.view(torch.uint8), t.view(torch.uint8))


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n", SIZES)
def test_bit_shuffle_equals_oracle_and_round_trips(dtype, n):
    x = _data(dtype, n, seed=1)
    t = torch.from_numpy(x).cuda()
    s = xs.bit_shuffle(t)
    assert np.array_equal(s.cpu().numpy(), so.bit_shuffle(x))
    back = xs.bit_unshuffle(s, t.dtype)
    assert torch.equal(back.view(torch.uint8), t.view(torch.uint8))


def test_unaligned_views_and_odd_counts():
    x = _data("float32", 5000 + 3)
    t = torch.from_numpy(x).cuda()
    v = t[3:]  # 12 bytes into the allocation: not 16-byte aligned; 5000 elements
    assert np.array_equal(xs.byte_shuffle(v).cpu().numpy(), so.byte_shuffle(x[3:]))
    assert np.array_equal(xs.bit_shuffle(v).cpu().numpy(), so.bit_shuffle(x[3:]))
    odd = t[:1001]  # byte shuffle takes any count, bit shuffle needs a multiple of 8
    assert np.array_equal(xs.byte_shuffle(odd).cpu().numpy(), so.byte_shuffle(x[:1001]))
    with pytest.raises(ValueError):
        xs.bit_shuffle(odd)


def test_argument_errors():
    t = torch.zeros(64, dtype=torch.float32, device="cuda")
    with pytest.raises(ValueError):  # overlap
        nat.check(nat.lib().xbi_shuffle(t.data_ptr(), t.data_ptr() + 16, 4, 32, nat.SHUFFLE_BYTES, 0, None))
    with pytest.raises(ValueError):
        nat.check(nat.lib().xbi_shuffle(t.data_ptr(), t.data_ptr(), 3, 8, nat.SHUFFLE_BYTES, 0, None))


def test_codecs_follow_the_numcodecs_protocol_and_help_the_entropy_coder():
    """bitround -> shuffle -> zlib: the pipeline of the reference's save path with the two
    array passes on the device.  The byte-shuffled, rounded stream must compress better than
    the plain one, and decode(encode(x)) must give the rounded array back bit for bit."""
    rng = np.random.default_rng(3)
    x = smooth_field(rng, (64, 4096), "float32", axis=1, scale=0.05, base=280.0)
    t = torch.from_numpy(x).cuda()
    bitround_(t, 7)
    rounded = t.cpu().numpy()
    assert np.array_equal(rounded.view("u4"), bro.bitround(x, 7).view("u4"))
    sizes = {"plain": len(zlib.compress(rounded.tobytes(), 6))}
    for codec in (xs.Shuffle(4), xs.BitShuffle(4)):
        enc = codec.encode(rounded)
        assert enc.dtype == np.uint8 and enc.size == rounded.nbytes
        dec = codec.decode(enc)
        assert np.array_equal(dec.view("u4"), rounded.reshape(-1).view("u4"))
        cfg = codec.get_config()
        assert type(codec).from_config(cfg).get_config() == cfg
        sizes[cfg["id"]] = len(zlib.compress(enc.tobytes(), 6))
    assert sizes["shuffle"] < sizes["plain"]
    # after rounding to 7 mantissa bits the 16 lowest bit planes are empty: the bit shuffle
    # turns them into one run of zero bytes at the head of the stream
    planes = xs.BitShuffle(4).encode(rounded).reshape(32, -1)
    assert not planes[:16].any() and planes[16:23].any()


def test_large_round_trip_float32():
    """1 GiB: shuffle then unshuffle gives the input back; the planes have the documented
    content on a sample (plane 3 of the byte shuffle = the top byte of every element)."""
    n = 1 << 28
    g = torch.Generator(device="cuda").manual_seed(9)
    t = torch.empty(n, dtype=torch.float32, device="cuda").normal_(280.0, 5.0, generator=g)
    s = xs.byte_shuffle(t)
    top = (t[:4096].view(torch.int32) >> 24).to(torch.uint8)
    assert torch.equal(s[3 * n: 3 * n + 4096], top)
    assert torch.equal(xs.byte_unshuffle(s, torch.float32), t)
    del s
    b = xs.bit_shuffle(t)
    sign_plane = b[31 * (n // 8): 31 * (n // 8) + 512]  # plane 31 = the sign bit
    assert int(sign_plane.count_nonzero()) == 0  # all values positive
    assert torch.equal(xs.bit_unshuffle(b, torch.float32), t)


@pytest.mark.parametrize("dtype,keeps", [("float16", (0, 3, 9, 10)), ("float32", (0, 7, 22, 23)), ("float64", (0, 11, 51, 52))])
@pytest.mark.parametrize("n", [8, 1000, 2048, 4096 + 64, (1 << 16) + 8, 262_144 + 1032])
@pytest.mark.parametrize("bits", [False, True])
def test_fused_bitround_shuffle_equals_round_then_shuffle(dtype, keeps, n, bits):
    """K3 fused into K4: one pass, same bytes as bitround followed by the shuffle (oracle for both)."""
    x = _data(dtype, n, seed=4)
    x[::97] = -x[::97]
    if n > 100:
        x[5], x[6], x[7] = np.inf, -np.inf, np.nan  # the integer arithmetic treats them like numcodecs does
    t = torch.from_numpy(x).cuda()
    for keep in keeps:
        got = xs.bitround_shuffle(t, keep, bits=bits).cpu().numpy()
        rounded = bro.bitround(x, keep)
        want = so.bit_shuffle(rounded) if bits else so.byte_shuffle(rounded)
        assert np.array_equal(got, want), (dtype, keep, n, bits)
    assert np.array_equal(t.cpu().numpy().view(f"u{x.dtype.itemsize}"), x.view(f"u{x.dtype.itemsize}"))  # input untouched
    with pytest.raises(ValueError):
        xs.bitround_shuffle(t, -1)
    with pytest.raises(ValueError):
        xs.bitround_shuffle(t, keeps[-1] + 1)
