"""CPU check (numpy, exact integers) of the arithmetic behind `csrc/contract_i8.cu`: the biased-byte digit extraction, the
exactness of the six-accumulator recombination, the INT32 bounds, and the size of the dropped digit products.  No GPU, no
library call: this pins the scheme itself, the GPU tests pin its implementation."""
import numpy as np

ND = 6
BIAS = 0x808080808080


def digits_via_bias(v):
    """What to_w48 + plane_word do: w = v + 0x808080808080, digit k = (byte k of w) XOR 0x80 as a signed byte."""
    w = (v.astype(np.int64) + BIAS).astype(np.uint64)
    d = np.empty(v.shape + (ND,), dtype=np.int64)
    for k in range(ND):
        b = ((w >> np.uint64(8 * k)) & np.uint64(0xFF)).astype(np.int64) ^ 0x80
        d[..., k] = np.where(b >= 128, b - 256, b)
    return d  # least significant first


def test_biased_bytes_are_the_balanced_base256_digits():
    rng = np.random.default_rng(0)
    v = rng.integers(-(1 << 46), (1 << 46) + 1, size=200_000, dtype=np.int64)
    v[:4] = [0, (1 << 46), -(1 << 46), -1]
    d = digits_via_bias(v)
    assert d.min() >= -128 and d.max() <= 127
    assert np.abs(d[:, ND - 1]).max() <= 65          # top digit: |v| <= 2^46 leaves one bit of head-room
    rec = sum(d[:, k] << (8 * k) for k in range(ND))
    assert np.array_equal(rec, v)


def test_magic_constant_rounding_matches_rint():
    """fma(x, scale, 1.5 * 2^52) puts rint(x * scale) into the low mantissa bits for |x * scale| < 2^51."""
    rng = np.random.default_rng(1)
    x = rng.standard_normal(100_000) * 2.0 ** rng.integers(-30, 1, 100_000)
    e = np.frexp(np.abs(x).max())[1]
    scale = 2.0 ** (46 - e)
    magic = 6755399441055744.0
    bits = (x * scale + magic).view(np.int64) - np.float64(magic).view(np.int64)
    assert np.array_equal(bits, np.rint(x * scale).astype(np.int64))


def test_recombination_is_exact_and_fits_int32():
    """sum_{s+t<=5} A^(s) B^(t)^T 256^(5-s-t), accumulated per diagonal in INT32 over K = 4096 (the wsyrk block), recombined as
    hi 2^24 + lo with both halves exact in FP64 -- equal to the integer reference; worst-case digits stay below 2^31."""
    rng = np.random.default_rng(2)
    M, N, K = 8, 6, 4096
    va = rng.integers(-(1 << 46), (1 << 46), size=(M, K), dtype=np.int64)
    vb = rng.integers(-(1 << 46), (1 << 46), size=(N, K), dtype=np.int64)
    A = digits_via_bias(va)[..., ::-1]  # plane s = 0 most significant
    B = digits_via_bias(vb)[..., ::-1]
    P = np.zeros((ND, M, N), dtype=np.int64)
    for s in range(ND):
        for t in range(ND - s):
            P[s + t] += A[..., s] @ B[..., t].T
    assert np.abs(P).max() < 2 ** 31
    worst = 6 * 128 * 128 * K  # six products of extreme digits on the last diagonal
    assert worst < 2 ** 31
    hi = (P[0] << 16) + (P[1] << 8) + P[2]
    lo = (P[3] << 16) + (P[4] << 8) + P[5]
    assert np.abs(hi).max() < 2 ** 53 and np.abs(lo).max() < 2 ** 53  # exact as doubles
    ref = sum(int(1) * (P[d].astype(object) * (256 ** (5 - d))) for d in range(ND))
    got = hi.astype(object) * (1 << 24) + lo.astype(object)
    assert (ref == got).all()
    # what the scheme drops: the products with s + t >= 6, relative to the exact integer product
    exact = va.astype(object) @ vb.astype(object).T
    kept = got * (256 ** 5)
    rel = np.abs(np.array((exact - kept) / (2.0 ** 92 * K), dtype=np.float64)).max()  # relative to (row max)(col max) K
    assert rel < 2.0 ** -45


def test_row_scaled_integers_reproduce_the_product_to_1e12():
    """End to end in float: operands -> 46-bit fixed point per row / column -> exact integer product -> one rounding."""
    rng = np.random.default_rng(3)
    M, N, K = 16, 12, 1000
    X = rng.standard_normal((M, K)) * np.exp(-3 * rng.random((M, 1)))
    Y = rng.standard_normal((N, K)) * np.exp(-3 * rng.random((N, 1)))
    ex = np.frexp(np.abs(X).max(1))[1]
    ey = np.frexp(np.abs(Y).max(1))[1]
    vx = np.rint(X * 2.0 ** (46 - ex)[:, None]).astype(np.int64)
    vy = np.rint(Y * 2.0 ** (46 - ey)[:, None]).astype(np.int64)
    A, B = digits_via_bias(vx)[..., ::-1], digits_via_bias(vy)[..., ::-1]
    acc = np.zeros((M, N), dtype=object)
    for s in range(ND):
        for t in range(ND - s):
            acc += (A[..., s] @ B[..., t].T).astype(object) * (256 ** (5 - s - t))
    got = np.array(acc, dtype=np.float64) * 2.0 ** (ex[:, None] + ey[None, :] - 52)
    ref = X @ Y.T
    assert np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max()
