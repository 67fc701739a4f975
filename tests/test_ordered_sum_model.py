"""Host model of csrc/ordered_sum.cu: the parallel evaluation of the left-to-right FP64 sum
    s = x[0];  s = fl(s + x[i])
(SURVEY A.10 area CDF, A.8 outlier statistics) restated with Python integers, checked bit for bit against the
scalar chain.  It documents WHY the kernels are exact: inside a binade the accumulator is m * u and
fl(m*u + x) = (m + rn(x / u)) * u unless x / u is an exact tie or the sum leaves the binade; every shortcut is
verified against the exact accumulator, a failed check falls back to the scalar chain for that chunk.
(The CUDA kernels themselves are compared with the scalar-chain kernel in tests/test_gpu_ordered_sum.py.)"""
import struct

import numpy as np
import pytest

L = 256
TWO52, TWO53 = 1 << 52, 1 << 53


def bits(x):
    return struct.unpack("<q", struct.pack("<d", x))[0]


def frombits(b):
    return struct.unpack("<d", struct.pack("<q", b))[0]


def quantise(v, e):
    """rn(v / 2^(e-52)) and the 'irregular' flag (exact tie, or a term that does not fit below 2^(e+1))."""
    b = bits(v)
    eb = (b >> 52) & 0x7FF
    if eb == 0:
        return 0, False
    ev, mv, sh = eb - 1023, (b & (TWO52 - 1)) | TWO52, e - (eb - 1023)
    if sh < 0:
        return 0, True
    if sh == 0:
        return mv, False
    if sh >= 54:
        return 0, False
    half = 1 << (sh - 1)
    rem, k = mv & ((half << 1) - 1), mv >> sh
    return (k + 1 if rem > half else k), rem == half


def from_int(M, e):
    return frombits((e + 1 + 1023) << 52) if M >= TWO53 else frombits(((e + 1023) << 52) | (M - TWO52))


def scalar_chain(x):
    s, out = float(x[0]), [float(x[0])]
    for v in x[1:]:
        s = s + float(v)
        out.append(s)
    return np.array(out)


def parallel_chain(x):
    n, nc = len(x), (len(x) + L - 1) // L
    csum = np.array([x[c * L:(c + 1) * L].sum() if (x[c * L:(c + 1) * L] >= 0).all() and np.isfinite(x[c * L:(c + 1) * L]).all() else np.nan
                     for c in range(nc)])
    cpre = np.concatenate([[0.0], np.cumsum(csum)[:-1]])
    ce, cK = [None] * nc, [0] * nc
    for c in range(nc):                                           # kernel B
        p0, p1 = cpre[c], cpre[c] + csum[c]
        if not (c > 0 and p0 > 1e-270 and p1 < 1e270):
            continue
        e = ((bits(float(p0)) >> 52) & 0x7FF) - 1023
        if not (p0 * (1 - 1e-6) > 2.0 ** e and p1 * (1 + 1e-6) < 2.0 ** (e + 1)):
            continue
        K, irregular = 0, False
        for v in x[c * L:(c + 1) * L]:
            k, t = quantise(float(v), e)
            K, irregular = K + k, irregular or t
        if not irregular:
            ce[c], cK[c] = e, K
    out, s, c, replayed = np.empty(n), 0.0, 0, 0
    while c < nc:                                                 # kernel C: 32 chunks per step
        sb = bits(s)
        e_s, m = ((sb >> 52) & 0xFFF) - 1023, (sb & (TWO52 - 1)) | TWO52
        inc, run = 0, 0
        for j in range(c, min(c + 32, nc)):
            carry_in = m + inc
            inc += cK[j]
            if not (ce[j] is not None and ce[j] == e_s and carry_in < TWO53 and m + inc <= TWO53):
                break
            run += 1
            base = carry_in                                       # kernel D: per-element prefixes of a committed chunk
            acc = 0
            for i, v in enumerate(x[j * L:(j + 1) * L]):
                acc += quantise(float(v), e_s)[0]
                out[j * L + i] = from_int(base + acc, e_s)
        if run:
            s = from_int(m + sum(cK[c:c + run]), e_s)
        c += run
        if run < 32 and c < nc:                                   # replay the run-breaker with the scalar chain
            replayed += 1
            for i, v in enumerate(x[c * L:(c + 1) * L]):
                s = float(v) if (c == 0 and i == 0) else s + float(v)
                out[c * L + i] = s
            c += 1
    return out, replayed


def _cases():
    rng = np.random.default_rng(1)
    yield "areas", 1e-5 * (0.25 + rng.random(40000))
    yield "wide", np.exp2(rng.integers(-64, 64, 20000).astype(np.float64)) * (1 + rng.random(20000))
    yield "grid", (1 + rng.integers(0, 1024, 30000)) / 1024.0
    yield "jumps", np.where(rng.random(20000) < 0.8, np.where(rng.random(20000) < 0.5, 0.0, 1e-3 * rng.random(20000)), 1e3 * rng.random(20000))
    ties = np.concatenate([[1.0], (2 * rng.integers(0, 50, 20000) + 1) * 2.0 ** -53])      # every term is an exact tie at first
    yield "ties", ties
    edge = np.concatenate([[1.0], np.full(5000, 2.0 ** -13), np.full(5000, 2.0 ** -60)])   # lands exactly on 2.0, then tiny terms
    yield "binade_edge", edge
    neg = 1e-3 * rng.random(10000)
    neg[5000] = -1.0
    yield "negative_term", neg


@pytest.mark.parametrize("name,x", list(_cases()), ids=[c[0] for c in _cases()])
def test_parallel_chain_is_bit_identical(name, x):
    want = scalar_chain(x)
    got, replayed = parallel_chain(x)
    assert (want.view(np.int64) == got.view(np.int64)).all()
    n_chunks = (len(x) + L - 1) // L
    if name in ("areas", "grid"):
        assert replayed < n_chunks // 4          # the shortcut carries almost all chunks
