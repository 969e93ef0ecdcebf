"""CPU-only executable specification of the S4-convolution kernels' algorithm (csrc/fftconv*.cu,
csrc/s4_kernelgen.cu), restated in numpy and checked against the reference formula
(models/s4.py:1391-1406: two-sided kernel, rfft/irfft product at n = 2l):

  * the packed real transform: z[i] = y[2i] + i y[2i+1], one M = n/2 point complex FFT, the 2x2 pair map
    (alpha, beta, gamma, delta) of kcoef_kernel that fuses untangle + spectrum product + re-tangle;
  * the split of the half-empty packed row into an even-frequency and an odd-frequency half transform
    (fftconv2/3_kernel), their pairings in bit-reversed slot order (partner = p ^ ((1 << msb p) - 1) resp. the
    complement), and the recombination a[i] + W_M^{-i} b[i];
  * the compact table of kcoef_compact_kernel: S = 2(K1 + K2), D = 2(K1 - K2), alpha = S + ws D, delta = S - ws D,
    beta = -gamma = i wc D, and the relation k_B = n/4 - k_A (w_B = -i conj(w_A)) between the two pairs of a centre item.
No product code runs here; the CUDA kernels are compared with the oracle in tests/test_gpu_ops.py."""
import numpy as np
import pytest


def brev(x, bits):
    r = 0
    for i in range(bits):
        r |= ((x >> i) & 1) << (bits - 1 - i)
    return r


def reference_conv(y, k0, k1):
    l = y.shape[-1]
    n = 2 * l
    kk = np.concatenate([k0, np.zeros(l)]) + np.concatenate([np.zeros(l), k1[::-1]])
    return np.fft.irfft(np.fft.rfft(y, n) * np.fft.rfft(kk, n), n)[:l]


def kernel_spectrum(k0, k1, log2M):
    """K'[f], f = 0..M, of the wrapped two-sided kernel at n = 2M, scaled by 1/(4M) (kf_kernel; D omitted)."""
    M, l = 1 << log2M, k0.shape[0]
    n = 2 * M
    kk = np.zeros(n)
    kk[:l] = k0
    kk[n - l:] += k1[::-1]                    # kk[n - s] = k1[s - 1], s = 1..l
    return np.fft.fft(kk)[: M + 1] / (4.0 * M)


def pair_coeffs(K, M, k):
    """(alpha, beta, gamma, delta) of kcoef_kernel for the pair (k, M - k)."""
    n = 2 * M
    K1, K2 = K[k], np.conj(K[M - k])
    w = np.exp(-2j * np.pi * k / n)
    u, v = 1 - 1j * w, 1 + 1j * w
    alpha = K1 * abs(u) ** 2 + K2 * abs(v) ** 2
    beta = K1 * v * np.conj(u) + K2 * u * np.conj(v)
    gamma = K1 * u * np.conj(v) + K2 * v * np.conj(u)
    delta = K1 * abs(v) ** 2 + K2 * abs(u) ** 2
    return alpha, beta, gamma, delta


def pair_map(a, b, c):
    alpha, beta, gamma, delta = c
    return alpha * a + beta * np.conj(b), np.conj(gamma * a + delta * np.conj(b))


def pack(y, M):
    z = np.zeros(M, complex)
    l = y.shape[0]
    z[: (l + 1) // 2] = np.pad(y, (0, l % 2))[0::2] + 1j * np.pad(y, (0, l % 2))[1::2]
    return z


def unpack(z, l):
    out = np.empty(2 * z.shape[0])
    out[0::2], out[1::2] = z.real, z.imag
    return out[:l]


@pytest.mark.parametrize("l,log2M", [(16, 4), (100, 7), (250, 8), (1000, 10)])
def test_packed_transform_with_pair_map(l, log2M):
    """v1: one M-point transform, natural order."""
    rng = np.random.default_rng(l)
    y, k0, k1 = rng.standard_normal(l), rng.standard_normal(l) * 0.3, rng.standard_normal(l) * 0.3
    M = 1 << log2M
    K = kernel_spectrum(k0, k1, log2M)
    Z = np.fft.fft(pack(y, M))
    Zp = np.zeros(M, complex)
    a = Z[0]                                              # DC and Nyquist, both real
    p0, pM = 2 * (a.real + a.imag) * K[0].real, 2 * (a.real - a.imag) * K[M].real
    Zp[0] = complex(p0 + pM, p0 - pM)
    for k in range(1, M // 2 + 1):
        oa, ob = pair_map(Z[k], Z[M - k], pair_coeffs(K, M, k))
        Zp[k] = oa
        if k != M - k:
            Zp[M - k] = ob
    got = unpack(np.fft.ifft(Zp) * M, l)                  # unnormalised inverse: the 1/(4M) sits in K'
    np.testing.assert_allclose(got, reference_conv(y, k0, k1), rtol=0, atol=1e-10)


@pytest.mark.parametrize("l,log2M", [(16, 4), (100, 7), (500, 9), (1000, 10)])
def test_split_transform_pairings_and_compact_table(l, log2M):
    """v2 / v3: two half transforms in bit-reversed slot order, radix-2 centre items, compact coefficients."""
    rng = np.random.default_rng(l + 1)
    y, k0, k1 = rng.standard_normal(l), rng.standard_normal(l) * 0.3, rng.standard_normal(l) * 0.3
    M, LH = 1 << log2M, log2M - 1
    Mh, n = M // 2, 2 * M
    K = kernel_spectrum(k0, k1, log2M)
    z = pack(y, M)
    assert np.all(z[Mh:] == 0)                            # the half-empty packed row the split relies on
    halves = [np.fft.fft(z[:Mh]), np.fft.fft(z[:Mh] * np.exp(-2j * np.pi * np.arange(Mh) / M))]
    Zfull = np.fft.fft(z)
    np.testing.assert_allclose(halves[0], Zfull[0::2], atol=1e-9)
    np.testing.assert_allclose(halves[1], Zfull[1::2], atol=1e-9)
    outs = []
    for odd, Zh in enumerate(halves):
        slots = np.array([Zh[brev(p, LH)] for p in range(Mh)])          # what the forward passes leave in shared memory
        new = slots.copy()
        G = Mh // 2                                                     # radix-2 centre: groups of two slots
        done = np.zeros(Mh, bool)
        for item in range(Mh // 4):
            ga = 2 * item
            gb = (ga ^ (G - 1)) if odd else (1 if ga == 0 else ga ^ ((1 << (ga.bit_length() - 1)) - 1))
            if not odd and item == 0:
                a = slots[0]                                            # slot 0: DC / Nyquist
                p0, pM = 2 * (a.real + a.imag) * K[0].real, 2 * (a.real - a.imag) * K[M].real
                new[0] = complex(p0 + pM, p0 - pM)
                new[1], _ = pair_map(slots[1], slots[1], pair_coeffs(K, M, 2 * brev(1, LH)))      # k = M/2, self-paired
                new[2], new[3] = pair_map(slots[2], slots[3], pair_coeffs(K, M, 2 * brev(2, LH)))
                done[:4] = True
                continue
            kA = 2 * brev(2 * ga, LH) + odd
            kB = 2 * brev(2 * gb, LH) + odd
            assert kB == n // 4 - kA                                    # one twiddle per item: w_B = -i conj(w_A)
            wA = np.exp(-2j * np.pi * kA / n)
            for (lead, part, k, w) in ((2 * ga, 2 * gb + 1, kA, wA), (2 * gb, 2 * ga + 1, kB, -1j * np.conj(wA))):
                assert brev(part, LH) == (Mh - 1 - brev(lead, LH) if odd else (Mh - brev(lead, LH)) % Mh)   # slot of M - k
                K1, K2 = K[k], np.conj(K[M - k])
                S, D = 2 * (K1 + K2), 2 * (K1 - K2)                     # the 16 bytes kcoef_compact_kernel stores
                wc, ws = w.real, w.imag
                c = (S + ws * D, 1j * wc * D, -1j * wc * D, S - ws * D)
                np.testing.assert_allclose(c, pair_coeffs(K, M, k), atol=1e-12)
                new[lead], new[part] = pair_map(slots[lead], slots[part], c)
                assert not done[lead] and not done[part]
                done[lead] = done[part] = True
        assert done.all()
        Zp = np.empty(Mh, complex)
        for p in range(Mh):
            Zp[brev(p, LH)] = new[p]
        outs.append(np.fft.ifft(Zp) * Mh)
    zz = outs[0] + np.exp(2j * np.pi * np.arange(Mh) / M) * outs[1]     # a[i] + W_M^{-i} b[i]
    np.testing.assert_allclose(unpack(zz, l), reference_conv(y, k0, k1), rtol=0, atol=1e-10)


def test_plane_padding_offset_identity():
    """mid_pass (fftconv3.cu) addresses element base + c*sub of a butterfly as padf(base) + c*sub + 2*((c*sub) >> 5)
    with a compile-time second term: valid for every span the plans use (sub a power of two, base = blk*16*sub + j,
    j < sub), and the re/im plane entries of butterflies j, j+1 (j even) are adjacent and 8-byte aligned."""
    padf = lambda i: i + 2 * (i >> 5)
    for log2sub in range(1, 11):
        sub = 1 << log2sub
        S = 16 * sub
        for blk in (0, 1, 3, 7):
            for j in range(0, min(sub, 64), 2):
                base = blk * S + j
                for c in range(16):
                    assert padf(base + c * sub) == padf(base) + c * sub + 2 * ((c * sub) >> 5)
                    assert padf(base + c * sub) % 2 == 0 and padf(base + c * sub + 1) == padf(base + c * sub) + 1
    # radix-4 pass of span 16 (mid4_pass): + 4p stays inside the 16-float block
    for blk in range(64):
        for j in (0, 2):
            base = blk * 16 + j
            for p in range(4):
                assert padf(base + 4 * p) == padf(base) + 4 * p


@pytest.mark.parametrize("log2M,radices", [(13, (4, 4, 4, 1)), (12, (4, 4, 2, 2)), (10, (4, 4, 2)), (11, (4, 4, 3))])
def test_dif_passes_leave_plain_bit_reversal(log2M, radices):
    """Forward passes (any radix split): u_q = sum_p x[j + p sub] w_R^{pq}, stored times W_S^{jq} at j + brev(q) sub.
    After all passes slot p holds X[brev(p)], whatever the radices — the property that lets one pointwise table serve
    every plan and gives the partner maps of the untangle step."""
    assert sum(radices) == log2M
    M = 1 << log2M
    rng = np.random.default_rng(log2M)
    x = rng.standard_normal(M) + 1j * rng.standard_normal(M)
    s = x.copy()
    span_log = log2M
    for lr in radices:
        R, S = 1 << lr, 1 << span_log
        sub = S // R
        out = np.empty_like(s)
        for blk in range(M // S):
            for j in range(sub):
                base = blk * S + j
                xin = s[base + sub * np.arange(R)]
                u = np.fft.fft(xin)                                   # radix-R DFT
                u = u * np.exp(-2j * np.pi * j * np.arange(R) / S)    # twiddle W_S^{jq}
                for q in range(R):
                    out[base + brev(q, lr) * sub] = u[q]
        s = out
        span_log -= lr
    X = np.fft.fft(x)
    np.testing.assert_allclose(s, X[[brev(p, log2M) for p in range(M)]], atol=1e-8)
