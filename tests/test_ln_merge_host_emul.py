"""Host emulation (numpy float32, same operation order) of the row-statistics exchange of the residual + LayerNorm GEMM
epilogue (beach_seg_b200/csrc/gemm.cuh, EPI_RESID_LN): per-thread 4-value chunks merged sequentially, 8-lane butterfly,
tagged 64-bit entry, 8-slice butterfly -> (mean, rstd).  Checks what DESIGN.md claims for it: no E[x^2] - mean^2
cancellation at |mean| = 100 std, the pairwise merge is bitwise symmetric (every CTA that merges the same partials gets
the same bits), and the 8-bit launch tag in the low mantissa bits costs 2^-20 relative."""
import numpy as np

F = np.float32


def merge_equal(mean_a, m2_a, mean_b, m2_b, half_count):
    """ln_merge_equal: two partials of equal count n = 2 * half_count."""
    d = F(mean_b - mean_a)
    m2 = F(F(d * d) * F(half_count) + F(m2_a + m2_b))  # (device: one fma; the difference is below what is asserted)
    return F(F(0.5) * F(mean_a + mean_b)), m2


def thread_partial(vals16):
    """gemm_epi_resid_ln_chunk: four chunks of four values of one row, merged in chunk order."""
    mean, m2 = F(0), F(0)
    w_mean = [F(1.0), F(0.5), F(1.0) / F(3.0), F(0.25)]
    w_m2 = [F(0.0), F(2.0), F(8.0) / F(3.0), F(3.0)]
    for ci in range(4):
        o = vals16[4 * ci:4 * ci + 4].astype(F)
        m4 = F(F(0.25) * F(F(o[0] + o[1]) + F(o[2] + o[3])))
        dd = (o - m4).astype(F)
        q4 = F(F(dd[0] * dd[0] + dd[1] * dd[1]) + F(dd[2] * dd[2] + dd[3] * dd[3]))
        delta = F(m4 - mean)
        mean = F(delta * w_mean[ci] + mean)
        m2 = F(m2 + F(F(delta * delta) * w_m2[ci] + q4))
    return mean, m2


def butterfly(parts, n_each):
    """xor-1, xor-2, xor-4 butterfly over 8 partials of n_each values; returns what every lane ends up with."""
    parts = list(parts)
    n = n_each
    for o in (1, 2, 4):
        parts = [merge_equal(*parts[i], *parts[i ^ o], n / 2) for i in range(8)]
        n *= 2
    return parts


def pack(mean, m2, tag):
    lo = (np.array(mean, F).view(np.uint32) & np.uint32(0xFFFFFFF0)) | np.uint32(tag >> 4)
    hi = (np.array(m2, F).view(np.uint32) & np.uint32(0xFFFFFFF0)) | np.uint32(tag & 15)
    return (np.uint64(hi) << np.uint64(32)) | np.uint64(lo)


def unpack(e):
    lo, hi = np.uint32(e & np.uint64(0xFFFFFFFF)), np.uint32(e >> np.uint64(32))
    tag = ((int(lo) & 15) << 4) | (int(hi) & 15)
    return (lo & np.uint32(0xFFFFFFF0)).view(F), (hi & np.uint32(0xFFFFFFF0)).view(F), tag


def row_stats(row, tag):
    """One row of 1024 values -> (mean, rstd) the way the four N-tile CTAs compute it."""
    entries = []
    for s in range(8):  # slot = 2 * (N tile) + column half: 128 columns, eight lanes x 16 values
        sl = row[128 * s:128 * (s + 1)]
        # lane l of the row's group holds columns 32 * chunk + 4 * l .. + 3 of the slice
        lanes = [thread_partial(np.concatenate([sl[32 * c + 4 * l:32 * c + 4 * l + 4] for c in range(4)])) for l in range(8)]
        after = butterfly(lanes, 16)
        assert all(a[0].tobytes() == after[0][0].tobytes() and a[1].tobytes() == after[0][1].tobytes() for a in after)
        entries.append(pack(after[0][0], after[0][1], tag))
    parts = []
    for e in entries:
        m, q, t = unpack(e)
        assert t == tag
        parts.append((m, q))
    final = butterfly(parts, 128)
    assert all(a[0].tobytes() == final[0][0].tobytes() and a[1].tobytes() == final[0][1].tobytes() for a in final)
    mean, m2 = final[0]
    return mean, F(1.0) / np.sqrt(F(m2 * F(1.0 / 1024.0) + F(1e-6)), dtype=F)


def test_statistics_exchange_is_accurate_and_symmetric():
    rng = np.random.default_rng(0)
    for offset, scale in ((0.0, 1.0), (30.0, 1.0), (100.0, 1.0), (-7.0, 0.02)):
        for tag in (0, 1, 47, 254):
            row = (rng.standard_normal(1024) * scale + offset).astype(F)
            mean, rstd = row_stats(row, tag)
            mean64 = row.astype(np.float64).mean()
            rstd64 = 1.0 / np.sqrt(row.astype(np.float64).var() + 1e-6)
            # fp32 arithmetic on 1024 values + 2^-20 from the tag nibbles; a bf16 ulp is 3.9e-3
            assert abs(float(mean) - mean64) <= 4e-6 * max(abs(mean64), scale), (offset, tag)
            # (the dropped nibble of a partial MEAN is 2^-20 of |mean|: it enters the between-slice term of M2, so the
            # bound grows with |mean| / std -- 350 in the last case, still 100x below a bf16 ulp)
            tol = 1e-5 * (1.0 + abs(offset) / scale / 100.0)
            assert abs(float(rstd) - rstd64) <= tol * rstd64, (offset, tag, float(rstd), rstd64)
            # the naive formulation the kernel avoids loses the variance at |mean| = 100 std
            if offset == 100.0:
                s, q = F(0), F(0)
                for v in row:
                    s, q = F(s + v), F(q + F(v * v))
                naive_var = float(F(q / F(1024)) - F(F(s / F(1024)) * F(s / F(1024))))
                assert abs(naive_var - row.astype(np.float64).var()) > 1e-3


def test_tag_round_trip_and_sentinel():
    for tag in range(255):
        m, q, t = unpack(pack(F(1.2345678), F(987.65432), tag))
        assert t == tag
        assert abs(float(m) - 1.2345678) <= 1.2345678 * 2.0 ** -19 and abs(float(q) - 987.65432) <= 987.65432 * 2.0 ** -19
    # a buffer filled with 0xFF bytes ("never written") decodes to tag 255, which no launch uses
    assert unpack(np.uint64(0xFFFFFFFFFFFFFFFF))[2] == 255
