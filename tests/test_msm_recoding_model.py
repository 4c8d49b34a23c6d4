"""Executable model of the device MSM's signed-digit recoding and window bookkeeping (csrc/msm.cu msm_count), in pure
Python: digits lie in [-2^(c-1), 2^(c-1)], reconstruct the scalar, need exactly W = 254 // c + 1 windows, and the
bucket-method identity  sum_i s_i P_i = sum_w 2^(cw) sum_b b * B_{w,b}  holds on a toy group (integers mod a prime)."""
import random

import pytest

R = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001


def recode(s, c):
    W = 254 // c + 1
    mask, half = (1 << c) - 1, 1 << (c - 1)
    out, carry = [], 0
    for _ in range(W):
        v = (s & mask) + carry
        s >>= c
        if v > half:
            out.append(v - (1 << c))  # negative digit (0 when v == 2^c), carry into the next window
            carry = 1
        else:
            out.append(v)
            carry = 0
    assert carry == 0 and s == 0, "top window must absorb the last carry"
    return out


@pytest.mark.parametrize("c", list(range(2, 25)))
def test_signed_digits_reconstruct_every_scalar(c):
    rng = random.Random(c)
    half = 1 << (c - 1)
    edge = [0, 1, half, half + 1, (1 << c) - 1, 1 << c, R - 1, R - 2, (1 << 253), (1 << 254) - 1 if (1 << 254) - 1 < R else R - 3]
    for s in edge + [rng.randrange(R) for _ in range(300)]:
        d = recode(s, c)
        assert len(d) == 254 // c + 1
        assert all(-half <= x <= half for x in d)
        assert sum(x << (c * w) for w, x in enumerate(d)) == s


@pytest.mark.parametrize("c", [2, 5, 13, 16, 22])
def test_bucket_method_identity_on_a_toy_group(c):
    """points are integers mod q (additive group): MSM = sum s_i P_i mod q computed through signed buckets"""
    q = (1 << 61) - 1
    rng = random.Random(100 + c)
    n = 200
    pts = [rng.randrange(q) for _ in range(n)]
    sc = [rng.randrange(R) for _ in range(n)]
    sc[0], sc[1] = 0, R - 1
    W, B = 254 // c + 1, 1 << (c - 1)
    buckets = [[0] * (B + 1) for _ in range(W)]
    for s, p in zip(sc, pts):
        for w, d in enumerate(recode(s, c)):
            if d > 0:
                buckets[w][d] = (buckets[w][d] + p) % q
            elif d < 0:
                buckets[w][-d] = (buckets[w][-d] - p) % q
    acc = 0
    for w in reversed(range(W)):  # Horner over windows
        acc = (acc << c) % q
        acc = (acc + sum(b * buckets[w][b] for b in range(1, B + 1))) % q
    assert acc == sum(s * p for s, p in zip(sc, pts)) % q
    # the row/column reduction used on the device: S = WS(Col) + 2^kc WS(Row) + sum(Row) over i = hi*2^kc + lo
    kc = c // 2
    cols, rows = 1 << kc, B >> kc
    for w in range(W):
        bw = buckets[w][1:]
        row = [sum(bw[h * cols + l] for l in range(cols)) % q for h in range(rows)]
        col = [sum(bw[h * cols + l] for h in range(rows)) % q for l in range(cols)]
        s2 = (sum(l * col[l] for l in range(cols)) + (sum(h * row[h] for h in range(rows)) << kc) + sum(row)) % q
        assert s2 == sum(b * buckets[w][b] for b in range(1, B + 1)) % q


@pytest.mark.parametrize("c", [2, 3, 5, 13, 17])
def test_bit_sliced_weighted_sums_and_batched_bucket_sets(c):
    """msm_bit_sums / msm_finish: WS(V) = sum_j j V_j is folded from the q subset sums S_b = sum_{j: bit b of j} V_j by Horner with
    doublings (acc = 2 acc + S_b from the top bit down), for the Row and the Col vector of every bucket set; and the batched
    pipeline's set index col * Ws + w keeps the columns of a batch apart (msm_run_batch)."""
    qmod = (1 << 61) - 1
    rng = random.Random(300 + c)
    B = 1 << (c - 1)
    kc = c // 2
    cols, rows = 1 << kc, B >> kc

    def ws_bit_sliced(v):
        m = len(v)
        q = 0
        while (1 << q) < m:
            q += 1
        h = 0
        for b in reversed(range(q)):
            h = (2 * h + sum(v[j] for j in range(m) if (j >> b) & 1)) % qmod
        return h

    for _ in range(3):
        bw = [rng.randrange(qmod) if rng.random() < 0.7 else 0 for _ in range(B)]
        row = [sum(bw[h * cols + l] for l in range(cols)) % qmod for h in range(rows)]
        col = [sum(bw[h * cols + l] for h in range(rows)) % qmod for l in range(cols)]
        assert ws_bit_sliced(row) == sum(j * x for j, x in enumerate(row)) % qmod
        assert ws_bit_sliced(col) == sum(j * x for j, x in enumerate(col)) % qmod
        s = (ws_bit_sliced(col) + (ws_bit_sliced(row) << kc) + sum(row)) % qmod
        assert s == sum((b + 1) * x for b, x in enumerate(bw)) % qmod  # bucket index b holds digit magnitude b + 1
    # a batch of 3 columns over the same points: one histogram / sort keyed by (col * Ws + w) * B + bucket
    n, batch = 60, 3
    W = 254 // c + 1
    pts = [rng.randrange(qmod) for _ in range(n)]
    scal = [[rng.randrange(R) for _ in range(n)] for _ in range(batch)]
    sets = [[0] * B for _ in range(batch * W)]
    for colj in range(batch):
        for s, p in zip(scal[colj], pts):
            for w, d in enumerate(recode(s, c)):
                if d:
                    key = colj * W + w
                    sets[key][abs(d) - 1] = (sets[key][abs(d) - 1] + (p if d > 0 else -p)) % qmod
    for colj in range(batch):
        acc = 0
        for w in reversed(range(W)):
            acc = ((acc << c) + sum((b + 1) * x for b, x in enumerate(sets[colj * W + w]))) % qmod
        assert acc == sum(s * p for s, p in zip(scal[colj], pts)) % qmod
