"""CPU checks (numpy, float32 arithmetic as on the device) of the two arguments the grouped leave-one-out count
kernel rests on (csrc/fast_count_loo_group.cu, csrc/fast_count.cuh SpecLoo):

1. the indicator-sum form counts exactly what the D separate maxima count:
       [max_{d != s} a_d < T]  ==  [sigma >= DP - 1] - [sigma == DP - 1] * i_s,   i_d = [a_d < T], sigma = sum_d i_d
2. a (query group, sub-tile) pair whose SECOND-LARGEST box gap is >= the group's largest threshold holds no pair
   that counts in any leave-one-out subspace (fp32 subtraction is monotone, so a box gap computed in fp32 is a
   lower bound of every |difference| computed in fp32), and the rule never fires when some pair does count."""
import numpy as np


def _loo_counts_by_maxima(a, T):
    """a: (pairs, D) |differences| (float32), T: (pairs,) thresholds -> (pairs, D) 0/1"""
    D = a.shape[1]
    out = np.zeros(a.shape, dtype=np.int64)
    for s in range(D):
        others = np.delete(a, s, axis=1)
        out[:, s] = others.max(axis=1) < T
    return out


def test_indicator_sums_equal_the_maxima_form():
    rng = np.random.default_rng(5)
    for D in (5, 6, 7, 8):
        DP = (D + 1) & ~1
        a = np.abs(rng.standard_normal((200_000, D))).astype(np.float32)
        # thresholds around the typical |difference|, plus exact ties with some coordinate
        T = np.abs(rng.standard_normal(a.shape[0])).astype(np.float32) * 1.5
        tie = rng.random(a.shape[0]) < 0.2
        T[tie] = a[tie, rng.integers(0, D, size=int(tie.sum()))]
        ref = _loo_counts_by_maxima(a, T)
        ap = np.zeros((a.shape[0], DP), dtype=np.float32)  # the pad coordinate differs by 0: always inside
        ap[:, :D] = a
        i = (ap < T[:, None]).astype(np.float32)
        sigma = i.sum(axis=1)
        ge = np.clip(sigma - (DP - 2), 0.0, 1.0)           # [sigma >= DP - 1]
        e = ge - np.clip(sigma - (DP - 1), 0.0, 1.0)       # [sigma == DP - 1]
        got = (ge[:, None] - e[:, None] * i[:, :D]).astype(np.int64)
        assert np.array_equal(got, ref), D
        # (the block-level kernel's form: [sigma == DP] + [sigma == DP - 1] (1 - i_s))
        allin = np.clip(sigma - (DP - 1), 0.0, 1.0)
        got2 = (allin[:, None] + e[:, None] * (1.0 - i[:, :D])).astype(np.int64)
        assert np.array_equal(got2, ref), D


def _second_largest(g):
    s = np.sort(g, axis=-1)
    return s[..., -2]


def test_second_largest_gap_rule_never_drops_a_counted_pair():
    rng = np.random.default_rng(11)
    fired = kept = 0
    for D in (5, 6, 8):
        for trial in range(400):
            scale = np.float32(10.0 ** rng.uniform(-3, 3))
            centre = (rng.standard_normal(D) * scale * 3).astype(np.float32)
            q = (centre + rng.standard_normal((8, D)).astype(np.float32) * scale * np.float32(0.3)).astype(np.float32)
            off = rng.standard_normal(D).astype(np.float32) * scale * np.float32(rng.uniform(0, 4))
            off[rng.random(D) < 0.5] = 0.0   # near in some coordinates, far in others: the interesting cases
            c = (centre + off + rng.standard_normal((32, D)).astype(np.float32) * scale * np.float32(0.3)).astype(np.float32)
            T = (np.abs(rng.standard_normal(8)).astype(np.float32) * scale).astype(np.float32)
            glo, ghi = q.min(axis=0), q.max(axis=0)
            blo, bhi = c.min(axis=0), c.max(axis=0)
            gap = np.maximum(blo - ghi, glo - bhi)          # float32, > 0 when the boxes are apart
            skip = _second_largest(gap) >= T.max()
            a = np.abs(q[:, None, :] - c[None, :, :])        # float32 |differences| of all 256 pairs
            counted = _loo_counts_by_maxima(a.reshape(-1, D), np.repeat(T, 32)).any()
            if skip:
                fired += 1
                assert not counted, (D, trial)
                # the argument itself: at least two coordinates are outside for every pair
                assert ((a >= T[:, None, None]).sum(axis=2) >= 2).all()
            else:
                kept += 1
    assert fired > 100 and kept > 100  # both outcomes were exercised


def test_group_classification_of_the_own_order_count_kernel():
    """count_gunits_kernel (csrc/fast_count_group.cuh): a (group, sub-tile) pair is OUTSIDE when the nearest box
    distance is >= the group's largest threshold (no pair counts), INSIDE when the farthest box distance is below the
    smallest (all 8 x 32 pairs count), else scanned -- in float32, for 2-D and 4-D sets."""
    rng = np.random.default_rng(13)
    seen = {"outside": 0, "inside": 0, "scan": 0}
    for D in (2, 4):
        for trial in range(1500):
            scale = np.float32(10.0 ** rng.uniform(-2, 2))
            q = (rng.standard_normal((8, D)).astype(np.float32) * scale * np.float32(0.05)).astype(np.float32)
            c = ((rng.standard_normal(D) * rng.uniform(0, 2)).astype(np.float32) * scale
                 + rng.standard_normal((32, D)).astype(np.float32) * scale * np.float32(0.05)).astype(np.float32)
            T = (np.float32(rng.uniform(0.05, 3.0)) * scale * (1 + np.float32(0.2) * rng.random(8).astype(np.float32))).astype(np.float32)
            glo, ghi, blo, bhi = q.min(axis=0), q.max(axis=0), c.min(axis=0), c.max(axis=0)
            a, b = blo - ghi, glo - bhi
            near = np.maximum(np.float32(0), np.maximum(a, b).max())
            far = np.maximum(-a, -b).max()
            m = np.abs(q[:, None, :] - c[None, :, :]).max(axis=2)   # Chebyshev distances, float32
            inside_pairs = m < T[:, None]
            if not near < T.max():
                seen["outside"] += 1
                assert not inside_pairs.any()
            elif far < T.min():
                seen["inside"] += 1
                assert inside_pairs.all()
            else:
                seen["scan"] += 1
    assert min(seen.values()) > 50, seen
