"""Developer tool: counts how many warp-wide count(key >= pivot) steps the top-k threshold search needs per
row, for the pivot rules discussed in DESIGN.md (3.4 and 9.4), on several value distributions.  Pure numpy, no GPU.

    python tools/topk_search_sim.py

rules:
  secant      what csrc/topk.cu does: bracket [min over lanes of the lane's ceil(k/32)-th largest key, row max],
              secant step on the counts in key space, plain bisection every fourth step
  +m2         the same with the free second point (M2 + 1, #lane maxima > M2) used for k <= 16
  lane-rank   first two pivots taken from the SORTED lane maxima at the rank where the k-th largest of the row
              is expected (distribution-free: the top-k of a row hit 32 (1 - (31/32)^k) distinct lanes), then
              the secant rule inside that bracket.  Needs a 32-key warp sort (~50 instructions, about one
              count step) that the numbers below do not include.
  est         the secant rule without the first count: the count at the lower end of the bracket is a rank statistic
              (distribution-free for iid rows), so a constant stands in for it
  warm        est + the previous row's threshold as the first pivot (helps when consecutive rows have similar scale)
"""
import numpy as np


def keys_of(v):
    b = (v.astype(np.float32) + np.float32(0)).view(np.uint32).astype(np.uint64)
    return np.where((b >> 31) & 1 == 1, (~b) & 0xffffffff, b | 0x80000000).astype(np.int64)


def recip31(d):
    return (1 << 31) // d


def secant(key, k, lo, hi, c_lo, c_hi, n):
    it = 0
    while hi - lo > 1:
        span = hi - lo
        it += 1
        sec = min(max((span * (((2 * (c_lo - k) + 1) * recip31(c_lo - c_hi)) & 0xffffffff)) >> 32, 1), span - 1)
        mid = lo + ((span >> 1) if it % 4 == 0 else sec)
        c = int((key >= mid).sum())
        n += 1
        if c == k:
            return n
        if c > k:
            lo, c_lo = mid, c
        else:
            hi, c_hi = mid, c
    return n


def lane_stats(key):
    srt = -np.sort(-key.reshape(32, 8), axis=1)
    return srt[:, 0], srt[:, 1]


def rule_secant(key, k, use_m2=False):
    m1, m2 = lane_stats(key)
    lo, hi = int(m1.min()) if k <= 32 else int(m2.min()), int(m1.max()) + 1
    c_lo, c_hi = int((key >= lo).sum()), 0
    if use_m2:
        p2, c2 = int(m2.max()) + 1, int((m1 > m2.max()).sum())
        if lo < p2 < hi:
            if c2 >= k:
                lo, c_lo = p2, c2
            else:
                hi, c_hi = p2, c2
    return 1 if c_lo == k else secant(key, k, lo, hi, c_lo, c_hi, 1)


RANK = {8: 7, 16: 13, 32: 20, 64: 19}       # expected rank among the sorted lane maxima (second maxima for k = 64)


def rule_lane_rank(key, k, slope=1.5):
    m1, m2 = lane_stats(key)
    s = -np.sort(-(m1 if k <= 32 else m2))
    lo, hi, c_lo, c_hi = int(m1.min()) if k <= 32 else int(m2.min()), int(m1.max()) + 1, None, 0
    i = RANK[k]
    p = int(s[i])
    c = int((key >= p).sum())
    n = 1
    if c == k:
        return n
    if c > k:
        lo, c_lo = p, c
        j = max(i - max(1, int(round((c - k + 0.5) / slope))), 0)
    else:
        hi, c_hi = p, c
        j = min(i + max(1, int(round((k - c + 0.5) / slope))), 31)
    p2 = int(s[j])
    if lo < p2 < hi:
        c2 = int((key >= p2).sum())
        n += 1
        if c2 == k:
            return n
        if c2 > k:
            lo, c_lo = p2, c2
        else:
            hi, c_hi = p2, c2
    if c_lo is None:
        c_lo = int((key >= lo).sum())
        n += 1
        if c_lo == k:
            return n
    return secant(key, k, lo, hi, c_lo, c_hi, n)


def secant_t(key, k, lo, hi, c_lo, c_hi, n):
    """secant() that also returns the threshold found (and tolerates an estimated c_lo)."""
    it = 0
    while hi - lo > 1:
        span = hi - lo
        it += 1
        d = max(c_lo - c_hi, 1)
        sec = min(max((span * (((2 * (c_lo - k) + 1) * recip31(d)) & 0xffffffff)) >> 32, 1), span - 1)
        mid = lo + ((span >> 1) if it % 4 == 0 else sec)
        c = int((key >= mid).sum())
        n += 1
        if c == k:
            return n, mid
        if c > k:
            lo, c_lo = mid, c
        else:
            hi, c_hi = mid, c
    return n, lo


EST = {8: 96, 16: 97, 32: 99, 64: 138}     # median count at the lower end of the bracket, iid rows of 256


def rule_warm(key, k, t_prev=None):
    m1, m2 = lane_stats(key)
    lo, hi = (int(m1.min()) if k <= 32 else int(m2.min())), int(m1.max()) + 1
    c_lo, c_hi, n = max(EST[k], k + 1), 0, 0
    if t_prev is not None and lo < t_prev < hi:
        c = int((key >= t_prev).sum())
        n = 1
        if c == k:
            return n, t_prev
        if c > k:
            lo, c_lo = t_prev, c
        else:
            hi, c_hi = t_prev, c
    return secant_t(key, k, lo, hi, c_lo, c_hi, n)


def main():
    rng = np.random.default_rng(0)
    dists = (("U[0,1)", lambda: rng.random(256, dtype=np.float32)),
             ("N(0,1)", lambda: rng.standard_normal(256).astype(np.float32)),
             ("relu(N(0,1))", lambda: np.maximum(rng.standard_normal(256), 0).astype(np.float32)),
             ("lognormal(0,2)", lambda: np.exp(2 * rng.standard_normal(256)).astype(np.float32)),
             ("N(0,1) x row scale", lambda: (np.exp(1.5 * rng.standard_normal()) * rng.standard_normal(256)).astype(np.float32)))
    print("%-18s %4s %8s %8s %10s %8s %8s" % ("distribution", "k", "secant", "+m2", "lane-rank", "est", "warm"))
    for name, gen in dists:
        for k in (8, 16, 32, 64):
            rows = [keys_of(gen()) for _ in range(300)]
            a = np.mean([rule_secant(r, k) for r in rows])
            b = np.mean([rule_secant(r, k, True) for r in rows])
            c = np.mean([rule_lane_rank(r, k) for r in rows])
            e = np.mean([rule_warm(r, k)[0] for r in rows])
            w, t_prev = [], None
            for r in rows:
                n, t_prev = rule_warm(r, k, t_prev)
                w.append(n)
            print("%-18s %4d %8.2f %8.2f %10.2f %8.2f %8.2f" % (name, k, a, b, c, e, np.mean(w)))


if __name__ == "__main__":
    main()
