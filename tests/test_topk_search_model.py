"""CPU: the threshold search of csrc/topk.cu (find_threshold) restated with the same integer arithmetic and run
on adversarial rows.  It cannot test the kernel (the -m gpu parity tests do that); it pins the ALGORITHM: the
bracket invariants, that every returned threshold T satisfies count(> T) <= k <= count(>= T), and the bound on
the number of count steps that the "plain bisection every fourth step" rule guarantees."""
import numpy as np
import pytest


def keys_of(v):
    """fast_key(): order-preserving uint32 image, -0 == +0, every NaN the largest key."""
    v = np.asarray(v, np.float32) + np.float32(0)
    b = v.view(np.uint32).astype(np.uint64)
    b = np.where(np.isnan(v), np.uint64(0x7fffffff), b)
    return np.where((b >> np.uint64(31)) & np.uint64(1) == 1, (~b) & np.uint64(0xffffffff), b | np.uint64(0x80000000)).astype(np.int64)


def find_threshold(key, k, p2=0, c2=0):
    """Mirror of find_threshold() in csrc/topk.cu.  Returns (T, exact, number of counts)."""
    lanes = key.reshape(32, 8)
    srt = -np.sort(-lanes, axis=1)
    m1, m2 = srt[:, 0], srt[:, 1]
    kmax = int(m1.max())
    kmin = int(m1.min()) if k <= 32 else (int(m2.min()) if k <= 64 else int(key.min()))
    steps = 0

    def count(t):
        nonlocal steps
        steps += 1
        return int((key >= t).sum())

    if kmax == 0xffffffff and count(kmax) >= k:
        return kmax, False, steps
    lo, hi = kmin, (kmax if kmax == 0xffffffff else kmax + 1)
    c_lo = count(lo)
    c_hi = count(kmax) if kmax == 0xffffffff else 0
    if lo < p2 < hi:
        if c2 >= k:
            lo, c_lo = p2, c2
        else:
            hi, c_hi = p2, c2
    assert c_lo >= k > c_hi
    span = 0 if c_lo == k else hi - lo
    it = 1
    while span > 1:
        num = 2 * (c_lo - k) + 1
        frac = (num * ((1 << 31) // (c_lo - c_hi))) & 0xffffffff
        assert num * ((1 << 31) // (c_lo - c_hi)) < (1 << 32)            # the Q32 fraction never overflows
        sec = min(max((span * frac) >> 32, 1), span - 1)
        mid = lo + ((span >> 1) if (it & 3) == 0 else sec)              # the kernel unrolls the step four times
        assert lo < mid < hi
        c = count(mid)
        if c >= k:
            lo, c_lo = mid, c
        if c <= k:                                                       # c == k collapses the bracket: lo = hi = mid
            hi, c_hi = mid, c
        span = hi - lo
        it += 1
    exact = c_lo == k
    return lo, exact, steps


def _check(row, k):
    key = keys_of(row)
    T, exact, steps = find_threshold(key, k)
    gt, ge = int((key > T).sum()), int((key >= T).sum())
    assert gt <= k <= ge, (k, gt, ge)
    assert exact == (ge == k) or not exact            # `exact` is only ever claimed when it holds
    if exact:
        assert ge == k
    assert steps <= 2 + 4 * 33                        # one bisection every 4 steps halves a span < 2^32
    return steps


@pytest.mark.parametrize("k", [1, 8, 16, 32, 33, 64, 65, 128, 255])
def test_search_returns_a_valid_threshold_on_adversarial_rows(k):
    rng = np.random.default_rng(k)
    rows = [
        np.full(256, 5.0), np.zeros(256), np.tile([0.0, -0.0], 128), np.arange(256, dtype=np.float32),
        np.arange(256, dtype=np.float32)[::-1].copy(), np.repeat(np.arange(8, dtype=np.float32), 32),
        np.where(np.arange(256) % 50 == 0, np.inf, -np.inf), np.where(np.arange(256) % 3 == 0, np.nan, 1.0),
        np.concatenate([np.full(40, np.nan), rng.standard_normal(216)]),
        np.concatenate([[3.0e38, -3.0e38], rng.standard_normal(254) * 1e-30]),          # 2^32-wide key range
        np.float32(1.0) + np.arange(256, dtype=np.float32) * np.float32(2.0 ** -23),       # adjacent floats
        np.exp(rng.standard_normal(256) * 20).astype(np.float32),                           # 17 decades
        -np.exp(rng.standard_normal(256) * 20).astype(np.float32),
        np.round(rng.standard_normal(256) * 2) / 2,                                         # heavy ties
        np.concatenate([np.full(k, 7.0), np.full(256 - k, 7.0 - 2.0 ** -20)]) if k < 256 else np.full(256, 7.0),
    ]
    for row in rows:
        _check(np.asarray(row, np.float32), k)
    worst = max(_check(rng.standard_normal(256).astype(np.float32) * rng.choice([1e-20, 1.0, 1e20]), k) for _ in range(200))
    assert worst <= 40                                # random data: nowhere near the worst-case bound


def test_second_point_from_the_lane_maxima_keeps_the_bracket_valid():
    """(M2 + 1, #lane maxima > M2), the free point the kernel uses for k <= 16: it is an exact value of the count
    function, so handing it to the search cannot change the result, only the number of steps."""
    rng = np.random.default_rng(3)
    saved = 0
    for _ in range(300):
        row = rng.standard_normal(256).astype(np.float32)
        key = keys_of(row)
        lanes = -np.sort(-key.reshape(32, 8), axis=1)
        big2 = int(lanes[:, 1].max())
        c2 = int((lanes[:, 0] > big2).sum())
        assert c2 == int((key >= big2 + 1).sum())
        for k in (8, 16):
            t0, _, s0 = find_threshold(key, k)
            t1, _, s1 = find_threshold(key, k, big2 + 1, c2)
            assert set(np.nonzero(key >= t0)[0]) == set(np.nonzero(key >= t1)[0])
            saved += s0 - s1
    assert saved > 0


def _popc(x):
    return bin(x).count("1")


def _mem_pos(p, k):
    """banked_mem_pos of csrc/topk.cu: sorted rank -> position in the CBSR row."""
    if k < 32:
        return p
    epl = k // 4
    t, i = divmod(p, epl)
    return 32 * (i // 8) + 8 * t + i % 8


def _banked_reference(cols, k):
    """MAXK_ORDER_BANKED by definition: classes column mod 4 by (size desc, class asc), columns ascending inside a
    class, rank p stored at _mem_pos(p)."""
    size = [sum(1 for c in cols if c % 4 == u) for u in range(4)]
    rank = [sum(1 for v in range(4) if v != u and (size[v] > size[u] or (size[v] == size[u] and v < u))) for u in range(4)]
    ordered = sorted(cols, key=lambda c: (rank[c % 4], c))
    out = [None] * k
    for p, c in enumerate(ordered):
        out[_mem_pos(p, k)] = c
    return out


def _banked_positions(sel_cols, k):
    """Mirror of the position computation of topk_banked_kernel<K>: lane l owns columns 8l .. 8l+7, bs[s] is
    the ballot of slot s; returns the column stored at every output position."""
    selb = np.zeros((32, 8), bool)
    for c in sel_cols:
        selb[c // 8, c % 8] = True
    bs = [sum(1 << lane for lane in range(32) if selb[lane, s]) for s in range(8)]
    sz = [_popc(bs[u]) + _popc(bs[u + 4]) for u in range(4)]
    out = {}
    for lane in range(32):
        lt = (1 << lane) - 1
        pos = [0] * 8
        for u in range(4):               # class == slot & 3; inside a class: (lane, slot < 4 first)
            base = sum(sz[v] for v in range(4) if v != u and (sz[v] > sz[u] or (sz[v] == sz[u] and v < u)))
            p = base + _popc(bs[u] & lt) + _popc(bs[u + 4] & lt)
            pos[u] = _mem_pos(p, k)
            pos[u + 4] = _mem_pos(p + (1 if selb[lane, u] else 0), k)
        for s in range(8):
            if selb[lane, s]:
                assert pos[s] not in out
                out[pos[s]] = 8 * lane + s
    return [out[i] for i in range(k)]


@pytest.mark.parametrize("k", [8, 16, 32, 64, 96, 128])
def test_banked_position_formulas_give_the_banked_order(k):
    """The ballot / popcount formulas of the specialised kernel produce exactly the MAXK_ORDER_BANKED permutation
    for any selected set, and it is the order the oracle emits."""
    rng = np.random.default_rng(k)
    for t in range(400):
        pool = 256 if t % 4 else max(k, 64)          # every fourth case: columns clustered in the first lanes
        cols = rng.choice(pool, k, replace=False).tolist()
        assert _banked_positions(cols, k) == _banked_reference(cols, k)
    import oracle
    x = rng.standard_normal((50, 256)).astype(np.float32)
    _, oc = oracle.topk(x, k, 2)
    for r in range(50):
        assert _banked_positions(oc[r].tolist(), k) == oc[r].tolist()
