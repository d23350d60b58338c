"""Developer tool: shared-memory wavefronts per edge of the forward accumulate for a CBSR entry order and a
slot layout, under the conflict model measured on B200: one LDS/STS warp instruction costs as
many wavefronts as its worst bank multiplicity.  Pure numpy, no GPU.

    python tools/bank_conflict_sim.py

Layout (csrc/slots.cuh): L lanes cover an edge, a lane owns EPL = k / L entries, S = 32 / L edge slots per warp
instruction, every slot has its own copy of the 256 columns in banks [q L, q L + L).  Inside a slot, instruction i
touches the i-th entry of each of the L lanes, and two of them collide when their columns are equal mod L; the
instruction costs the worst slot.  Orders: "random" (the selection in arbitrary order), "plain" = sorted by
(column mod L, column), "by size" = residue classes with the largest class first (MAXK_ORDER_BANKED for L = 4): a
class with more than EPL entries necessarily puts two of them into some instruction; starting with the largest class
makes those doubled instructions the SAME (first) ones in all slots, so the worst-slot cost is paid once.
"""
import numpy as np


def wavefronts_per_edge(rows, lanes):
    k = rows.shape[1]
    epl = k // lanes
    slots = 32 // lanes
    n = (len(rows) // slots) * slots
    res = (rows[:n] % lanes).reshape(n, lanes, epl)             # [edge, lane t, entry i of the lane]
    mult = np.zeros((n, epl), np.int64)
    for i in range(epl):
        counts = np.zeros((n, lanes), np.int64)
        np.add.at(counts, (np.repeat(np.arange(n), lanes), res[:, :, i].reshape(-1)), 1)
        mult[:, i] = counts.max(axis=1)                        # worst bank of this slot in instruction i
    per_group = mult.reshape(n // slots, slots, epl).max(axis=1).sum(axis=1)   # worst slot per instruction, summed
    return 2.0 * per_group.mean() / slots                      # read + write, per edge


def by_size(row, m):
    cls = [[c for c in sorted(row) if c % m == j] for j in range(m)]
    cls.sort(key=lambda l: -len(l))                            # stable: ties keep the lower class first
    return np.array([c for l in cls for c in l])


def epilogue_wavefronts(lane_map):
    """Wavefronts of one 16-byte access of the per-row epilogue: a quarter-warp (8 consecutive lanes) is served
    per pass, a pass costs its worst bank multiplicity.  lane_map(lane) -> (copy q, word row j): the float4 at
    words 32 j + 4 q .. + 3, i.e. banks [4 q, 4 q + 4)."""
    total = 0
    for quarter in range(4):
        hits = {}
        for lane in range(8 * quarter, 8 * quarter + 8):
            q, j = lane_map(lane)
            for b in range(4 * q, 4 * q + 4):
                hits.setdefault(b, set()).add(32 * j + b)
        total += max(len(v) for v in hits.values())
    return total


def main():
    # the epilogue's two lane maps (csrc/spgemm_fwd.cu): the accumulation's own (4 lanes per copy) against lane l on
    # copy l % 8 -- ncu measured 16 and 4 wavefronts per LDS.128 / STS.128 (profiles/r02_ncu_full_fwd_yelp_k32_epilogue_fix.txt)
    print("epilogue, wavefronts per 16-byte access: lanes (q, t) = (l / 4, l %% 4): %d   lanes (q, t) = (l %% 8, l / 8): %d"
          % (epilogue_wavefronts(lambda l: (l // 4, 2 * (l % 4))), epilogue_wavefronts(lambda l: (l % 8, 2 * (l // 8)))))
    rng = np.random.default_rng(0)
    print("%4s %22s %8s %8s %8s %8s" % ("k", "layout", "floor", "random", "plain", "by size"))
    for k in (8, 16, 32, 64, 128):
        cols = np.stack([rng.choice(256, k, replace=False) for _ in range(8000)])
        for lanes in (4, 8, 16):
            if k // lanes < 1 or (k == 8 and lanes == 16):
                continue
            plain = np.stack([np.array(sorted(r, key=lambda c: (c % lanes, c))) for r in cols])
            sized = np.stack([by_size(r, lanes) for r in cols])
            print("%4d %22s %8.2f %8.2f %8.2f %8.2f" % (k, "%d slots x %d banks" % (32 // lanes, lanes), 2.0 * k / 32,
                                                        wavefronts_per_edge(cols, lanes), wavefronts_per_edge(plain, lanes),
                                                        wavefronts_per_edge(sized, lanes)))


if __name__ == "__main__":
    main()
