"""Developer tool: shared-memory wavefronts per edge of the forward accumulate (and backward look-up) for a CBSR
entry order, under the conflict model measured on B200: one LDS/STS warp instruction costs
as many wavefronts as its worst bank multiplicity.  Pure numpy, no GPU.

    python tools/bank_conflict_sim.py

Layout of Lay<K> (csrc/maxk_common.cuh): a lane owns EPL consecutive entries of one edge, L = k / EPL lanes cover
an edge, EPI = 32 / L edges are processed per warp instruction, every edge slot has its own copy of the 256
columns in banks [q L, q L + L): inside a slot, instruction i touches the entries {EPL t + i : t < L}, and two of
them collide when their columns are equal mod L.  The instruction costs the worst slot.
"""
import numpy as np


def wavefronts_per_edge(rows, k):
    epl = 2 if k == 8 else 4
    lanes = k // epl
    epi = 32 // lanes
    n = (len(rows) // epi) * epi
    res = rows[:n] % lanes                                   # bank residue of every entry
    res = res.reshape(n, lanes, epl)                          # [edge, lane t, entry i of the lane]
    mult = np.zeros((n, epl), np.int64)
    for i in range(epl):
        counts = np.zeros((n, lanes), np.int64)
        np.add.at(counts, (np.repeat(np.arange(n), lanes), res[:, :, i].reshape(-1)), 1)
        mult[:, i] = counts.max(axis=1)                       # worst bank of this slot in instruction i
    per_group = mult.reshape(n // epi, epi, epl).max(axis=1).sum(axis=1)   # worst slot per instruction, summed
    return 2.0 * per_group.mean() / epi                       # read + write, per edge


def main():
    rng = np.random.default_rng(0)
    print("%4s %28s %10s %10s %10s" % ("k", "layout", "no-conflict", "value order", "banked"))
    for k, m in ((8, 4), (16, 4), (32, 8), (64, 16)):
        cols = np.stack([rng.choice(256, k, replace=False) for _ in range(20000)])
        banked = np.stack([np.array(sorted(r, key=lambda c: (c % m, c))) for r in cols])
        epl = 2 if k == 8 else 4
        lanes = k // epl
        print("%4d %28s %10.2f %10.2f %10.2f" % (k, "%d slots x %d banks, %d entries/lane" % (32 // lanes, lanes, epl),
                                                  2.0 * epl / (32 // lanes), wavefronts_per_edge(cols, k),
                                                  wavefronts_per_edge(banked, k)))


if __name__ == "__main__":
    main()
