"""Offline model (numpy, no GPU) of the merge loop's lane utilisation on the bench data under
different ways of cutting a tile into per-thread slices.  The merge loop runs one iteration
per DISTINCT end in a thread's slice; a warp takes as long as its busiest lane.  Printed:
mean iterations per lane / mean over warps of the busiest lane (= active-lane fraction)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench

T, TILE = 512, 5461
n, r = 10_000_000, 100_000
ops = [bench.gen_operand(2000 + j, n, r)[0].astype(np.int64) for j in range(4)]
merged = np.unique(np.concatenate(ops))              # distinct ends = merge iterations
n_tiles = len(merged) // TILE


def efficiency(cuts_of_tile):
    eff = []
    for t in range(n_tiles):
        ends = merged[t * TILE:(t + 1) * TILE]
        cuts = cuts_of_tile(ends)                    # T + 1 boundaries (coordinates), cuts[0] < ends[0]
        per_thread = np.histogram(ends, bins=cuts)[0]
        assert per_thread.sum() == len(ends)
        busiest = per_thread.reshape(T // 32, 32).max(axis=1)
        eff.append(per_thread.mean() / busiest.mean())
    return float(np.mean(eff))


def uniform_coordinate(ends):                        # the kernel today
    return np.linspace(ends[0] - 1, ends[-1], T + 1)


def by_merged_rank(ends):                            # exact merge-path balance (upper bound)
    idx = np.linspace(0, len(ends), T + 1).astype(int)
    c = np.concatenate([[ends[0] - 1], ends[np.minimum(idx[1:], len(ends)) - 1]])
    c[-1] = ends[-1]
    return c


def by_longest_source(ends, src=0):                  # every g-th run of ONE source, others follow
    own = ops[src][(ops[src] >= ends[0]) & (ops[src] <= ends[-1])]
    if len(own) < T:
        return uniform_coordinate(ends)
    idx = np.linspace(0, len(own), T + 1).astype(int)
    c = np.concatenate([[ends[0] - 1], own[np.minimum(idx[1:], len(own)) - 1]])
    c[-1] = ends[-1]
    return np.maximum.accumulate(c)


def uniform_in_sub_tiles(ends, parts=16):            # coordinate split inside 16 equal-count sub-tiles
    cuts = [ends[0] - 1]
    per = T // parts
    for q in range(parts):
        seg = ends[q * len(ends) // parts:(q + 1) * len(ends) // parts]
        lo = cuts[-1]
        cuts.extend(np.linspace(lo, seg[-1], per + 1)[1:])
    return np.asarray(cuts, dtype=float)


for name, fn in [("uniform coordinate split (current)", uniform_coordinate),
                 ("16 equal-count sub-tiles, uniform inside", uniform_in_sub_tiles),
                 ("every g-th run of source 0", by_longest_source),
                 ("equal merged rank (merge path)", by_merged_rank)]:
    print(f"{name:45s} active-lane fraction {efficiency(fn):.3f}")
