"""CPU study (numpy): what a 128-row output tile of a 3x3x3 submanifold convolution touches under different row orders, on the
synthetic 300 k-point frame of bench.py (159.9 k voxels at 0.2 m, 5.57 present neighbours per voxel).

    python tools/tile_reuse_study.py

For each order: unique input rows per tile (what a tile-local cache of inputs would hold), the reuse factor against the
(row, offset) pairs, and the number of offsets with at least one present neighbour (= MMA stages per K chunk, each a full 128-row
tile whatever its density).  Result: the shipped popcount-mask order halves the MMA stages (11 against 21-24)
but leaves no input reuse (x1.2); spatial orders reuse inputs x3.4-3.9 and double the stages."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fullysparsefusion_b200 import synth  # noqa: E402
from oracle import fsf_oracle as O  # noqa: E402  (tools may use the checker; the product path never does)


def main():
    pts = synth.ring_points(300000, sweeps=10, seed=0)
    uniq, _, _ = O.unique_rows(O.voxelize(pts, synth.NUSC_VOXEL, synth.NUSC_RANGE, floor_mode=1))
    M = len(uniq)
    z, y, x = (uniq[:, i].astype(np.int64) for i in range(3))
    key = (z * 514 + y + 1) * 514 + x + 1 + 514 * 514
    order = np.argsort(key)
    skey = key[order]
    nbr = np.full((27, M), -1, np.int64)
    k = 0
    for dz in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                q = key + (dz * 514 + dy) * 514 + dx
                pos = np.minimum(np.searchsorted(skey, q), M - 1)
                hit = skey[pos] == q
                nbr[k, hit] = order[pos[hit]]
                k += 1
    pairs = int((nbr >= 0).sum())
    print(f"{M} voxels, {pairs} (row, offset) pairs = {pairs / M:.2f} per voxel")

    def stats(perm, name):
        nb = nbr[:, perm]
        tiles = (M + 127) // 128
        uniq_rows = stages = 0
        for t in range(tiles):
            blk = nb[:, t * 128:(t + 1) * 128]
            uniq_rows += len(np.unique(blk[blk >= 0]))
            stages += int((blk >= 0).any(1).sum())
        print(f"{name:28s} unique input rows/tile {uniq_rows / tiles:6.1f}   pairs/tile {pairs / tiles:6.1f}   reuse x{pairs / uniq_rows:4.2f}"
              f"   active offsets/tile {stages / tiles:5.2f}   rows present per stage {pairs / stages:5.1f} of 128")

    def spread(v):
        v = v.astype(np.uint64)
        r = np.zeros_like(v)
        for i in range(10):
            r |= ((v >> np.uint64(i)) & np.uint64(1)) << np.uint64(3 * i)
        return r

    stats(np.arange(M), "ranked (z,y,x) order")
    stats(np.argsort(spread(x) | (spread(y) << np.uint64(1)) | (spread(z) << np.uint64(2)), kind="stable"), "Morton order")
    mask = np.zeros(M, np.uint32)
    for k in range(27):
        mask |= (nbr[k] >= 0).astype(np.uint32) << np.uint32(k)
    pc = np.array([bin(int(m)).count("1") for m in mask])
    stats(np.lexsort((mask, pc)), "popcount-mask order (shipped)")
    stats(np.lexsort((x, y, z, x // 8, y // 8)), "8x8 BEV blocks")


if __name__ == "__main__":
    main()
