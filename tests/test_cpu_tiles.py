"""CPU: the arithmetic of the photon kernel's tile look-up, restated in NumPy.

`pyxsim_b200/csrc/pipeline.cu` numbers photon PAIRS cell by cell (``qoff``) and hands 128-slot
tiles to warps.  Inside a tile the cell of a slot is found without a search: the starts of the
tile's cells are staged relative to the tile's first slot (the first cell's true, negative start
included), the rank of a slot's cell is "cells starting at or before it, minus one" -- a ballot
over the cells starting before the round's 32 slots plus an OR of start bits inside them -- and
the pair index inside the cell is ``slot - start[rank]`` for every rank.  This file checks that
arithmetic (the ballot / popcount formulation, the packing of four rank bytes in one word, the
``mark_tiles`` map from tiles to first cells) against a plain bisection of ``qoff`` on random and
adversarial layouts.  The kernel itself is compared with the oracle in the GPU tests."""
import numpy as np
import pytest

TSLOTS, ROUNDS, MAXC = 128, 4, 132
REL_SEARCH = 0xFFFFFFFF


def mark_tiles(qoff):
    """tile -> cell holding the tile's first slot (k_prep_*: every tile that STARTS inside a
    cell's slot range is marked with that cell)"""
    n_pairs = int(qoff[-1])
    ntiles = (n_pairs + TSLOTS - 1) // TSLOTS
    tile_cell = np.full(ntiles, -1, dtype=np.int64)
    for k in range(len(qoff) - 1):
        qa, qb = int(qoff[k]), int(qoff[k + 1])
        t = (qa + TSLOTS - 1) // TSLOTS
        while t * TSLOTS < qb:
            tile_cell[t] = k
            t += 1
    return tile_cell


def popc(x):
    return bin(x & 0xFFFFFFFF).count("1")


def tile_lookup(qoff, tile_cell, tile):
    """(rel, m) of every slot of the tile the way `decide` computes them, or None when the tile
    falls back to the global bisection"""
    n_cells = len(qoff) - 1
    slot0 = tile * TSLOTS
    c0 = int(tile_cell[tile])
    c1 = int(tile_cell[tile + 1]) if tile + 1 < len(tile_cell) else n_cells - 1
    ncl = c1 - c0 + 1
    if ncl > MAXC:
        return None
    starts = [int(qoff[c0 + q]) - slot0 for q in range(ncl)]
    ends = [int(qoff[c0 + q + 1]) - slot0 for q in range(ncl)]
    if any(a == b for a, b in zip(starts, ends)) or any(a < -(1 << 30) for a in starts):
        return None
    cnt = [0] * 32                                       # one packed word per lane
    for qb in range(0, ncl, 32):
        qs = [starts[qb + l] if qb + l < ncl else 0x7FFFFFFF for l in range(32)]
        for r in range(ROUNDS):
            before = sum(1 << l for l in range(32) if qs[l] < r * 32)
            mask = 0
            for l in range(32):
                d = (qs[l] - r * 32) & 0xFFFFFFFF
                if d < 32:
                    mask |= 1 << d
            for lane in range(32):
                byte = popc(before) + popc(mask & ((2 << lane) - 1))
                cnt[lane] = (cnt[lane] + (byte << (8 * r))) & 0xFFFFFFFF
    out = []
    for r in range(ROUNDS):
        for lane in range(32):
            relpack = (cnt[lane] - 0x01010101) & 0xFFFFFFFF
            assert relpack != REL_SEARCH
            rel = (relpack >> (8 * r)) & 255
            srel = r * 32 + lane
            out.append((rel, srel - starts[rel]))
    return out


def reference_lookup(qoff, slot):
    k = int(np.searchsorted(qoff, slot, side="right")) - 1
    return k, slot - int(qoff[k])


def layouts():
    rng = np.random.RandomState(5)
    yield "ones", np.ones(700, dtype=np.int64)                      # every slot its own cell
    yield "big", np.array([5, 3000, 1, 1, 700, 2], dtype=np.int64)  # cells spanning many tiles
    yield "mixed", rng.randint(1, 90, size=400).astype(np.int64)
    yield "small", rng.randint(1, 4, size=2000).astype(np.int64)
    yield "exact", np.full(37, TSLOTS, dtype=np.int64)              # cells ending on tile boundaries
    yield "empties", (rng.randint(0, 3, size=3000) * rng.randint(0, 2, size=3000)).astype(np.int64)


@pytest.mark.parametrize("name,npairs", list(layouts()), ids=[n for n, _ in layouts()])
def test_tile_lookup_matches_bisection(name, npairs):
    qoff = np.concatenate([[0], np.cumsum(npairs)])
    tile_cell = mark_tiles(qoff)
    n_pairs = int(qoff[-1])
    assert (tile_cell >= 0).all()
    fast = 0
    for tile in range(len(tile_cell)):
        c0 = int(tile_cell[tile])
        assert qoff[c0] <= tile * TSLOTS < qoff[c0 + 1]             # first slot lies in a NON-EMPTY cell
        got = tile_lookup(qoff, tile_cell, tile)
        if got is None:
            assert name == "empties"                                 # only lists with empty cells search
            continue
        fast += 1
        for srel, (rel, m) in enumerate(got):
            slot = tile * TSLOTS + srel
            if slot >= n_pairs:
                continue
            k, mref = reference_lookup(qoff, slot)
            assert (c0 + rel, m) == (k, mref), (name, tile, srel)
    assert fast > 0 or name == "empties"


def test_staged_cells_fit_and_rank_bytes_never_collide_with_the_search_marker():
    # a tile of 128 slots holds at most 128 non-empty cells plus the one of the next tile's first slot
    assert TSLOTS + 1 <= MAXC < 255
    qoff = np.arange(0, 1000, dtype=np.int64)
    tile_cell = mark_tiles(qoff)
    for tile in range(len(tile_cell) - 1):
        assert tile_cell[tile + 1] - tile_cell[tile] + 1 <= MAXC
