"""TEST INFRASTRUCTURE (oracle): CPU restatement of the expert-parallel exchange of csrc/snb_ep.cu.

Reference semantics being restated (tutel_moe_layer_nobatch.py:155-218): every rank routes its
own chunk (capacity from the local S and the global E), kept rows travel to the rank that owns their expert
(expert e on rank e // E_local), come back, and dropped rows never leave.  Because capacity/drops are decided
per source rank and the expert MLP is row-wise, the result on every rank equals the all-local result.

This module simulates the slot algebra of the CUDA kernels (record slots, per-source kept counts, the owner's
tile plan) in numpy so the protocol itself can be checked without GPUs: tests/test_host_logic.py."""
import numpy as np

TILE = 128


def dispatch(rank, world, E, capmax, idx, loc, cap):
    """k_ep_dispatch of one source rank.  Returns {owner: [(slot, sample)]}, the dropped list [(slot, sample)]
    (slots inside the source's own region) and kept counts kc[e]."""
    EL = E // world
    out = {w: [] for w in range(world)}
    dropped = []
    for s, (e, l) in enumerate(zip(idx.tolist(), loc.tolist())):
        if l < cap:
            owner, el = divmod(e, EL)
            out[owner].append(((rank * EL + el) * capmax + l, s))
        else:
            dropped.append((world * EL * capmax + len(dropped), s))
    counts = np.bincount(idx, minlength=E)
    return out, dropped, np.minimum(counts, cap)


def plan(rank, world, E, capmax, kc_by_source, n_dropped):
    """k_ep_plan of one owner.  kc_by_source[w][el] = kept rows source w sent for local expert el.
    Returns tiles [(global expert or -1, row0, rows)] and row2slot (-1 = padding)."""
    EL = E // world
    tiles, rows = [], []
    for el in range(EL + 1):
        if el < EL:
            slots = [(w * EL + el) * capmax + l for w in range(world) for l in range(int(kc_by_source[w][el]))]
            expert = rank * EL + el
        else:
            slots = [world * EL * capmax + j for j in range(n_dropped)]
            expert = -1
        row0 = len(rows)
        for i in range(0, len(slots), TILE):
            tiles.append((expert, row0 + i, min(TILE, len(slots) - i)))
        pad = (-len(slots)) % TILE
        rows.extend(slots + [-1] * pad)
    return tiles, np.array(rows, dtype=np.int64)


def exchange(world, E, capmax, routing, row_fn, dropped_fn):
    """Full round trip.  routing[r] = (idx, loc, cap, payload[S, ...]).  row_fn(expert, payload_row) evaluates a
    kept row on its owner, dropped_fn(payload_row) a dropped row at home.  Returns ret[r][s]."""
    EL = E // world
    rx = [dict() for _ in range(world)]            # slot -> (source rank, sample, payload row)
    kc = [[None] * world for _ in range(world)]    # kc[owner][source] = counts of the owner's local experts
    ndrop = [0] * world
    for r, (idx, loc, cap, payload) in enumerate(routing):
        sent, dropped, kept = dispatch(r, world, E, capmax, idx, loc, cap)
        for owner, items in sent.items():
            for slot, s in items:
                assert slot not in rx[owner], "two records in one slot"
                rx[owner][slot] = (r, s, payload[s])
        for slot, s in dropped:
            rx[r][slot] = (r, s, payload[s])
        ndrop[r] = len(dropped)
        for owner in range(world):
            kc[owner][r] = kept[owner * EL:(owner + 1) * EL]
    ret = [dict() for _ in range(world)]
    for owner in range(world):
        tiles, row2slot = plan(owner, world, E, capmax, kc[owner], ndrop[owner])
        seen = set()
        for expert, row0, nrows in tiles:
            for slot in row2slot[row0:row0 + nrows].tolist():
                assert slot >= 0 and slot not in seen
                seen.add(slot)
                src, s, row = rx[owner][slot]
                assert s not in ret[src], "a sample evaluated twice"
                ret[src][s] = row_fn(expert, row) if expert >= 0 else dropped_fn(row)
        assert seen == set(rx[owner]), "records the plan never visits"
    return ret
