"""Executable model (numpy, CPU) of the flagged segmented-scan sweep of lux_b200/csrc/seg.cuh and of the chained fix-up of
pull.cuh (pull_fixup_fused_kernel), checked against per-vertex sums computed directly from the CSC:

  * stream layout: one word per edge, head flag on the first in-edge of every non-empty vertex, every block padded with
    1 .. stage head-flagged dummy words to whole stages; close_vtx[j] = the vertex completed by head j (entry 0 and the
    entries of pad heads are dummies); tile_v[t] = heads before piece t                         (build_seg_stream, api.cu)
  * one piece: rounds of R edges; per round the completed sums, the carry across rounds, the piece's head partial (the
    sum in front of its first head, deferred to the fix-up) and tail partial                     (seg_tile_kernel)
  * fix-up: chained scan over pieces with decoupled look-back — blocks of 256 pieces publish (flag, value) aggregates,
    a block walks back over its predecessors until it meets an aggregate that contains a head  (pull_fixup_fused_kernel)
  * vertices without in-edges get update(identity) separately                                   (empties_kernel)
Integer edge values make every summation order exact, so the comparison is bit-exact."""
import numpy as np
import pytest

DUMMY = 0xFFFFFFFF
FIX_BLOCK = 256


def build_stream(row_end, vals, stage, blocks=None):
    """blocks: list of (v_lo, v_hi) vertex ranges (panel: one per source block); default one block."""
    nv = len(row_end)
    starts = np.concatenate([[0], row_end[:-1]]).astype(np.int64)
    blocks = blocks or [(0, nv)]
    words, heads, own = [], [], []  # per word: value, head flag; per head: owning vertex
    for (lo, hi) in blocks:
        e0, e1 = (starts[lo] if lo < nv else int(row_end[-1])), (int(row_end[hi - 1]) if hi > lo else (starts[lo] if lo < nv else 0))
        for v in range(lo, hi):
            for k, e in enumerate(range(starts[v], int(row_end[v]))):
                words.append(vals[e])
                heads.append(1 if k == 0 else 0)
                if k == 0:
                    own.append(v)
        n_edges = e1 - e0
        pad = stage - n_edges % stage  # 1 .. stage: the first pad closes the block's last vertex
        for _ in range(pad):
            words.append(0)
            heads.append(1)
            own.append(DUMMY)
    close = np.array([DUMMY] + own, np.uint64)  # entry j + 1 = owner of head j  ->  close[j] = vertex completed by head j
    return np.array(words, np.int64), np.array(heads, np.int64), close


def sweep_piece(words, heads, close, jbase, rnd):
    """One warp piece, `rnd` edges per round.  Returns (stores {vertex: sum}, head_partial or None, tail_partial)."""
    stores, carry, seen, n_closed, head_partial = {}, 0, False, 0, None
    for r0 in range(0, len(words), rnd):
        w, h = words[r0:r0 + rnd], heads[r0:r0 + rnd]
        pos = np.nonzero(h)[0]
        if len(pos) == 0:
            carry += int(w.sum())
            continue
        first = carry + int(w[:pos[0]].sum())  # the carry of the previous rounds belongs to the round's first head
        sums = [first] + [int(w[pos[i]:pos[i + 1]].sum()) for i in range(len(pos) - 1)]
        for li, s in enumerate(sums):
            if li == 0 and not seen:
                head_partial = s  # the piece's first completion may have begun in earlier pieces: fix-up
                continue
            v = int(close[jbase + n_closed + li])
            if v != DUMMY:
                stores[v] = s
        carry = int(w[pos[-1]:].sum())
        seen = True
        n_closed += len(pos)
    return stores, head_partial, carry


def comb(f2, v2, f1, v1):
    return (f2 | f1, v2 if f2 else v1 + v2)


def chained_fixup(flags, tails, order=None):
    """pull_fixup_fused_kernel: returns the carry into every piece.  `order` = arrival order of the blocks (tickets)."""
    n = len(flags)
    nb = (n + FIX_BLOCK - 1) // FIX_BLOCK
    agg, incl, out = [None] * nb, [None] * nb, [0] * n
    for b in (order if order is not None else range(nb)):
        # a block only waits for blocks with a smaller ticket: in ticket order every predecessor has published
        lo, hi = b * FIX_BLOCK, min(n, (b + 1) * FIX_BLOCK)
        f, v, ex = 0, 0, []
        for t in range(lo, hi):
            ex.append((f, v))
            f, v = comb(int(flags[t]), int(tails[t]), f, v)
        agg[b] = (f, v)
        pf, pv, q = 0, 0, b - 1
        while q >= 0 and not pf:  # look-back stops at the first aggregate / prefix that contains a head
            have_prefix = incl[q] is not None
            qf, qv = incl[q] if have_prefix else agg[q]
            pf, pv = comb(pf, pv, qf, qv)
            if have_prefix:
                break
            q -= 1
        incl[b] = comb(f, v, pf, pv)
        for t in range(lo, hi):
            ef, ev = ex[t - lo]
            out[t] = comb(ef, ev, pf, pv)[1]
    return out


def run_model(row_end, vals, piece, rnd, stage, blocks=None, ticket_order=None):
    words, heads, close = build_stream(row_end, vals, stage, blocks)
    assert len(words) % stage == 0 and stage % piece == 0 and piece % rnd == 0
    n_pieces = len(words) // piece
    tile_v = np.concatenate([[0], np.cumsum(heads.reshape(n_pieces, piece).sum(1))]).astype(np.int64)
    out, head_p, tail_p = {}, [None] * n_pieces, [0] * n_pieces
    for t in range(n_pieces):
        st, hp, tl = sweep_piece(words[t * piece:(t + 1) * piece], heads[t * piece:(t + 1) * piece], close, int(tile_v[t]), rnd)
        out.update(st)
        head_p[t], tail_p[t] = hp, tl
    flags = (np.diff(tile_v) > 0).astype(np.int64)
    carry = chained_fixup(flags, tail_p, ticket_order(len(flags)) if ticket_order else None)
    for t in range(1, n_pieces):
        if flags[t]:
            v = int(close[tile_v[t]])
            if v != DUMMY:
                out[v] = carry[t] + head_p[t]
    return out


def direct_sums(row_end, vals):
    starts = np.concatenate([[0], row_end[:-1]]).astype(np.int64)
    return {v: int(vals[starts[v]:int(row_end[v])].sum()) for v in range(len(row_end)) if row_end[v] > starts[v]}


def random_csc(nv, seed, hub=None):
    rng = np.random.default_rng(seed)
    deg = rng.integers(0, 6, nv)
    deg[rng.random(nv) < 0.45] = 0  # many vertices without in-edges, like RMAT
    if hub is not None:
        deg[hub[0]] = hub[1]  # a hub spanning many pieces (and several fix-up blocks)
    row_end = np.cumsum(deg).astype(np.uint64)
    vals = rng.integers(1, 50, int(row_end[-1])).astype(np.int64)
    return row_end, vals


@pytest.mark.parametrize("nv,piece,rnd,stage,hub", [(50, 8, 4, 16, None), (400, 16, 8, 32, (7, 300)), (3000, 8, 8, 64, (0, 9000)),
                                                   (5000, 32, 16, 64, (4999, 40000)), (20000, 8, 4, 8, (100, 300000))])
def test_flagged_stream_sweep_equals_direct_sums(nv, piece, rnd, stage, hub):
    row_end, vals = random_csc(nv, nv, hub)
    assert run_model(row_end, vals, piece, rnd, stage) == direct_sums(row_end, vals)


def test_blocked_stream_with_padding_between_blocks():
    """Panel layout: several vertex blocks, each padded to whole stages (pads close the block's last vertex, then dummies)."""
    row_end, vals = random_csc(900, 5, (450, 2000))
    blocks = [(0, 300), (300, 300), (300, 451), (451, 900)]  # includes an empty block: a whole stage of pads
    assert run_model(row_end, vals, 8, 4, 32, blocks) == direct_sums(row_end, vals)


def test_hub_spanning_many_fixup_blocks():
    """~ 20 K pieces -> ~ 80 fix-up blocks; the hub's in-edge list covers most of them, so most look-backs walk over
    head-less aggregates until they meet the block where the hub began."""
    row_end, vals = random_csc(6000, 11, (3, 150000))
    assert run_model(row_end, vals, 8, 4, 8) == direct_sums(row_end, vals)
