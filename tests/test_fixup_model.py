"""Executable model (numpy, CPU) of the pull kernel's cross-tile fix-up — the three-kernel segmented scan of
lux_b200/csrc/pull.cuh (pull_fixup_scan / _blocks / _apply) — checked against the obvious sequential definition:
carry into tile t = sum of the tail partials of the tiles since, and including, the last tile before t that completed
a vertex.  Mirrors the CUDA control flow (256-tile blocks, 1024 serial chunks + shuffle scans) so that a change of the
kernel's structure has a CPU-side regression check of the algebra, including n_blocks > 1024 (chunk length > 1)."""
import numpy as np
import pytest

FIX_BLOCK = 256


def comb(f2, v2, f1, v1):
    """(f1, v1) earlier, (f2, v2) later — seg_combine in pull.cuh."""
    return (f2 | f1, v2 if f2 else v1 + v2)


def seq_scan_exclusive(flags, vals):
    out_f, out_v = [], []
    f, v = 0, 0.0
    for ff, vv in zip(flags, vals):
        out_f.append(f)
        out_v.append(v)
        f, v = comb(int(ff), float(vv), f, v)
    return out_f, out_v, (f, v)


def model_fixup(flags, tails):
    n = len(flags)
    nb = (n + FIX_BLOCK - 1) // FIX_BLOCK
    carry_f, carry_v, agg = [0] * n, [0.0] * n, []
    for b in range(nb):  # pull_fixup_scan_kernel: exclusive in-block scan + block aggregate
        lo, hi = b * FIX_BLOCK, min(n, (b + 1) * FIX_BLOCK)
        ef, ev, total = seq_scan_exclusive(flags[lo:hi], tails[lo:hi])
        carry_f[lo:hi], carry_v[lo:hi] = ef, ev
        agg.append(total)
    # pull_fixup_blocks_kernel: 1024 threads, serial chunk of `per` blocks each, scan of chunk aggregates, rewrite
    per = (nb + 1023) // 1024
    chunk = []
    for k in range(1024):
        b0, b1 = min(k * per, nb), min(min(k * per, nb) + per, nb)
        f, v = 0, 0.0
        for b in range(b0, b1):
            f, v = comb(agg[b][0], agg[b][1], f, v)
        chunk.append((f, v))
    # two-level inclusive scan over the 1024 chunk aggregates (warps of 32, then warp aggregates)
    incl = []
    for w in range(32):
        f, v = 0, 0.0
        for lane in range(32):
            f, v = comb(chunk[w * 32 + lane][0], chunk[w * 32 + lane][1], f, v)
            incl.append((f, v))
    prefix = []
    for k in range(1024):
        w, lane = divmod(k, 32)
        wf, wv = 0, 0.0
        for ww in range(w):
            wf, wv = comb(incl[ww * 32 + 31][0], incl[ww * 32 + 31][1], wf, wv)
        pf, pv = (0, 0.0) if lane == 0 else incl[k - 1]
        prefix.append(comb(pf, pv, wf, wv))
    block_prefix = [None] * nb
    for k in range(1024):
        b0, b1 = min(k * per, nb), min(min(k * per, nb) + per, nb)
        pf, pv = prefix[k]
        for b in range(b0, b1):
            block_prefix[b] = (pf, pv)
            pf, pv = comb(agg[b][0], agg[b][1], pf, pv)
    # pull_fixup_apply_kernel
    out = []
    for t in range(n):
        c = carry_v[t] if carry_f[t] else block_prefix[t // FIX_BLOCK][1] + carry_v[t]
        out.append(c)
    return np.array(out)


@pytest.mark.parametrize("n,density", [(1, 1.0), (255, 0.5), (256, 0.0), (257, 0.01), (5000, 0.02), (300000, 0.0005), (300000, 0.3)])
def test_fixup_model_equals_sequential_definition(n, density):
    rng = np.random.default_rng(n)
    flags = (rng.random(n) < density).astype(np.int64)
    tails = rng.integers(0, 100, n).astype(np.float64)  # integers: every summation order is exact
    want = np.array(seq_scan_exclusive(flags, tails)[1])
    got = model_fixup(flags, tails)
    assert np.array_equal(got, want)
