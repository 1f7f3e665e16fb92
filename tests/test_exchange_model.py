"""Executable model (CPU processes + shared memory) of the multi-GPU PageRank exchange protocol of lux_b200/csrc/api.cu
(pagerank_publish) and build.cuh (pack_push_kernel, flag_barrier_kernel, chunk_pull_kernel):

  * XT = [P equal hot chunks | P equal cold chunks], two buffers per rank; iteration i reads XT[cur], fills XT[1 - cur]
  * pack+push: every owned entry is stored into the buffer of the rank HOLDING its chunk (peer memory = a shared tensor)
  * flag barrier: rank r stores the epoch into word r of every peer's flag array and spins on its own words — ONE barrier
    per iteration
  * chunk pull: every rank copies the chunks it does not hold from their holders; the cold half later ("second stream"),
    only the next sweep waits for it
  * the next iteration pushes into the buffer the previous sweep read

Every rank is a process with random delays; each sweep checks that the buffer it reads holds exactly the values of the
previous iteration everywhere — i.e. one barrier per iteration plus double buffering is race-free, whatever the skew
between the ranks — and the index arithmetic (owner ranges vs equal chunks, padding) fills every position exactly once.
The device kernels are not involved (they need a GPU: tests/test_gpu_multi.py); this covers the host-side protocol."""
import random
import time

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp


def layout(P, hot_counts, cold_counts):
    """Owner ranges and equal chunk sizes as build_hot_layout computes them (api.cu: xt_hot_chunk / xt_cold_chunk =
    ceil(n / P) rounded up to a multiple of 32)."""
    hot_off = np.concatenate([[0], np.cumsum(hot_counts)]).astype(np.int64)
    cold_off = np.concatenate([[0], np.cumsum(cold_counts)]).astype(np.int64)
    pad32 = lambda n: (int(n) + 31) // 32 * 32  # noqa: E731
    Ch = pad32((hot_off[-1] + P - 1) // P)
    Cc = pad32((cold_off[-1] + P - 1) // P)
    return hot_off, cold_off, Ch, Cc


def value(pos, it):
    return (pos * 7919 + it * 104729) % 1000003


def _rank(r, P, hot_off, cold_off, Ch, Cc, xt, flags, iters, seed, errors):
    rng = random.Random(seed * 131 + r)
    cold_base = Ch * P
    H, C = int(hot_off[-1]), int(cold_off[-1])
    cur, epoch = 0, 0
    hot_own = np.arange(hot_off[r], hot_off[r + 1])
    cold_own = np.arange(cold_off[r], cold_off[r + 1])

    def barrier():
        nonlocal epoch
        epoch += 1
        for k in range(P):
            if k != r:
                flags[k][r] = epoch  # "st.release.sys" into the peer's flag array
        t0 = time.time()
        for k in range(P):
            while k != r and int(flags[r][k]) < epoch:
                if time.time() - t0 > 60:
                    errors[r] = -1
                    return False
        return True

    def pull(buf, base, chunk):
        for k in range(P):
            if k != r:
                lo = base + k * chunk
                xt[r][buf][lo:lo + chunk] = xt[k][buf][lo:lo + chunk]

    try:
        # iteration 0 publishes the initial values (luxb_init), then every iteration: sweep -> publish
        for it in range(iters):
            if it > 0:
                # ---- sweep: reads the whole XT[cur]; it must hold iteration it-1's values, everywhere ----
                time.sleep(rng.random() * 0.002)
                got_h = xt[r][cur][:H].numpy()
                got_c = xt[r][cur][cold_base:cold_base + C].numpy()
                if not (np.array_equal(got_h, value(np.arange(H), it - 1))
                        and np.array_equal(got_c, value(np.arange(C) + 10 ** 6, it - 1))):
                    errors[r] = it
                    return
                time.sleep(rng.random() * 0.002)  # a second look later in the sweep: nobody may have overwritten it meanwhile
                if not np.array_equal(xt[r][cur][:H].numpy(), got_h):
                    errors[r] = it
                    return
            nxt = 1 - cur
            # ---- pack + push: owner -> holder of the equal chunk ----
            for pos in hot_own:
                xt[int(pos // Ch)][nxt][pos] = int(value(pos, it))
            for pos in cold_own:
                xt[int(pos // Cc)][nxt][cold_base + pos] = int(value(pos + 10 ** 6, it))
            if not barrier():
                return
            pull(nxt, 0, Ch)          # hot: compute stream
            time.sleep(rng.random() * 0.001)  # the panel kernel runs here
            pull(nxt, cold_base, Cc)  # cold: second stream; the main sweep waits for it
            cur = nxt
        errors[r] = 0
    except Exception:  # noqa: BLE001
        errors[r] = -2
        raise


@pytest.mark.parametrize("P,hot_counts,cold_counts", [
    (2, [40, 3], [10, 150]),
    (3, [100, 20, 1], [5, 60, 400]),       # rank 0 owns most hot entries, rank 2 most cold ones (like RMAT)
    (4, [0, 64, 0, 33], [31, 0, 97, 1]),   # ranks owning nothing of a region; sizes around the 32-element padding
])
def test_one_barrier_per_iteration_with_double_buffering_is_race_free(P, hot_counts, cold_counts):
    hot_off, cold_off, Ch, Cc = layout(P, hot_counts, cold_counts)
    n = (Ch + Cc) * P
    xt = [[torch.full((n,), -1, dtype=torch.int64).share_memory_() for _ in range(2)] for _ in range(P)]
    flags = [torch.zeros(P, dtype=torch.int64).share_memory_() for _ in range(P)]
    errors = torch.full((P,), -3, dtype=torch.int64).share_memory_()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_rank, args=(r, P, hot_off, cold_off, Ch, Cc, xt, flags, 40, 5, errors)) for r in range(P)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
    alive = [p for p in procs if p.is_alive()]
    for p in alive:
        p.terminate()
    assert not alive
    assert errors.tolist() == [0] * P, errors.tolist()


def test_every_position_has_exactly_one_owner_and_one_holder():
    for P, hc, cc in [(2, [40, 3], [10, 150]), (8, [500, 90, 40, 20, 9, 5, 2, 1], [3, 9, 30, 90, 200, 500, 900, 4000])]:
        hot_off, cold_off, Ch, Cc = layout(P, hc, cc)
        assert Ch % 32 == 0 and Cc % 32 == 0 and Ch * P >= hot_off[-1] and Cc * P >= cold_off[-1]
        for off, chunk in ((hot_off, Ch), (cold_off, Cc)):
            owners = np.searchsorted(off, np.arange(off[-1]), side="right") - 1
            assert np.array_equal(np.bincount(owners, minlength=P), np.diff(off))
            holders = np.arange(off[-1]) // chunk
            assert holders.max() < P
