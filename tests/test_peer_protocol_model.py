"""Model check of the peer-halo ordering (PeerArgs, soap_bubble_simulation_b200/csrc/sbx_internal.cuh; wait_neighbour /
peer_signal_done in sbx_tma.cuh): ranks run their steps at their own pace, the band next to a neighbour reads the
neighbour's boundary row in place from the ping-pong half of the previous step, and a single monotonic flag per band
("finished step e", published in the writer's own memory) has to cover both hazards:

  RAW  the row a band reads at step e is the neighbour's row of step e-1, complete;
  WAR  a band does not overwrite (at step e+1) the row its neighbour is still reading for step e.

The model executes the micro-operations of every boundary band (wait, remote read, local write, publish) under random
interleavings and checks the version every read observes.  The multi-GPU parity tests (tests/test_gpu_multi.py) check the
real kernels; this one documents why one flag per band is enough and that dropping either wait is caught.
"""
import random

import pytest

TOP, BOT = 0, 1


def run(ranks: int, steps: int, seed: int, wait_top: bool = True, wait_bot: bool = True):
    """Returns the list of violations (empty = every read saw the version it should)."""
    rng = random.Random(seed)
    # row[r][band][half]: version of rank r's first (TOP) / last (BOT) owned row in each ping-pong half.  The upload
    # fills the half step 1 reads (half 0) with version 0; the other half holds nothing useful yet.
    row = [[[0, -1], [0, -1]] for _ in range(ranks)]
    done = [[0, 0] for _ in range(ranks)]  # published flags: done[r][band] = last step that band finished
    step = [1] * ranks                       # the launch each rank is executing
    pc = [[0, 0] for _ in range(ranks)]      # micro-op counter of each boundary band inside that launch
    bad = []

    def has_band(r, b):
        return (b == TOP and r > 0) or (b == BOT and r < ranks - 1)

    def neighbour(r, b):  # (rank, band) whose row this band reads and whose flag it polls
        return (r - 1, BOT) if b == TOP else (r + 1, TOP)

    while True:
        runnable = []
        for r in range(ranks):
            if step[r] > steps:
                continue
            for b in (TOP, BOT):
                if not has_band(r, b) or pc[r][b] == 4:
                    continue
                if pc[r][b] == 0 and (wait_top if b == TOP else wait_bot):
                    nr, nb = neighbour(r, b)
                    if done[nr][nb] < step[r] - 1:
                        continue  # still polling the neighbour's flag
                runnable.append((r, b))
        if not runnable:
            break
        r, b = rng.choice(runnable)
        e = step[r]
        nr, nb = neighbour(r, b)
        op = pc[r][b]
        if op == 1:    # bulk copy of the neighbour's row out of the half step e reads
            got = row[nr][nb][(e - 1) % 2]
            if got != e - 1:
                bad.append((r, b, e, got))
        elif op == 2:  # my boundary row of step e goes into the half step e writes
            row[r][b][e % 2] = e
        elif op == 3:  # __threadfence_system + st.release.sys of the flag in my own memory
            done[r][b] = e
        pc[r][b] += 1
        # the next launch's CTAs pass griddepcontrol.wait only when every band of this one has finished
        if all(pc[r][bb] == 4 or not has_band(r, bb) for bb in (TOP, BOT)):
            step[r] += 1
            pc[r] = [0, 0]
    assert all(s == steps + 1 for s in step), f"deadlock: ranks stopped at steps {step}"
    return bad


@pytest.mark.parametrize("ranks", [2, 3, 8])
def test_one_flag_per_band_orders_reads_and_overwrites(ranks):
    for seed in range(200):
        assert run(ranks, steps=12, seed=seed) == []


def test_dropping_a_wait_is_caught():
    # without the wait a rank that runs ahead reads a half its neighbour has not written yet (RAW), or overwrites a
    # row the neighbour has not read yet (WAR): both show up as a wrong version on some interleaving
    assert any(run(2, 12, seed, wait_top=False) for seed in range(50))
    assert any(run(2, 12, seed, wait_bot=False) for seed in range(50))
    assert any(run(8, 12, seed, wait_top=False, wait_bot=False) for seed in range(20))


def test_raw_and_war_are_both_exercised():
    """Both hazard kinds occur when the waits are off (so the green run above is not vacuous for either)."""
    stale, clobbered = False, False
    for seed in range(200):
        for (_, _, e, got) in run(3, 12, seed, wait_top=False, wait_bot=False):
            stale |= got < e - 1      # the neighbour had not produced step e-1 yet
            clobbered |= got > e - 1  # the neighbour had already overwritten it with step e+1
    assert stale and clobbered
