"""Host-side model of the fused cross-GPU exchange (PeerExchange, csrc/kernels.h / payoff_reduce.cuh).

Threads stand in for the last blocks of G ranks.  The model has the same steps as the kernel code -- write my totals
into slot [epoch parity][my rank] of every rank's buffer, raise my flag there, wait until every rank's flag on MY
buffer has reached this epoch, fold the slots in rank order -- and random delays between all of them.  It checks the
protocol's bookkeeping (two parities are enough because a rank can be at most one launch ahead of the slowest one, flags
only grow, every rank folds the same values); memory ordering is the GPU's business and is covered by the -m gpu tests.
"""
import random
import threading
import time

import pytest


def _run(G, epochs, seed):
    slots = [[[None] * G for _ in range(2)] for _ in range(G)]     # slots[owner][parity][source rank]
    flags = [[0] * G for _ in range(G)]                            # flags[owner][source rank] = last epoch published
    results = [[None] * (epochs + 1) for _ in range(G)]
    max_lead = [0]
    errors = []

    def value(rank, epoch):
        return (rank + 1) * 1000003 + epoch * 7919

    def rank_main(me):
        rnd = random.Random(seed * 97 + me)
        try:
            for epoch in range(1, epochs + 1):
                par = epoch & 1
                for p in rnd.sample(range(G), G):                  # any order of peers
                    slots[p][par][me] = value(me, epoch)
                    if rnd.random() < 0.2:
                        time.sleep(rnd.random() * 1e-4)
                for p in rnd.sample(range(G), G):
                    flags[p][me] = epoch
                deadline = time.time() + 20.0
                while any(flags[me][r] < epoch for r in range(G)):
                    if time.time() > deadline:
                        raise TimeoutError(f"rank {me} epoch {epoch}")
                    time.sleep(0)
                max_lead[0] = max(max_lead[0], max(flags[me]) - epoch)
                if rnd.random() < 0.3:
                    time.sleep(rnd.random() * 2e-4)                # a slow reader: the others may already be writing
                results[me][epoch] = sum(slots[me][par][r] for r in range(G))
        except Exception as exc:      # noqa: BLE001
            errors.append(exc)

    threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(G)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(60)
    assert not errors, errors
    for epoch in range(1, epochs + 1):
        expect = sum(value(r, epoch) for r in range(G))
        assert all(results[r][epoch] == expect for r in range(G)), (epoch, [results[r][epoch] for r in range(G)])
    return max_lead[0]


@pytest.mark.parametrize("G", [2, 3, 8])
def test_two_slot_parities_are_enough(G):
    lead = _run(G, epochs=300, seed=G)
    assert lead <= 1          # nobody is ever more than one launch ahead of a rank that is still waiting


def test_single_rank_and_many_epochs():
    assert _run(1, epochs=50, seed=1) == 0


# ---- slot tags (round 2): launch number + call signature travel with the data and are verified by the folding rank

def _run_tagged(G, plan, seed, timeout=20.0):
    """plan[rank] = list of call signatures the rank issues, one per launch.  Returns results[rank][launch] where a
    result is the folded sum, "mismatch" (a slot carried another call's tag) or "timeout" (a peer never arrived)."""
    slots = [[[None] * G for _ in range(2)] for _ in range(G)]     # (value, tag)
    flags = [[0] * G for _ in range(G)]
    results = [[] for _ in range(G)]

    def rank_main(me):
        rnd = random.Random(seed * 131 + me)
        for i, sig in enumerate(plan[me]):
            epoch = i + 1
            par = epoch & 1
            tag = (sig, epoch)
            for p in range(G):
                slots[p][par][me] = ((me + 1) * 1000 + sig, tag)
            if rnd.random() < 0.3:
                time.sleep(rnd.random() * 1e-4)
            for p in range(G):
                flags[p][me] = epoch
            deadline = time.time() + timeout
            timed_out = False
            while any(flags[me][r] < epoch for r in range(G)):
                if time.time() > deadline:
                    timed_out = True
                    break
                time.sleep(0)
            if timed_out:
                results[me].append("timeout")
            elif any(slots[me][par][r][1] != tag for r in range(G)):
                results[me].append("mismatch")
            else:
                results[me].append(sum(slots[me][par][r][0] for r in range(G)))

    threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(G)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(120)
    return results


def test_matching_call_sequences_fold_to_the_same_sum():
    G = 4
    plan = [[5, 6, 7, 8, 9]] * G
    res = _run_tagged(G, plan, seed=3)
    for launch, sig in enumerate(plan[0]):
        expect = sum((r + 1) * 1000 + sig for r in range(G))
        assert all(res[r][launch] == expect for r in range(G))


def test_diverged_call_sequences_are_never_folded_silently():
    """One rank issues a different call in the middle of the sequence (what a failed launch or a mismatched host
    program produces): every rank sees 'mismatch' for that launch instead of a sum of unrelated numbers, and the
    launches after it are fine again."""
    G = 3
    plan = [[1, 2, 3, 4], [1, 2, 3, 4], [1, 9, 3, 4]]
    res = _run_tagged(G, plan, seed=5)
    ok = lambda sig: sum((r + 1) * 1000 + sig for r in range(G))      # noqa: E731
    assert [res[r][0] for r in range(G)] == [ok(1)] * G
    assert [res[r][1] for r in range(G)] == ["mismatch"] * G
    assert [res[r][2] for r in range(G)] == [ok(3)] * G and [res[r][3] for r in range(G)] == [ok(4)] * G


def test_a_rank_that_stops_launching_times_the_others_out():
    G = 3
    plan = [[1, 2, 3], [1, 2, 3], [1]]                                  # rank 2 stops after the first launch
    res = _run_tagged(G, plan, seed=7, timeout=2.0)
    assert res[2] == [sum((r + 1) * 1000 + 1 for r in range(G))]
    for r in (0, 1):
        assert res[r][0] == res[2][0] and res[r][1] == "timeout" and res[r][2] == "timeout"
