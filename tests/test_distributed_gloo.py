"""world_size-2 run of the sharding logic on CPU (gloo): each rank takes its pair shard, the oracle
plays the kernel, torch.distributed sums the partials, and the result equals the unsharded run.
Also covers the unique-id broadcast plumbing ShardedPricer uses."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

WORLD = 2


def _worker(rank, port, R, nT, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(WORLD),
                      LOCAL_RANK=str(rank))
    import torch
    import torch.distributed as dist
    from cocos_b200 import distributed as cdist
    from oracle import heston_numpy as hn
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    r, w, _ = cdist.env_rank_world()
    assert (r, w) == (rank, WORLD)
    first, count = cdist.pair_shard(R, w, r)
    P = hn.REFERENCE_PARAMS
    H = (R + 1) // 2
    # the oracle simulates exactly the global pairs [first, first+count): partners exist for pairs < floor(R/2)
    n_partner = max(0, min(first + count, R // 2) - first)
    x, _ = hn.fused_model(P["x0"], P["v0"], P["r"], P["rho"], P["sigma_v"], P["kappa"], P["v_bar"], P["T"], nT,
                          2 * count, seed=31, subsequence=2, pair_offset=first, dtype=np.float64)
    x_local = np.concatenate([x[:count], x[count:count + n_partner]])
    pay = np.maximum(np.exp(x_local.astype(np.float64)) - P["K"], 0.0)
    t = torch.tensor([pay.sum(), float(len(pay))], dtype=torch.float64)
    dist.all_reduce(t)                                   # the single collective of the path
    # object broadcast used for the NCCL unique id
    box = [b"\x07" * 128 if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    assert box[0] == b"\x07" * 128
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), t.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("R", [2001, 4000])
def test_two_rank_shards_equal_single_run(tmp_path, R):
    import torch.multiprocessing as mp
    from oracle import heston_numpy as hn
    nT = 21
    port = 29500 + (os.getpid() + R) % 2000
    mp.start_processes(_worker, args=(port, R, nT, str(tmp_path)), nprocs=WORLD, join=True, start_method="spawn")
    got = [np.load(tmp_path / f"rank{r}.npy") for r in range(WORLD)]
    assert np.array_equal(got[0], got[1])                # every rank holds the reduced sums
    P = hn.REFERENCE_PARAMS
    x, _ = hn.fused_model(P["x0"], P["v0"], P["r"], P["rho"], P["sigma_v"], P["kappa"], P["v_bar"], P["T"], nT, R,
                          seed=31, subsequence=2, dtype=np.float64)   # float64: np.dot's size-dependent
    pay = np.maximum(np.exp(x.astype(np.float64)) - P["K"], 0.0)
    assert got[0][1] == R                                # shards cover every path exactly once
    assert got[0][0] == pytest.approx(pay.sum(), rel=1e-12)   # float32 rounding would blur the comparison


def test_pair_shards_cover_all_pairs():
    from cocos_b200 import distributed as cdist
    for R in (1, 2, 9, 10, 10_000_001):
        for world in (1, 2, 4, 8):
            spans = [cdist.pair_shard(R, world, r) for r in range(world)]
            assert sum(c for _, c in spans) == (R + 1) // 2
            pos = 0
            for first, count in spans:
                assert count == 0 or first == pos
                pos += count
