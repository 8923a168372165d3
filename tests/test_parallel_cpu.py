"""Multi-rank host logic on the CPU (gloo, world_size 2): video assignment, frame shards with the
window halo, and the label gather -- the only collective of the path."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from playaid_core_b200 import parallel
from playaid_core_b200.dataset_utils import window_index_table


def test_assign_videos_balanced():
    lengths = [10800, 25200, 600, 10800, 3600, 10800, 64, 25200]
    parts = parallel.assign_videos(lengths, 4)
    assert sorted(i for p in parts for i in p) == list(range(len(lengths)))
    loads = [sum(lengths[i] for i in p) for p in parts]
    assert max(loads) - min(loads) <= max(lengths)
    assert parallel.assign_videos(lengths, 1) == [list(range(len(lengths)))]


@pytest.mark.parametrize("n,world", [(10800, 8), (64, 2), (100, 3), (7, 4)])
def test_frame_shards_cover_video_and_halo_suffices(n, world):
    wf = window_index_table(np.arange(n), 7, 3, max_frames=n)
    covered = []
    for r in range(world):
        lo, hi, hlo, hhi = parallel.frame_shard(n, r, world, reach=27)
        covered += list(range(lo, hi))
        if hi > lo:
            rows = wf[lo:hi]
            assert rows.min() >= hlo and rows.max() < hhi  # every window index lies inside the halo
            idx = parallel.shard_window_rows(wf, lo, hi, hlo, 2)
            assert idx.shape == ((hi - lo) * 2, 7) and idx.min() >= 0 and idx.max() < (hhi - hlo) * 2
            assert (idx[1::2] - idx[0::2] == 1).all()
    assert covered == list(range(n))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 101  # odd: ragged shards
    lo, hi, _, _ = parallel.frame_shard(n, rank, world)
    local = torch.stack([torch.arange(lo, hi, dtype=torch.int32) * 2, torch.arange(lo, hi, dtype=torch.int32) * 2 + 1], 1)
    gathered, lens = parallel.gather_labels(local)
    merged = parallel.merge_frame_shards(gathered, lens)
    ok = merged.shape == (n, 2) and bool((merged.reshape(-1) == torch.arange(2 * n, dtype=torch.int32)).all())
    lp = torch.full((hi - lo, 2, 3), float(rank))
    g2, l2 = parallel.gather_labels(lp, fill=0)
    ok = ok and g2.shape[0] == world and int(l2.sum()) == n and float(g2[1, 0, 0, 0]) == 1.0
    q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_labels_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]
