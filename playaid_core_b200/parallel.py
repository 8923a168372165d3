"""Multi-GPU partitioning of the path: one process per GPU, no data-path collective.

Every (frame, fighter) crop and every window is independent given read-only weights, so the work
shards with plain index arithmetic:

* many videos  -- whole videos are dealt to ranks, longest first (`assign_videos`), the B200 form of
  reference playaid/multi_manuscript.sh:1-7 (one `manuscript.py` process per video);
* one long video -- contiguous frame ranges per rank; windows reach +-reach frames
  (27 at delta 3, playaid/dataset_utils.py:123), so each rank also computes features for a
  `reach`-frame halo on either side and no feature exchange is needed (`frame_shard`).

The only communication is the final gather of per-frame labels (and optionally log-probs):
`gather_labels` pads each rank's block to the longest shard and issues ONE
`torch.distributed.all_gather_into_tensor` -- NCCL over NVLink on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def assign_videos(lengths, world_size: int) -> list[list[int]]:
    """Greedy longest-first assignment of videos (by frame count) to ranks; returns video ids per rank."""
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    load = [0] * world_size
    out: list[list[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += int(lengths[i])
    return [sorted(v) for v in out]


def frame_shard(n_frames: int, rank: int, world_size: int, reach: int = 27):
    """Contiguous shard of one video for `rank`: (own_lo, own_hi, halo_lo, halo_hi).
    Frames [own_lo, own_hi) are labelled by this rank; features are needed for [halo_lo, halo_hi)."""
    base, rem = divmod(n_frames, world_size)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi, max(0, lo - reach), min(n_frames, hi + reach)


def shard_window_rows(win_frames: np.ndarray, own_lo: int, own_hi: int, halo_lo: int, n_fighters: int) -> np.ndarray:
    """Window table of the whole video (frame numbers [N,S]) -> int32 feature-row indices
    [(own_hi-own_lo)*F, S] relative to a feature table that starts at frame `halo_lo`."""
    wf = win_frames[own_lo:own_hi].astype(np.int64) - halo_lo
    idx = wf[:, None, :] * n_fighters + np.arange(n_fighters)[None, :, None]
    return np.ascontiguousarray(idx.reshape(-1, wf.shape[1]).astype(np.int32))


def gather_labels(local: torch.Tensor, max_len: int | None = None, group=None, fill: int = -1):
    """All-gather per-rank label blocks [n_local, ...] of unequal length.
    Returns (gathered [world, max_len, ...], lengths [world]) on every rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    n = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    if world == 1:
        return local[None], n
    lens = torch.empty((world,), dtype=torch.int64, device=local.device)
    dist.all_gather_into_tensor(lens, n, group=group)
    if max_len is None:
        max_len = int(lens.max().item())
    pad = torch.full((max_len,) + tuple(local.shape[1:]), fill, dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world,) + tuple(pad.shape), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out.view(-1), pad.view(-1), group=group)
    return out, lens


def merge_frame_shards(gathered: torch.Tensor, lens: torch.Tensor) -> torch.Tensor:
    """Concatenate the valid prefix of every rank's block (contiguous frame shards -> whole video)."""
    return torch.cat([gathered[r, : int(lens[r])] for r in range(gathered.shape[0])], dim=0)
