"""Multi-GPU execution of the hot path: one process per GPU, ``torch.distributed`` (NCCL over NVLink) for the
exchange steps.  The reference has no distributed code; what shards is dictated by the algorithm:

* sub-videos of a long clip are independent units up to halos  -> ``bench.py`` weak scaling: one sub-video per rank,
  no data-path collective;
* inside ONE sub-video (strong scaling, ``inpaint_clip_distributed``):
    RAFT frame pairs are independent          -> contiguous pair ranges per rank, one all-gather of the flows
    flow completion / image propagation       -> recurrent in time: chunks (when there are several) go round-robin
                                                 to ranks, a single chunk is computed redundantly by every rank
    sliding windows of the generator          -> contiguous window ranges per rank, one all-gather of the window
                                                 predictions; the order-dependent uint8 composite then runs on every
                                                 rank over the gathered predictions (identical result everywhere)

The collective payloads are small next to the compute (flows 2x[T-1,2,H,W] fp32, predictions
[sum l_t,H,W,4] fp16).  Sharding helpers are pure functions so they are covered by world-size-2 ``gloo`` tests on CPU.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous near-equal split of range(n): returns [lo, hi) of `rank` (earlier ranks get the remainder)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_sizes(n: int, world: int) -> List[int]:
    return [shard_range(n, world, r)[1] - shard_range(n, world, r)[0] for r in range(world)]


def round_robin(n: int, world: int, rank: int) -> List[int]:
    return list(range(rank, n, world))


def all_gather_variable(local: torch.Tensor, sizes: Sequence[int], group=None) -> torch.Tensor:
    """All-gather of tensors whose dim-0 lengths differ per rank (``sizes[r]``); returns their concatenation.

    One collective: shards are padded to the largest length, gathered, and trimmed."""
    world = len(sizes)
    if world == 1:
        return local
    mx = max(sizes)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad.contiguous(), group=group)
    return torch.cat([o[:s] for o, s in zip(out, sizes)], 0)


def window_shards(n_windows: int, world: int) -> List[Tuple[int, int]]:
    return [shard_range(n_windows, world, r) for r in range(world)]


def composite_order(schedule) -> Tuple[List[int], List[int]]:
    """Flat (frame id, first-visit flag) lists of the composite, in window order (propainter_inference.py:294-307)."""
    seen, ids, first = set(), [], []
    for nb, _ in schedule:
        for i in nb:
            ids.append(i)
            first.append(0 if i in seen else 1)
            seen.add(i)
    return ids, first


def inpaint_clip_distributed(models, frames, flow_masks, masks_dilated, orig_u8, cfg, group=None) -> torch.Tensor:
    """Strong-scaling pass over ONE clip shared by all ranks of `group`.  Inputs are replicated on every rank
    (reference layouts, see propainter_inference.process_inpainting); returns the composited uint8 frames
    [T,H,W,3] on every rank."""
    from . import propainter_inference as PI

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    eng = models.raft_model.engine
    T = cfg.video_length

    # ---- RAFT: pairs [lo, hi) need frames [lo, hi]
    n_pairs = T - 1
    lo, hi = shard_range(n_pairs, world, rank)
    if hi > lo:
        ff, fb = eng.raft_bidir(frames[0, lo:hi + 1], cfg.raft_iter)
    else:
        H, W = frames.shape[-2:]
        ff = torch.zeros(0, 2, H, W, device=eng.device)
        fb = torch.zeros_like(ff)
    sizes = shard_sizes(n_pairs, world)
    both = all_gather_variable(torch.stack([ff, fb], 1), sizes, group)   # one collective (fp32: N-GPU == 1-GPU bit for bit)
    gt = (both[:, 0].unsqueeze(0), both[:, 1].unsqueeze(0))

    # ---- recurrent stages: a single chunk is computed by every rank (no exchange needed)
    pred = PI.complete_flow(models.flow_model, gt, flow_masks, cfg.subvideo_length)
    uf, um = PI.image_propagation(models.inpaint_model, frames, masks_dilated, pred, cfg)

    # ---- generator windows: contiguous ranges, one all-gather of the predictions
    sched = PI.window_schedule(cfg)
    wlo, whi = shard_range(len(sched), world, rank)
    md = masks_dilated[0].to(device=eng.device, dtype=torch.float32).contiguous()
    eng.gen_begin(uf[0], md, um[0], pred[0][0], pred[1][0])
    H, W = frames.shape[-2:]
    if whi > wlo:
        mine = eng.gen_run(sched[wlo:whi])
    else:
        mine = torch.zeros(0, H, W, 4, device=eng.device, dtype=torch.float16)
    eng.gen_end()
    wsizes = [sum(len(nb) for nb, _ in sched[a:b]) for a, b in window_shards(len(sched), world)]
    preds = all_gather_variable(mine, wsizes, group)

    ids, first = composite_order(sched)
    ids_dev = torch.tensor(ids, dtype=torch.int32, device=eng.device)
    first_dev = torch.tensor(first, dtype=torch.int32, device=eng.device)
    orig = orig_u8.to(eng.device).contiguous()
    comp = torch.zeros_like(orig)
    o = 0
    for nb, _ in sched:
        n = len(nb)
        eng.composite(preds[o:o + n], md, orig, comp, ids_dev[o:o + n], first_dev[o:o + n])
        o += n
    return comp
