"""Data-parallel collection of (pooled hidden states, actions) over the GPUs of one box.

SURVEY.md 8(e): every observation is independent (the reference never batches at all), so the path shards with NO
data-path collective: weights are replicated, rank r takes a contiguous slice of the observations and runs them through
`predict_action_and_capture` in micro-batches.  The only optional exchange is a gather of the results to rank 0 when one
process should write the `episode_*.pt` files (run_libero_eval_object.py:357-366); otherwise every rank writes the
episodes it owns.  Launch: `torchrun --nproc-per-node N ...` (one process per GPU), or a single process.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


def world_info(group=None) -> Tuple[int, int]:
    """(world_size, rank) of the initialised process group, (1, 0) without one."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def shard_rows(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Rows [lo, hi) of rank `rank`: contiguous, in rank order, sizes differ by at most one (the first n % world ranks
    take the extra row); empty for ranks >= n."""
    if n < 0 or world < 1 or not 0 <= rank < world:
        raise ValueError(f"shard_rows(n={n}, world={world}, rank={rank})")
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def episode_owner(episode_id: int, world: int) -> int:
    """Rank that writes `episode_{id}.pt` when every rank writes its own files (round-robin over episode numbers)."""
    return episode_id % world


class ShardedCollector:
    """Runs a batch of observations through `vla.predict_action_and_capture` on this rank's shard.

    `vla` is any object with the reference-facing signature
    `predict_action_and_capture(input_ids, unnorm_key=..., layer_indices=..., pooling_method=..., pixel_values=...)`
    returning `({layer: float32 [b, D]}, float64 [b, action_dim])` (modeling_prismatic.OpenVLAForActionPrediction).
    """

    def __init__(self, vla, micro_batch: int = 256, group=None):
        if micro_batch < 1:
            raise ValueError("micro_batch must be >= 1")
        self.vla, self.micro_batch, self.group = vla, micro_batch, group
        self.world, self.rank = world_info(group)

    def run_local(self, input_ids: torch.Tensor, pixel_values: torch.Tensor, unnorm_key: Optional[str],
                  layer_indices: Sequence[int], pooling_method: str = "mean", attention_mask: Optional[torch.Tensor] = None):
        """This rank's rows of the GLOBAL inputs [n, ...].  Returns (lo, hi, pooled float32 [L, hi-lo, D] or None when the
        shard is empty, actions float64 [hi-lo, A]).  `attention_mask` [n, P] marks right-padded (ragged) prompts; each
        micro-batch is trimmed to its own longest row before it is sent to the device."""
        n = input_ids.shape[0]
        if pixel_values.shape[0] != n:
            raise ValueError("input_ids and pixel_values disagree on the number of observations")
        lo, hi = shard_rows(n, self.world, self.rank)
        layers = list(layer_indices)
        pooled, actions = [], []
        for s in range(lo, hi, self.micro_batch):
            e = min(s + self.micro_batch, hi)
            ids_mb, kw = input_ids[s:e], {}
            if attention_mask is not None:
                m = attention_mask[s:e]
                width = max(1, int((m != 0).sum(1).max()))
                ids_mb, kw = ids_mb[:, :width], {"attention_mask": m[:, :width]}
            embeds, act = self.vla.predict_action_and_capture(
                ids_mb, unnorm_key=unnorm_key, layer_indices=layers, pooling_method=pooling_method,
                pixel_values=pixel_values[s:e], **kw)
            pooled.append(np.stack([np.asarray(embeds[l], dtype=np.float32).reshape(e - s, -1) for l in layers]))
            actions.append(np.asarray(act, dtype=np.float64).reshape(e - s, -1))
        if not pooled:
            return lo, hi, None, np.zeros((0, 0), dtype=np.float64)
        return lo, hi, np.concatenate(pooled, 1), np.concatenate(actions, 0)

    def run(self, input_ids: torch.Tensor, pixel_values: torch.Tensor, unnorm_key: Optional[str],
            layer_indices: Sequence[int], pooling_method: str = "mean", gather: bool = False,
            attention_mask: Optional[torch.Tensor] = None):
        """Shard, run, and (gather=True) assemble the global result on rank 0 in the original row order.

        Returns `(pooled [L, n, D], actions [n, A])` on rank 0 (and on every rank when world == 1); other ranks get
        `(None, None)` with gather=True, or their local `(lo, hi, pooled, actions)` tuple with gather=False."""
        lo, hi, pooled, actions = self.run_local(input_ids, pixel_values, unnorm_key, layer_indices, pooling_method,
                                                 attention_mask)
        if self.world == 1:
            return pooled, actions
        if not gather:
            return lo, hi, pooled, actions
        # The one optional exchange of the path: variable-sized shards -> gather_object keeps it simple (138 MB at n = 256;
        # it is off the per-step hot path -- once per collected batch, to the writer rank only).
        payload = (lo, hi, pooled, actions)
        bucket = [None] * self.world if self.rank == 0 else None
        dist.gather_object(payload, bucket, dst=0, group=self.group)
        if self.rank != 0:
            return None, None
        parts = sorted((p for p in bucket if p[2] is not None), key=lambda p: p[0])
        if not parts:
            return None, np.zeros((0, 0), dtype=np.float64)
        cover = 0
        for p in parts:
            if p[0] != cover:
                raise RuntimeError(f"shards do not tile the batch: expected row {cover}, got {p[0]}")
            cover = p[1]
        return np.concatenate([p[2] for p in parts], 1), np.concatenate([p[3] for p in parts], 0)


def embeds_dict(pooled: np.ndarray, layer_indices: Sequence[int]) -> Dict[int, np.ndarray]:
    """[L, n, D] -> {layer: [n, D]} in the reference's return convention (openvla_utils.py:193-207)."""
    return {l: pooled[i] for i, l in enumerate(layer_indices)}
