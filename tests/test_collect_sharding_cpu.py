"""Data-parallel collection host logic on CPU with the gloo backend, world_size 2: contiguous row shards,
no data-path collective, optional gather of the results to rank 0 in the original order.  The model is a deterministic
stand-in with the reference-facing signature."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from openvla_probe_b200.collect import ShardedCollector, embeds_dict, episode_owner, shard_rows

LAYERS = [0, 5, -1]


class FakeVLA:
    """pooled[l][b] and actions[b] are fixed functions of observation b's inputs, so any routing error shows."""

    def __init__(self):
        self.calls = []

    def predict_action_and_capture(self, input_ids, unnorm_key=None, layer_indices=None, pooling_method="mean",
                                   pixel_values=None, **kw):
        b = input_ids.shape[0]
        self.calls.append(b)
        key = input_ids.double().sum(1) + pixel_values.double().flatten(1).sum(1)
        embeds = {l: (key[:, None] * (1 + abs(l)) + torch.arange(4)[None]).float().numpy() for l in layer_indices}
        actions = (key[:, None] * 0.5 + torch.arange(7)[None]).double().numpy()
        return embeds, actions


def _inputs(n):
    g = torch.Generator().manual_seed(0)
    return torch.randint(1, 100, (n, 6), generator=g), torch.randn(n, 3, 4, 4, generator=g)


def test_shard_rows_tile_the_batch():
    for n in (0, 1, 2, 7, 256, 257):
        for world in (1, 2, 3, 8):
            pieces = [shard_rows(n, world, r) for r in range(world)]
            assert pieces[0][0] == 0 and pieces[-1][1] == n
            assert all(pieces[i][1] == pieces[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in pieces]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    with pytest.raises(ValueError):
        shard_rows(4, 2, 2)
    assert [episode_owner(e, 4) for e in range(1, 6)] == [1, 2, 3, 0, 1]


def test_single_process_micro_batches():
    ids, px = _inputs(11)
    vla = FakeVLA()
    pooled, actions = ShardedCollector(vla, micro_batch=4).run(ids, px, "k", LAYERS)
    assert vla.calls == [4, 4, 3] and pooled.shape == (3, 11, 4) and actions.shape == (11, 7)
    want_e, want_a = FakeVLA().predict_action_and_capture(ids, layer_indices=LAYERS, pixel_values=px)
    assert np.array_equal(actions, want_a)
    d = embeds_dict(pooled, LAYERS)
    assert all(np.array_equal(d[l], want_e[l]) for l in LAYERS)


def test_ragged_micro_batches_are_trimmed_to_their_longest_row():
    """attention_mask of right-padded prompts travels with the rows and every micro-batch is cut to its own width."""
    ids, px = _inputs(7)
    lens = [6, 2, 3, 1, 4, 4, 5]
    mask = (torch.arange(6)[None] < torch.tensor(lens)[:, None]).long()

    class Rec(FakeVLA):
        def predict_action_and_capture(self, input_ids, attention_mask=None, **kw):
            self.seen = getattr(self, "seen", []) + [(tuple(input_ids.shape), attention_mask.sum(1).tolist())]
            return super().predict_action_and_capture(input_ids * attention_mask, **kw)

    vla = Rec()
    pooled, actions = ShardedCollector(vla, micro_batch=3).run(ids, px, "k", LAYERS, attention_mask=mask)
    assert vla.seen == [((3, 6), [6, 2, 3]), ((3, 4), [1, 4, 4]), ((1, 5), [5])]
    _, want_a = FakeVLA().predict_action_and_capture(ids * mask, layer_indices=LAYERS, pixel_values=px)
    assert np.array_equal(actions, want_a)


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ids, px = _inputs(n)
    col = ShardedCollector(FakeVLA(), micro_batch=3)
    lo, hi, p_loc, a_loc = col.run(ids, px, "k", LAYERS, gather=False)
    assert (lo, hi) == shard_rows(n, world, rank) and (p_loc is None) == (hi == lo)
    pooled, actions = col.run(ids, px, "k", LAYERS, gather=True)
    if rank == 0:
        q.put((pooled, actions))
    else:
        assert pooled is None and actions is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [7, 1, 8])
def test_two_rank_gather_equals_single_process(n):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    pooled, actions = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ids, px = _inputs(n)
    want_e, want_a = FakeVLA().predict_action_and_capture(ids, layer_indices=LAYERS, pixel_values=px)
    assert np.array_equal(actions, want_a) and pooled.shape == (3, n, 4)
    for i, l in enumerate(LAYERS):
        assert np.array_equal(pooled[i], want_e[l])
