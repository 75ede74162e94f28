"""Host logic of the multi-GPU probe step, exercised on CPU with the gloo backend (world_size 2):
batch sharding, the single flat allreduce of [dW | db | loss, count], and the global `mask.sum()` normaliser.
The per-rank gradient here is a plain-torch stand-in for the CUDA kernels (same un-normalised definition)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from openvla_probe_b200.probes import (adamw_reference_step, allreduce_flat, layer_matrix, prepare_dual,
                                        prepare_object, prepare_spatial, shard_batches, split_episodes)


def test_shard_batches_partition():
    for n, batch, world, drop in [(1000, 256, 2, False), (1000, 256, 3, True), (7, 4, 8, False), (4096, 4096, 8, False)]:
        per_rank = [shard_batches(n, batch, world, r, drop) for r in range(world)]
        n_steps = n // batch if drop else -(-n // batch)
        assert all(len(p) == n_steps for p in per_rank)
        for s in range(n_steps):
            lo, hi = s * batch, min((s + 1) * batch, n)
            pieces = [per_rank[r][s] for r in range(world)]
            assert pieces[0][0] == lo and pieces[-1][1] == hi
            assert all(pieces[i][1] == pieces[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in pieces]
            assert max(sizes) - min(sizes) <= 1


def _local_unnormalised(W, b, X, Y, pw):
    """Masked pos-weighted BCE, gradient and loss summed (not averaged) over the local rows."""
    z = X @ W.t() + b
    t, m = (Y == 1).float(), (Y != -1).float()
    lw = 1 + (pw - 1) * t
    loss = (m * ((1 - t) * z + lw * torch.nn.functional.softplus(-z))).sum()
    g = m * (torch.sigmoid(z) * lw - pw * t)
    return g.t() @ X, g.sum(0), loss, m.sum()


def _worker(rank, world, port, X, Y, pw, W0, b0, perm, batch, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    W, b = W0.clone(), b0.clone()
    P = torch.cat([W.flatten(), b])
    M, V = torch.zeros_like(P), torch.zeros_like(P)
    steps = shard_batches(X.shape[0], batch, world, rank, False)
    for s, (lo, hi) in enumerate(steps):
        idx = perm[lo:hi]
        K, D = W0.shape
        Wc, bc = P[: K * D].view(K, D), P[K * D:]
        dW, db, loss, cnt = _local_unnormalised(Wc, bc, X[idx], Y[idx], pw)
        flat = torch.cat([dW.flatten(), db, torch.stack([loss, cnt])])
        allreduce_flat(flat)
        g = flat[:-2] / flat[-1]
        adamw_reference_step(P, g, M, V, s + 1)
    if rank == 0:
        out.put(P.clone())
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_step_equals_single_process_step():
    torch.manual_seed(0)
    N, D, K, batch = 300, 32, 11, 64
    X = torch.randn(N, D)
    Y = torch.randint(-1, 2, (N, K)).to(torch.int8)
    pw = 0.5 + torch.rand(K) * 4
    lin = torch.nn.Linear(D, K)
    W0, b0 = lin.weight.detach().clone(), lin.bias.detach().clone()
    perm = torch.randperm(N)
    # single process, torch autograd + torch.optim.AdamW = the reference's step (train_object_probes.py:184-189)
    params = [W0.clone().requires_grad_(True), b0.clone().requires_grad_(True)]
    opt = torch.optim.AdamW(params, lr=1e-3, weight_decay=1e-4)
    bce = torch.nn.BCEWithLogitsLoss(reduction="none", pos_weight=pw)
    for s in range(-(-N // batch)):
        idx = perm[s * batch: (s + 1) * batch]
        z = X[idx] @ params[0].t() + params[1]
        mask = (Y[idx] != -1)
        loss = (bce(z, (Y[idx] == 1).float()) * mask.float()).sum() / mask.sum()
        opt.zero_grad()
        loss.backward()
        opt.step()
    want = torch.cat([params[0].detach().flatten(), params[1].detach()])
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, X, Y, pw, W0, b0, perm, batch, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert torch.allclose(got, want, rtol=0, atol=2e-6)


def _fake_cache(n_eps=20, T=15, n_rel=12, n_act=4, D=8, spatial=False, seed=0):
    rng = np.random.default_rng(seed)
    cache = {}
    for i in range(n_eps):
        lab = rng.integers(0 if spatial else -1, 2, (T, n_rel + n_act)).astype(np.int8)
        lab[:, 0] = 1            # constant column -> dropped by the keep filters
        cache[i] = {
            "visual_semantic_encoding": {0: torch.from_numpy(rng.normal(size=(T, D)).astype(np.float32))},
            "symbolic_state_object_relations": torch.from_numpy(lab[:, :n_rel]),
            "symbolic_state_action_subgoals": torch.from_numpy(lab[:, n_rel:]),
        }
    return cache


def test_split_keep_and_pos_weight_rules():
    cache = _fake_cache()
    tr, va = split_episodes(cache)
    assert len(va) == 2 and len(tr) == 18 and not set(tr) & set(va)
    import random
    ids = list(cache.keys()); random.Random(0).shuffle(ids)
    assert va == ids[:2] and tr == ids[2:]                     # train_object_probes.py:72-75
    sp = prepare_object(cache)
    assert 0 not in sp.keep.tolist() and len(sp.keep) == 15 and float(sp.pos_weight.max()) <= 20
    Y = torch.cat([torch.cat([cache[i]["symbolic_state_object_relations"], cache[i]["symbolic_state_action_subgoals"]], 1)
                   for i in sp.train_ids])
    k = int(sp.keep[3])
    want = (float(((Y[:, k] == 0)).sum()) + 1) / (float((Y[:, k] == 1).sum()) + 1)
    assert abs(float(sp.pos_weight[3]) - min(want, 20.0)) < 1e-6
    sd = prepare_dual(cache)
    assert sd.pos_weight.dim() == 0 and 0 not in sd.keep.tolist()
    ss = prepare_spatial(_fake_cache(spatial=True))
    assert 0 not in ss.keep.tolist()
    X, Yl = layer_matrix(cache, sp.train_ids, 0)
    assert X.shape == (18 * 15, 8) and Yl.shape == (18 * 15, 16) and Yl.dtype == torch.int8
    assert layer_matrix(cache, sp.train_ids, 5)[0].numel() == 0  # layer absent -> empty (StepDS `layer in cache[...]`)


def test_packed_episode_store_round_trip(tmp_path):
    """Packed memory-mapped store <-> the reference's episode_*.pt dict layout (run_libero_eval_object.py:357-366)."""
    from openvla_probe_b200.probes import PackedEpisodeStore, load_episodes

    rng = np.random.default_rng(0)
    st = PackedEpisodeStore(str(tmp_path / "store"), n_layers=3, dim=16, n_rel=5, n_act=2, capacity=64)
    ref = []
    for ep in range(3):
        for _ in range(2):
            pooled = rng.normal(size=(3, 4, 16)).astype(np.float32)
            rel = rng.integers(-1, 2, (4, 5)).astype(np.int8)
            act = rng.integers(-1, 2, (4, 2)).astype(np.int8)
            st.append_batch(pooled, rel, act)
            ref.append((ep, pooled, rel, act))
        st.end_episode()
    st.flush()
    rd = PackedEpisodeStore(str(tmp_path / "store"), mode="r")
    assert rd.n_rows == 24 and rd.episodes == [(0, 8), (8, 8), (16, 8)] and rd.layer(2).shape == (24, 16)
    rd.export_reference_episodes(str(tmp_path / "logs"))
    cache = load_episodes(str(tmp_path / "logs"))
    assert len(cache) == 3
    d = cache[1]
    assert set(d) == {"visual_semantic_encoding", "symbolic_state_object_relations", "symbolic_state_action_subgoals"}
    want = np.concatenate([p[1][2] for p in ref if p[0] == 1])
    assert np.array_equal(d["visual_semantic_encoding"][2].numpy(), want)
    assert d["symbolic_state_object_relations"].dtype == torch.int8 and d["symbolic_state_object_relations"].shape == (8, 5)


def test_metrics_from_confusion_counts_match_sklearn():
    """acc / F1 computed from integer confusion counts == the sklearn calls the reference makes on gathered arrays
    (train_object_probes.py:190-206, train_dual_head_final.py:196-232, train_3class_direct.py:196-207)."""
    from sklearn.metrics import f1_score

    from openvla_probe_b200.probes import KIND_3CLASS, KIND_DUAL, KIND_OBJECT, KIND_SPATIAL, metrics_from_counts

    rng = np.random.default_rng(0)
    for trial in range(6):
        n = 500
        t = rng.integers(0, 2, n) if trial < 4 else np.zeros(n, dtype=np.int64)       # degenerate: a single class
        p = rng.integers(0, 2, n) if trial != 5 else np.zeros(n, dtype=np.int64)
        tp, fp = int(((p == 1) & (t == 1)).sum()), int(((p == 1) & (t == 0)).sum())
        fn, tn = int(((p == 0) & (t == 1)).sum()), int(((p == 0) & (t == 0)).sum())
        for kind in (KIND_OBJECT, KIND_SPATIAL):
            m = metrics_from_counts(kind, [tp, fp, fn, tn])
            assert abs(m["val_acc"] - float((p == t).mean())) < 1e-12
            assert abs(m["val_f1"] - f1_score(t, p, average="macro", zero_division=0)) < 1e-12
        m = metrics_from_counts(KIND_DUAL, [tp, fp, fn, tn, tp, fp, fn, tn])
        assert abs(m["pres_f1_va"] - f1_score(t, p, average="binary", pos_label=1, zero_division=0)) < 1e-12
        assert abs(m["truth_f1_va"] - f1_score(t, p, labels=[0, 1], average="macro", zero_division=0)) < 1e-12
        assert abs(m["truth_acc_va"] - float((p == t).mean())) < 1e-12
    t3, p3 = rng.integers(0, 3, 700), rng.integers(0, 3, 700)
    conf = [int(((t3 == a) & (p3 == b)).sum()) for a in range(3) for b in range(3)]
    m = metrics_from_counts(KIND_3CLASS, conf)
    assert abs(m["val_acc"] - float((p3 == t3).mean())) < 1e-12
    assert abs(m["val_f1"] - f1_score(t3, p3, labels=[0, 1, 2], average="macro", zero_division=0)) < 1e-12


def test_per_label_metrics_match_sklearn():
    """per_label_metrics (from integer counts) == the sklearn calls of eval_probes_per_label.py:77-96, including the
    single-class NaN convention and the empty-mask skip."""
    from sklearn.metrics import balanced_accuracy_score, matthews_corrcoef, precision_recall_fscore_support

    from openvla_probe_b200.probes import per_label_metrics

    rng = np.random.default_rng(3)
    N, K = 400, 9
    y = rng.integers(-1, 2, (N, K))
    p = rng.integers(0, 2, (N, K))
    y[:, 2] = -1                                  # empty mask -> skipped
    y[:, 3] = np.where(y[:, 3] == 0, 1, y[:, 3])  # single class (all positives) -> mcc / bal_acc NaN
    p[:, 4] = 0                                   # never predicted positive -> precision 0 by zero_division
    counts = np.zeros((K, 4), dtype=np.int64)
    for k in range(K):
        m = y[:, k] != -1
        t, q = (y[m, k] == 1), p[m, k] == 1
        counts[k] = [(t & q).sum(), (~t & q).sum(), (t & ~q).sum(), (~t & ~q).sum()]
    keep = list(range(10, 10 + K))
    recs = per_label_metrics(counts, keep)
    assert [r["label_idx"] for r in recs] == [10, 11, 13, 14, 15, 16, 17, 18]
    for r in recs:
        k = r["label_idx"] - 10
        m = y[:, k] != -1
        targ, pred = (y[m, k] == 1).astype(int), p[m, k]
        pr, rc, f1, _ = precision_recall_fscore_support(targ, pred, average="binary", pos_label=1, zero_division=0)
        assert abs(r["prec"] - pr) < 1e-12 and abs(r["recall"] - rc) < 1e-12 and abs(r["f1"] - f1) < 1e-12
        if len(np.unique(targ)) > 1:
            assert abs(r["mcc"] - matthews_corrcoef(targ, pred)) < 1e-12
            assert abs(r["bal_acc"] - balanced_accuracy_score(targ, pred)) < 1e-12
        else:
            assert np.isnan(r["mcc"]) and np.isnan(r["bal_acc"])
