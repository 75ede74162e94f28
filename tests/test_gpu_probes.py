"""GPU parity of the probe-training kernels against torch autograd (fp32) on identical inputs.

Stated tolerance: the two GEMMs of a step multiply in TF32 (fp32 accumulate) where the reference multiplies in fp32,
so logits / dW are compared with rel-L2 <= 2e-3; loss <= 1e-3 relative; everything that is not a GEMM (BCE gradient,
bias gradient, AdamW) is compared at fp32 round-off (1e-5).
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _data(N=700, D=256, L=37, seed=0, spatial=False):
    g = torch.Generator().manual_seed(seed)
    X = torch.randn(N, D, generator=g)
    if spatial:
        Y = (torch.rand(N, L, generator=g) > 0.6).to(torch.int8)
    else:
        Y = torch.randint(-1, 2, (N, L), generator=g).to(torch.int8)
    keep = torch.tensor(sorted(np.random.default_rng(seed).choice(L, size=L - 4, replace=False).tolist()))
    return X, Y, keep


def _ref_loss(kind, Wd, X, Yk, pw):
    """The reference's loss expressions, verbatim semantics, on fp32 torch CPU."""
    import torch.nn.functional as F

    if kind == "dual":
        zp = X @ Wd["presence_head.weight"].t() + Wd["presence_head.bias"]
        zt = X @ Wd["truth_head.weight"].t() + Wd["truth_head.bias"]
        pres_t, truth_t, mask = (Yk != -1).float(), (Yk == 1).float(), (Yk != -1)
        lp = F.binary_cross_entropy_with_logits(zp, pres_t, pos_weight=pw, reduction="mean")
        lt_el = F.binary_cross_entropy_with_logits(zt, truth_t, reduction="none")
        lt = (lt_el * mask.float()).sum() / mask.sum()
        return lp + lt, torch.cat([zp, zt], 1)
    z = X @ Wd["weight"].t() + Wd["bias"]
    if kind == "object":
        mask = (Yk != -1)
        el = F.binary_cross_entropy_with_logits(z, (Yk == 1).float(), pos_weight=pw, reduction="none")
        return (el * mask.float()).sum() / mask.sum(), z
    return F.binary_cross_entropy_with_logits(z, Yk.float(), pos_weight=pw, reduction="mean"), z


@pytest.mark.parametrize("kind", ["object", "spatial", "dual"])
def test_probe_steps_vs_autograd_and_adamw(kind):
    from openvla_probe_b200.probes import ProbeTrainer

    X, Y, keep = _data(spatial=(kind == "spatial"))
    K, D = len(keep), X.shape[1]
    pw = torch.tensor(3.7) if kind == "dual" else (0.5 + 3 * torch.rand(K, generator=torch.Generator().manual_seed(5)))
    torch.manual_seed(3)
    tr = ProbeTrainer(kind, D, K, pw, batch=256)
    sd0 = {k: v.clone() for k, v in tr.state_dict().items()}
    params = {k: v.clone().requires_grad_(True) for k, v in sd0.items()}
    opt = torch.optim.AdamW(list(params.values()), lr=1e-3, weight_decay=1e-4)
    perm = torch.randperm(X.shape[0], generator=torch.Generator().manual_seed(9))
    tr.load_epoch(X.cuda(), Y.cuda(), keep, perm, drop_last=(kind == "dual"))
    n_steps = len(tr.steps)
    assert n_steps == (X.shape[0] // 256 if kind == "dual" else -(-X.shape[0] // 256))
    for s in range(n_steps):
        idx = perm[s * 256: (s + 1) * 256]
        Xb, Yb = X[idx], Y[idx][:, keep]
        loss_ref, z_ref = _ref_loss(kind, params, Xb, Yb, pw)
        opt.zero_grad()
        loss_ref.backward()
        # forward logits of the same step, before the update
        z = tr.logits(Xb.cuda()).cpu()
        zk = torch.cat([z[:, :K], z[:, tr.Kpad:tr.Kpad + K]], 1) if kind == "dual" else z[:, :K]
        assert (zk - z_ref.detach()).norm() / z_ref.detach().norm() < 2e-3
        tr.train_step(s)
        assert abs(tr.step_loss() - float(loss_ref)) <= 1e-3 * abs(float(loss_ref)) + 1e-5
        # gradient check (un-normalised device gradient / global count == autograd gradient)
        G = tr.G.cpu()
        stats = G[tr.n_total:]
        dW = G[: tr.n_w].view(tr.rows, D)
        db = G[tr.n_w: tr.n_total]
        heads = [("presence_head.", 0, stats[1]), ("truth_head.", tr.Kpad, stats[3])] if kind == "dual" else [("", 0, stats[1])]
        for pre, off, cnt in heads:
            gw, gb = params[pre + "weight"].grad, params[pre + "bias"].grad
            assert (dW[off: off + K] / cnt - gw).norm() / gw.norm() < 2e-3
            assert (db[off: off + K] / cnt - gb).norm() / gb.norm() < 2e-3
        opt.step()
        # Adam divides by |g|: a TF32-level perturbation of a near-zero gradient can flip the sign of an lr-sized
        # update, so parameters are compared in norm against the distance travelled (the AdamW kernel itself is
        # checked at fp32 round-off in test_adamw_kernel_exact).
        for k, v in tr.state_dict().items():
            moved = (params[k].detach() - sd0[k]).norm()
            assert (v - params[k].detach()).norm() <= 0.05 * moved + 1e-6, k
    # padding label rows never move
    W = tr.P[: tr.n_w].view(tr.rows, D)
    assert float(W[K: tr.Kpad].abs().max()) == 0.0


def test_bce_grad_exact_fp32_pieces():
    """BCE gradient / loss / counts / bias row-sums against closed forms (no GEMM involved -> fp32 round-off)."""
    import ctypes as C

    from openvla_probe_b200 import _lib

    lib = _lib.load()
    n, K, Kpad = 77, 13, 16
    g = torch.Generator().manual_seed(1)
    Z = (torch.randn(n, Kpad, generator=g) * 3).cuda()
    Y = torch.randint(-1, 2, (n, Kpad), generator=g).to(torch.int8)
    Y[:, K:] = -1
    pw = (0.5 + torch.rand(Kpad, generator=g)).cuda()
    dZT = torch.zeros(Kpad, 80, device="cuda")
    stats = torch.zeros(4, device="cuda")
    _lib.check(lib.ovla_probe_bce_grad(C.c_void_p(Z.data_ptr()), C.c_longlong(Kpad), C.c_void_p(Y.cuda().data_ptr()), n, K,
                                       Kpad, 0, 1, C.c_void_p(pw.data_ptr()), C.c_float(1.0), C.c_void_p(dZT.data_ptr()),
                                       C.c_longlong(80), C.c_void_p(stats.data_ptr()), None))
    z, y, p = Z.cpu()[:, :K].double(), Y[:, :K], pw.cpu()[:K].double()
    t, m = (y == 1).double(), (y != -1).double()
    sig = torch.sigmoid(z)
    grad = m * (sig * (1 + (p - 1) * t) - p * t)
    loss = (m * ((1 - t) * z + (1 + (p - 1) * t) * torch.nn.functional.softplus(-z))).sum()
    assert torch.allclose(dZT.cpu()[:K, :n].double(), grad.t(), atol=1e-5)
    assert float(dZT.cpu()[K:, :].abs().max()) == 0.0
    s = stats.cpu().double()
    assert abs(s[0] - loss) < 1e-3 and s[1] == m.sum()
    out = torch.zeros(Kpad, device="cuda")
    _lib.check(lib.ovla_probe_rowsum(C.c_void_p(dZT.data_ptr()), C.c_longlong(80), Kpad, n, C.c_void_p(out.data_ptr()), None))
    assert torch.allclose(out.cpu()[:K].double(), grad.sum(0), atol=1e-4)


def test_adamw_kernel_exact():
    """ovla_probe_adamw == torch.optim.AdamW on identical gradients (fp32 round-off), incl. the on-device 1/count."""
    import ctypes as C

    from openvla_probe_b200 import _lib

    lib = _lib.load()
    rows, D = 24, 64
    n_w, n_tot = rows * D, rows * D + rows
    g = torch.Generator().manual_seed(2)
    p0 = torch.randn(n_tot, generator=g) * 0.1
    ref = p0.clone().requires_grad_(True)
    opt = torch.optim.AdamW([ref], lr=1e-3, weight_decay=1e-4)
    P, M, V = p0.clone().cuda(), torch.zeros(n_tot).cuda(), torch.zeros(n_tot).cuda()
    cnt = [37.0, 11.0]
    stats = torch.tensor([0.0, cnt[0], 0.0, cnt[1]]).cuda()
    for step in range(1, 6):
        graw = torch.randn(n_tot, generator=g)
        scale = torch.empty(n_tot)
        rows_idx = torch.cat([torch.arange(n_w) // D, torch.arange(rows)])
        scale = torch.where(rows_idx < rows // 2, torch.tensor(cnt[0]), torch.tensor(cnt[1]))
        ref.grad = graw / scale
        opt.step()
        G = graw.cuda()
        _lib.check(lib.ovla_probe_adamw(C.c_void_p(P.data_ptr()), C.c_void_p(G.data_ptr()), C.c_void_p(M.data_ptr()),
                                        C.c_void_p(V.data_ptr()), C.c_longlong(n_w), D, rows // 2, C.c_longlong(n_tot),
                                        C.c_void_p(stats.data_ptr()), C.c_float(1e-3), C.c_float(0.9), C.c_float(0.999),
                                        C.c_float(1e-8), C.c_float(1e-4), step, None))
        assert torch.allclose(P.cpu(), ref.detach(), rtol=1e-5, atol=1e-6), step


def test_episode_round_trip_and_training_loop(tmp_path):
    """Files written by EpisodeWriter train through the per-layer driver and produce the reference's output files."""
    from openvla_probe_b200.probes import EpisodeWriter, train_probes

    rng = np.random.default_rng(0)
    n_rel, n_act, D = 30, 6, 128
    w_true = rng.normal(size=(D, n_rel + n_act))
    for ep in range(1, 13):
        wr = EpisodeWriter(layers=[0, 1])
        T = 40
        feats = rng.normal(size=(2, T, D)).astype(np.float32)
        lab = (feats[1] @ w_true > 0).astype(np.int8)              # layer 1 is linearly decodable, layer 0 is noise
        lab[rng.random(lab.shape) < 0.3] = -1
        wr.append_batch(feats, lab[:, :n_rel], lab[:, n_rel:])
        wr.save(str(tmp_path / "logs" / f"episode_{ep}.pt"))
    d = torch.load(str(tmp_path / "logs" / "episode_3.pt"), weights_only=False)
    assert set(d) == {"visual_semantic_encoding", "symbolic_state_object_relations", "symbolic_state_action_subgoals"}
    assert d["visual_semantic_encoding"][1].shape == (40, D) and d["visual_semantic_encoding"][1].dtype == torch.float32
    assert d["symbolic_state_object_relations"].dtype == torch.int8
    recs = train_probes("object", str(tmp_path / "logs"), [0, 1], epochs=120, batch=128, out_dir=str(tmp_path / "out"),
                        verbose=False)
    # torch AdamW on the same data reaches ~0.83 after 100 epochs (440 samples); the noise layer stays near chance
    assert recs[1]["val_acc"] > 0.78 and recs[0]["val_acc"] < 0.62
    ck = torch.load(str(tmp_path / "out" / "linear_probe_L01.pth"), weights_only=False)
    assert set(ck) == {"state_dict", "layer", "kept"} and ck["state_dict"]["weight"].shape == (len(ck["kept"]), D)
    recs = train_probes("dual", str(tmp_path / "logs"), [1], epochs=120, batch=128, out_dir=str(tmp_path / "out"),
                        verbose=False)
    ck = torch.load(str(tmp_path / "out" / "linear_probe_dual_head_final_L01.pth"), weights_only=False)
    assert ck["model_type"] == "DualHeadProbe" and "presence_head.weight" in ck["state_dict"]
    assert recs[0]["truth_acc_va"] > 0.72


def test_3class_probe_steps_vs_autograd():
    """Direct 3-class probe (train_3class_direct.py:147-212): weighted CrossEntropy over [B*K, 3] logits."""
    from openvla_probe_b200.probes import ProbeTrainer

    X, Y, keep = _data(N=520, D=128, L=21)
    K, D = len(keep), X.shape[1]
    cw = torch.tensor([0.6, 1.1, 1.3])
    torch.manual_seed(4)
    tr = ProbeTrainer("3class", D, K, cw, batch=256)
    sd0 = tr.state_dict()
    assert sd0["weight"].shape == (3 * K, D)
    params = {k: v.clone().requires_grad_(True) for k, v in sd0.items()}
    opt = torch.optim.AdamW(list(params.values()), lr=1e-3, weight_decay=1e-4)
    crit = torch.nn.CrossEntropyLoss(weight=cw)
    perm = torch.randperm(X.shape[0], generator=torch.Generator().manual_seed(9))
    tr.load_epoch(X.cuda(), Y.cuda(), keep, perm, drop_last=True)
    assert len(tr.steps) == 2
    for s in range(2):
        idx = perm[s * 256: (s + 1) * 256]
        Xb, Yb = X[idx], Y[idx][:, keep].long()
        z = Xb @ params["weight"].t() + params["bias"]
        loss = crit(z.view(-1, 3), (Yb + 1).view(-1))
        opt.zero_grad()
        loss.backward()
        tr.train_step(s)
        assert abs(tr.step_loss() - float(loss.detach())) <= 1e-3 * float(loss.detach()) + 1e-5
        G = tr.G.cpu()
        cnt = G[tr.n_total + 1]
        dW = G[: tr.n_w].view(tr.rows, D)[: 3 * K] / cnt
        db = G[tr.n_w: tr.n_w + 3 * K] / cnt
        assert (dW - params["weight"].grad).norm() / params["weight"].grad.norm() < 2e-3
        assert (db - params["bias"].grad).norm() / params["bias"].grad.norm() < 2e-3
        opt.step()
        for k, v in tr.state_dict().items():
            moved = (params[k].detach() - sd0[k]).norm()
            assert (v - params[k].detach()).norm() <= 0.05 * moved + 1e-6, k


@pytest.mark.parametrize("kind", ["object", "spatial", "dual", "3class"])
def test_on_device_validation_counters_match_host_metrics(kind):
    """evaluate(on_device=True): confusion counts computed next to the logits (ovla_probe_confusion) give the same
    accuracy / F1 as the reference's gather-to-host + sklearn path, and the counts themselves equal numpy's."""
    from openvla_probe_b200.probes import ProbeTrainer, confusion_counts, evaluate

    X, Y, keep = _data(N=900, D=128, L=29, seed=3, spatial=(kind == "spatial"))
    K = len(keep)
    pw = torch.tensor(1.7) if kind == "dual" else (torch.ones(3) if kind == "3class" else torch.full((K,), 1.3))
    tr = ProbeTrainer(kind, 128, K, pw, batch=256, device=0)
    tr.fit(X, Y, keep, epochs=2, seed=0)
    host = evaluate(kind, tr, X, Y, keep)
    dev = evaluate(kind, tr, X, Y, keep, on_device=True)
    for k, v in dev.items():
        assert abs(v - host[k]) < 1e-6, (k, v, host[k])      # the host path averages in fp32
    Z = tr.logits(X.cuda().contiguous())
    counts = confusion_counts(kind, tr, Z, Y, keep)
    y = Y[:, keep].long().numpy()
    z = Z.cpu().numpy()
    if kind == "3class":
        pred = z[:, :3 * K].reshape(-1, 3).argmax(1)
        tgt = (y + 1).reshape(-1)
        want = [int(((tgt == a) & (pred == b)).sum()) for a in range(3) for b in range(3)]
        assert counts == want
    elif kind == "dual":
        assert sum(counts[:4]) == y.size and sum(counts[4:8]) == int((y != -1).sum())
    else:
        assert sum(counts[:4]) == (int((y != -1).sum()) if kind == "object" else y.size)


def test_full_size_object_probe_step_vs_autograd():
    """BASELINE config [3] at full size: batch 4096 x 4096-d features x 439 kept labels (461 + 20 minus 42 dropped),
    one object-probe step against fp32 autograd (run on the GPU with TF32 disabled as the checker), same tolerances as
    the small-shape test: logits / dW rel-L2 <= 2e-3, loss <= 1e-3 relative."""
    from openvla_probe_b200.probes import ProbeTrainer

    N, D, Lbl = 4096, 4096, 481
    g = torch.Generator().manual_seed(11)
    X = torch.randn(N, D, generator=g)
    Y = torch.randint(-1, 2, (N, Lbl), generator=g).to(torch.int8)
    keep = torch.tensor(sorted(np.random.default_rng(2).choice(Lbl, size=439, replace=False).tolist()))
    K = len(keep)
    pw = 0.5 + 3 * torch.rand(K, generator=g)
    torch.manual_seed(3)
    tr = ProbeTrainer("object", D, K, pw, batch=N)
    sd0 = tr.state_dict()
    perm = torch.arange(N)
    tr.load_epoch(X.cuda(), Y.cuda(), keep, perm, drop_last=False)
    assert len(tr.steps) == 1
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        Wr = sd0["weight"].cuda().clone().requires_grad_(True)
        br = sd0["bias"].cuda().clone().requires_grad_(True)
        Xd, Yk = X.cuda(), Y[:, keep].cuda()
        z_ref = Xd @ Wr.t() + br
        mask = (Yk != -1)
        el = torch.nn.functional.binary_cross_entropy_with_logits(z_ref, (Yk == 1).float(), pos_weight=pw.cuda(), reduction="none")
        loss_ref = (el * mask.float()).sum() / mask.sum()
        loss_ref.backward()
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    z = tr.logits(Xd)[:, :K]
    assert float((z - z_ref.detach()).norm() / z_ref.detach().norm()) < 2e-3
    tr.train_step(0)
    assert abs(tr.step_loss() - float(loss_ref.detach())) <= 1e-3 * abs(float(loss_ref.detach()))
    G = tr.G
    cnt = G[tr.n_total + 1]
    dW = G[: tr.n_w].view(tr.rows, D)[:K] / cnt
    db = G[tr.n_w: tr.n_total][:K] / cnt
    assert float((dW - Wr.grad).norm() / Wr.grad.norm()) < 2e-3
    assert float((db - br.grad).norm() / br.grad.norm()) < 2e-3
    assert int(cnt.item()) == int(mask.sum().item())


def test_per_label_confusion_counts_on_device():
    """ovla_probe_confusion_per_label: [K, 4] counts (tp, fp, fn, tn) of every kept label equal numpy's on the same
    logits (eval_probes_per_label.py:59-96 semantics: mask y != -1, target y == 1, pred sigmoid(z) > 0.5)."""
    from openvla_probe_b200.probes import ProbeTrainer, per_label_counts, per_label_metrics

    X, Y, keep = _data(N=777, D=128, L=33, seed=5)
    K = len(keep)
    tr = ProbeTrainer("object", 128, K, torch.full((K,), 1.5), batch=256, device=0)
    tr.fit(X, Y, keep, epochs=1, seed=0)
    Z = tr.logits(X.cuda().contiguous())
    counts = per_label_counts(tr, Z, Y, keep)
    z = Z[:, :K].cpu()
    pred = (torch.sigmoid(z) > 0.5).numpy()
    y = Y[:, keep].numpy()
    want = np.zeros((K, 4), dtype=np.int64)
    for k in range(K):
        m = y[:, k] != -1
        t, q = y[m, k] == 1, pred[m, k]
        want[k] = [(t & q).sum(), (~t & q).sum(), (t & ~q).sum(), (~t & ~q).sum()]
    assert counts.shape == (K, 4) and np.array_equal(counts, want)
    assert len(per_label_metrics(counts, keep.tolist())) == K


# ----------------------------------------------------------------------------------------------- all layers at once
def _oracle_model(kind, D, K, sd):
    from oracle import probe_oracle as PO

    m = PO.DualHeadProbe(D, K) if kind == "dual" else torch.nn.Linear(D, K)
    m.load_state_dict(sd)
    return m


@pytest.mark.parametrize("kind", ["object", "spatial", "dual"])
def test_grouped_trainer_follows_the_oracle_and_the_single_layer_trainer(kind):
    """G probes trained concurrently (grouped GEMMs + fused deterministic BCE / bias-gradient kernel + grouped AdamW)
    against (a) the oracle's loss (the reference scripts' expressions) + torch.optim.AdamW per layer on the same batch
    order, (b) the per-layer device trainer.  Tolerances as for the single-layer path: TF32 GEMMs -> loss 1e-3 relative,
    weights rel-L2 2e-3, accumulated update rel-L2 5e-2."""
    from oracle import probe_oracle as PO
    from openvla_probe_b200.probes import MultiLayerProbeTrainer, ProbeTrainer

    G, B = 3, 256
    X0, Y, keep = _data(N=900, spatial=(kind == "spatial"))
    Xs = torch.stack([X0 * (1.0 + 0.25 * g) + 0.1 * g for g in range(G)])               # [G, N, D]
    K, D = len(keep), X0.shape[1]
    pw = torch.tensor(3.7) if kind == "dual" else (0.5 + 3 * torch.rand(K, generator=torch.Generator().manual_seed(5)))
    torch.manual_seed(11)
    singles = [ProbeTrainer(kind, D, K, pw, batch=B) for _ in range(G)]
    inits = [t.state_dict() for t in singles]
    multi = MultiLayerProbeTrainer(kind, G, D, K, pw, batch=B, init_states=inits)
    models = [_oracle_model(kind, D, K, sd) for sd in inits]
    opts = [torch.optim.AdamW(m.parameters(), lr=1e-3, weight_decay=1e-4) for m in models]
    drop_last = kind == "dual"
    gen = torch.Generator().manual_seed(9)
    Xd, Yd = Xs.cuda(), Y.cuda()
    for epoch in range(2):
        perm = torch.randperm(X0.shape[0], generator=gen)
        multi.load_epoch(Xd, Yd, keep, perm, drop_last)
        for g in range(G):
            singles[g].load_epoch(Xd[g], Yd, keep, perm, drop_last)
        assert len(multi.steps) == len(singles[0].steps) == (X0.shape[0] // B if drop_last else -(-X0.shape[0] // B))
        for s in range(len(multi.steps)):
            idx = perm[s * B:(s + 1) * B]
            ref_losses = []
            for g in range(G):
                loss, _ = PO.loss_fn(kind, models[g], Xs[g][idx], Y[idx][:, keep], pw)
                opts[g].zero_grad(); loss.backward(); opts[g].step()
                ref_losses.append(float(loss))
                singles[g].train_step(s)
            multi.train_step(s)
            got = multi.step_losses()
            np.testing.assert_allclose(got, ref_losses, rtol=1e-3, atol=1e-5)
            np.testing.assert_allclose(got, [t.step_loss() for t in singles], rtol=2e-4, atol=1e-6)
    multi.finish()
    for g in range(G):
        sd, ref, one = multi.state_dict(g), models[g].state_dict(), singles[g].state_dict()
        for k in ref:
            a, r, i0 = sd[k].float(), ref[k].float(), inits[g][k].float()
            assert float((a - r).norm() / r.norm()) < 2e-3, (kind, g, k)
            assert float(((a - i0) - (r - i0)).norm() / (r - i0).norm()) < 5e-2, (kind, g, k)
            assert float((a - one[k]).norm() / one[k].norm()) < 2e-3, (kind, g, k)
    # padded rows / columns never move
    assert float(multi.P[:, : multi.n_w].view(G, multi.rows, D)[:, K:multi.Kpad].abs().max()) == 0.0


def test_grouped_step_is_deterministic_and_handles_ragged_batches():
    """Two trainers with the same initial weights and batch order end bit-identical (the grouped BCE / bias-gradient kernel
    reduces in a fixed order, no floating-point atomics); the last, ragged batch (n % 128 != 0, n % 4 != 0) is covered."""
    from openvla_probe_b200.probes import MultiLayerProbeTrainer

    G, B = 5, 192
    X0, Y, keep = _data(N=777, D=128, L=70)
    Xs = torch.stack([X0 + 0.05 * g for g in range(G)]).cuda()
    K, D = len(keep), X0.shape[1]
    pw = 0.5 + 3 * torch.rand(K, generator=torch.Generator().manual_seed(5))
    torch.manual_seed(1)
    a = MultiLayerProbeTrainer("object", G, D, K, pw, batch=B)
    inits = [a.state_dict(g) for g in range(G)]
    b = MultiLayerProbeTrainer("object", G, D, K, pw, batch=B, init_states=inits, chunks=2)
    perm = torch.randperm(X0.shape[0], generator=torch.Generator().manual_seed(2))
    for t in (a, b):
        t.load_epoch(Xs, Y.cuda(), keep, perm, drop_last=False)
        assert t.steps[-1][1] - t.steps[-1][0] == 777 % B
        for s in range(len(t.steps)):
            t.train_step(s)
        t.finish()
    assert torch.equal(a.P, b.P) and torch.equal(a.Gbuf, b.Gbuf)
    assert bool(torch.isfinite(a.P).all())


def test_full_size_grouped_probe_step_vs_fp32():
    """BASELINE configs[3] shape for several layers at once: batch 4096, D = 4096, K = 439, dual heads -- logits and the
    un-normalised gradient of every layer against fp32 torch on the device (TF32 tolerance 2e-3 rel-L2)."""
    from openvla_probe_b200.probes import MultiLayerProbeTrainer

    G, B, D, L = 4, 4096, 4096, 481
    g = torch.Generator().manual_seed(0)
    X = torch.randn(G, B, D, generator=g)
    Y = torch.randint(-1, 2, (B, L), generator=g).to(torch.int8)
    keep = torch.arange(439)
    torch.manual_seed(0)
    tr = MultiLayerProbeTrainer("dual", G, D, 439, torch.tensor(1.7), batch=B)
    Xd = X.cuda()
    tr.load_epoch(Xd, Y.cuda(), keep, torch.arange(B), drop_last=True)
    P0 = tr.P.clone()
    tr.train_step(0)
    tr.finish()
    torch.backends.cuda.matmul.allow_tf32 = False
    Yk = Y[:, keep].cuda()
    for gi in range(G):
        W = P0[gi, : tr.n_w].view(tr.rows, D)
        bias = P0[gi, tr.n_w:]
        z = Xd[gi] @ W.t() + bias
        zp, zt = z[:, :439], z[:, tr.Kpad:tr.Kpad + 439]
        pres_t, truth_t, mask = (Yk != -1).float(), (Yk == 1).float(), (Yk != -1).float()
        gp = (torch.sigmoid(zp) * (1 + 0.7 * pres_t) - 1.7 * pres_t)          # d/dz of pos-weighted BCE, un-normalised
        gt = (torch.sigmoid(zt) - truth_t) * mask
        dW_ref = torch.cat([gp.t() @ Xd[gi], gt.t() @ Xd[gi]])
        dW = torch.cat([tr.Gbuf[gi, : tr.n_w].view(tr.rows, D)[:439], tr.Gbuf[gi, : tr.n_w].view(tr.rows, D)[tr.Kpad:tr.Kpad + 439]])
        assert float((dW - dW_ref).norm() / dW_ref.norm()) < 2e-3, gi
        db = tr.Gbuf[gi, tr.n_w: tr.n_total]
        db_ref = torch.cat([gp.sum(0), gt.sum(0)])
        assert float((torch.cat([db[:439], db[tr.Kpad:tr.Kpad + 439]]) - db_ref).norm() / db_ref.norm()) < 1e-4, gi
        st = tr.Gbuf[gi, tr.n_total:]
        assert float(st[1]) == B * 439 and float(st[3]) == float(mask.sum())
