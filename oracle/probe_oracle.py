"""CPU restatement of the reference's probe trainers -- TEST INFRASTRUCTURE ONLY.

Only tests/, tests/golden/make_probe_golden.py, __graft_entry__.smoke() and bench.py's CPU legs may import this file;
the product path (openvla_probe_b200/probes.py) never does.

Restates, in plain PyTorch fp32 on the CPU, the three training scripts of helenlu66/openvla-probe
(`experiment_utils/train_object_probes.py`, `train_spatial_probes.py`, `train_dual_head_final.py`) and
`train_3class_direct.py`: episode split, label filter, pos_weight / class weights, loss, AdamW, validation metrics.
It issues the same torch calls in the same order as the scripts (nn.Linear init, DataLoader iterators, BCEWithLogitsLoss,
AdamW), so that under the same global torch seed it reproduces a script run BIT FOR BIT.

Pinned: tests/golden/make_probe_golden.py runs the reference's UNMODIFIED scripts here (runpy, global seed set by a
wrapper, synthetic episode files written by the product's EpisodeWriter) and asserts that this restatement yields the
identical `kept` list, pos_weight, saved `.pth` state dicts (torch.equal) and CSV metrics; it then stores the per-step
batch orders, losses and states as tests/golden/probe_golden.pt for the GPU parity tests.
"""
from __future__ import annotations

import random
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn as nn
import torch.optim as optim
from torch.utils.data import DataLoader, Dataset

OBJECT, SPATIAL, DUAL, THREE = "object", "spatial", "dual", "3class"


def stack_labels(cache: Dict[int, dict], ids: Sequence[int]) -> torch.Tensor:
    """train_object_probes.py:78-84 / train_dual_head_final.py:85-90."""
    return torch.cat([torch.cat([cache[i]["symbolic_state_object_relations"],
                                 cache[i]["symbolic_state_action_subgoals"]], 1) for i in ids], 0)


def split_episodes(cache: Dict[int, dict], seed: int = 0):
    """train_object_probes.py:50,72-75 (random.Random(0)); train_dual_head_final.py:76-82 (random.Random(args.seed))."""
    rng = random.Random(seed)
    ep_ids = list(cache.keys())
    rng.shuffle(ep_ids)
    val_len = max(1, int(0.10 * len(ep_ids)))
    return ep_ids[val_len:], ep_ids[:val_len]


def prepare_object(cache):
    """train_object_probes.py:86-102: keep = labels that show both 0 and 1 (over train U val);
    POS_W = clamp((neg + 1) / (pos + 1), max=20)[keep] counted on the TRAIN split, -1 masked out."""
    train_ids, val_ids = split_episodes(cache)
    Y_full = stack_labels(cache, train_ids + val_ids)
    mask_full = Y_full != -1
    keep = (((Y_full == 1) & mask_full).any(0) & ((Y_full == 0) & mask_full).any(0)).nonzero(as_tuple=True)[0]
    Y_tr = stack_labels(cache, train_ids)
    mask_tr = Y_tr != -1
    pos_cnt = ((Y_tr == 1) & mask_tr).sum(0).float()
    neg_cnt = ((Y_tr == 0) & mask_tr).sum(0).float()
    pos_w = ((neg_cnt + 1.0) / (pos_cnt + 1.0))[keep].clamp(max=20)
    return train_ids, val_ids, keep, pos_w


def prepare_spatial(cache, all_label_mats: Optional[Sequence[torch.Tensor]] = None):
    """train_spatial_probes.py:96-131: keep is decided over ALL episode files (also the excluded ones, `quick_stack(all_files)`)
    by `(Y == 1).any(0) & (Y == 0).any(0)`; pos_cnt = Y_train.sum(0) (labels assumed {0,1}), neg_cnt = N_train - pos_cnt,
    POS_W = clamp((neg + 1) / (pos + 1), max=20)[keep]."""
    train_ids, val_ids = split_episodes(cache)
    Y_full = torch.cat(list(all_label_mats), 0) if all_label_mats is not None else stack_labels(cache, sorted(cache))
    keep = ((Y_full == 1).any(0) & (Y_full == 0).any(0)).nonzero(as_tuple=True)[0]
    Y_train = stack_labels(cache, train_ids)
    pos_cnt = Y_train.sum(0).float()
    neg_cnt = Y_train.shape[0] - pos_cnt
    pos_w = ((neg_cnt + 1.0) / (pos_cnt + 1.0))[keep].clamp(max=20)
    return train_ids, val_ids, keep, pos_w


def _freq_keep(Y_tr: torch.Tensor) -> torch.Tensor:
    """train_dual_head_final.py:96-113 / train_3class_direct.py:96-110: keep labels whose TRAIN 0/1 frequency is in (1 %, 99 %)."""
    m = Y_tr != -1
    cnt = m.sum(0)
    freq = torch.zeros_like(cnt, dtype=torch.float32)
    ok = cnt > 0
    freq[ok] = ((Y_tr == 1) & m).sum(0)[ok].float() / cnt[ok]
    freq[~ok] = -1.0
    keep = ((freq > 0.01) & (freq < 0.99)).nonzero(as_tuple=True)[0]
    if len(keep) == 0:
        keep = torch.arange(Y_tr.shape[1])
    return keep


def prepare_dual(cache, seed: int = 0):
    """train_dual_head_final.py:76-127: presence pos_weight = #absent / (#present + 1e-9) over kept TRAIN labels."""
    train_ids, val_ids = split_episodes(cache, seed)
    Y_tr = stack_labels(cache, train_ids)
    keep = _freq_keep(Y_tr)
    present = Y_tr[:, keep] != -1
    n_present = present.sum().item()
    n_absent = present.numel() - n_present
    return train_ids, val_ids, keep, torch.tensor(n_absent / (n_present + 1e-9))


def prepare_3class(cache, seed: int = 0):
    """train_3class_direct.py:75-131: inverse-frequency weights of {N/A, False, True}, normalised to sum to 3."""
    train_ids, val_ids = split_episodes(cache, seed)
    Y_tr = stack_labels(cache, train_ids)
    keep = _freq_keep(Y_tr)
    Yk = Y_tr[:, keep]
    total = Yk.numel()
    if total == 0:
        w = torch.tensor([1.0, 1.0, 1.0])
    else:
        counts = [(Yk == -1).sum().item(), (Yk == 0).sum().item(), (Yk == 1).sum().item()]
        w = torch.tensor([total / (3 * (c + 1e-6)) for c in counts], dtype=torch.float32)
        w = w / w.sum() * 3
    return train_ids, val_ids, keep, w


class _IndexDS(Dataset):
    """Stands in for the scripts' StepDS (train_object_probes.py:129-145): yields the sample index, so that a DataLoader
    built with the script's arguments draws from the global torch RNG exactly as the script's loader does."""

    def __init__(self, n):
        self.n = n

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        return i


def layer_samples(cache, ids, layer, dual_rule: bool = False):
    """The (episode, step) enumeration of StepDS, as dense X fp32 [N, D] / Y [N, n_labels] in that order."""
    xs, ys = [], []
    for i in ids:
        enc = cache[i]["visual_semantic_encoding"]
        if layer not in enc:
            continue
        x = enc[layer].float()
        y = torch.cat([cache[i]["symbolic_state_object_relations"], cache[i]["symbolic_state_action_subgoals"]], 1)
        n = min(y.shape[0], x.shape[0]) if dual_rule else x.shape[0]   # train_dual_head_final.py:138 vs train_object_probes.py:136
        xs.append(x[:n])
        ys.append(y[:n])
    if not xs:
        return torch.empty(0, 0), torch.empty(0, 0, dtype=torch.int8)
    return torch.cat(xs), torch.cat(ys)


class DualHeadProbe(nn.Module):
    """train_dual_head_final.py:147-153."""

    def __init__(self, input_dim, num_labels):
        super().__init__()
        self.presence_head = nn.Linear(input_dim, num_labels)
        self.truth_head = nn.Linear(input_dim, num_labels)

    def forward(self, x):
        return self.presence_head(x), self.truth_head(x)


def loss_fn(kind, model, x, y, pos_w):
    """The scripts' training losses: train_object_probes.py:184-188, train_spatial_probes.py:153,161-163,
    train_dual_head_final.py:175-191, train_3class_direct.py:155,182-190.  y already restricted to the kept columns."""
    if kind == OBJECT:
        logits = model(x)
        mask = y != -1
        target = (y == 1).float()
        bce = nn.BCEWithLogitsLoss(reduction="none", pos_weight=pos_w)
        return (bce(logits, target) * mask.float()).sum() / mask.sum(), logits
    if kind == SPATIAL:
        logits = model(x)
        return nn.BCEWithLogitsLoss(pos_weight=pos_w)(logits, y.float()), logits
    if kind == DUAL:
        pl, tl = model(x)
        presence_target = (y != -1).float()
        truth_target = (y == 1).float()
        truth_mask = y != -1
        loss_p = nn.BCEWithLogitsLoss(reduction="mean", pos_weight=pos_w)(pl, presence_target)
        el = nn.BCEWithLogitsLoss(reduction="none")(tl, truth_target)
        cnt = truth_mask.sum().item()
        loss_t = (el * truth_mask.float()).sum() / cnt if cnt > 0 else torch.tensor(0.0)
        return loss_p + loss_t, torch.cat([pl, tl], 1)
    logits = model(x)
    tgt = torch.zeros_like(y, dtype=torch.long)
    tgt[y == -1] = 0
    tgt[y == 0] = 1
    tgt[y == 1] = 2
    return nn.CrossEntropyLoss(weight=pos_w)(logits.view(-1, 3), tgt.view(-1)), logits


def val_metrics(kind, model, Xva, Yva_kept, batch, thresh=0.5):
    """Validation metrics as the scripts compute them with sklearn (train_object_probes.py:190-206,
    train_spatial_probes.py:165-176, train_dual_head_final.py:194-232, train_3class_direct.py:192-215)."""
    from sklearn.metrics import average_precision_score, f1_score

    with torch.no_grad():
        outs = [model(Xva[i:i + batch]) for i in range(0, Xva.shape[0], batch)]
    y = Yva_kept
    if kind == DUAL:
        pl = torch.cat([o[0] for o in outs])
        tl = torch.cat([o[1] for o in outs])
        pt, tt, tm = (y != -1).long(), (y == 1).long(), (y != -1)
        pp, tp = (pl.sigmoid() > 0.5).long(), (tl.sigmoid() > 0.5).long()
        out = dict(pres_acc_va=(pp == pt).sum().item() / pt.numel(),
                   truth_acc_va=((tp == tt) & tm).sum().item() / tm.sum().item() if tm.any() else 0.0,
                   pres_f1_va=f1_score(pt.view(-1).numpy(), pp.view(-1).numpy(), average="binary", pos_label=1, zero_division=0),
                   truth_f1_va=f1_score(tt[tm].numpy(), tp[tm].numpy(), labels=[0, 1], average="macro", zero_division=0)
                   if tm.any() else 0.0)
        return out
    z = torch.cat(outs)
    if kind == THREE:
        tgt = (y.long() + 1).view(-1)
        pred = z.view(-1, 3).argmax(1)
        return dict(val_acc=(pred == tgt).sum().item() / tgt.numel(),
                    val_f1=f1_score(tgt.numpy(), pred.numpy(), labels=[0, 1, 2], average="macro", zero_division=0))
    probs = z.sigmoid()
    if kind == OBJECT:
        mask, target = (y != -1), (y == 1).float()
        pred = (probs > thresh).long()
        ok = (pred[mask] == target[mask]).sum().item()
        tot = mask.sum().item()
        yt, yp, pr = target[mask].numpy(), pred[mask].numpy(), probs[mask].numpy()
    else:
        target = y.float()
        pred = (probs > thresh).float()
        ok, tot = (pred == target).sum().item(), target.numel()
        yt, yp, pr = target.numpy(), pred.numpy(), probs.numpy()
    return dict(val_acc=ok / tot, val_f1=f1_score(yt, yp, average="macro", zero_division=0),
                val_ap=average_precision_score(yt, pr, average="macro"))


def train_layers(kind: str, cache: Dict[int, dict], layers: Sequence[int], epochs: int, batch: int,
                 all_label_mats=None, seed: int = 0, record: bool = True) -> dict:
    """The per-layer loop of the scripts (train_object_probes.py:208-232, train_spatial_probes.py:179-204,
    train_dual_head_final.py:236-290, train_3class_direct.py:217-259) with the scripts' RNG consumption order:
    per layer nn.Linear init(s) -> per epoch one train DataLoader iterator (+ one val iterator per epoch for the dual /
    3-class scripts, one after the last epoch for object / spatial).  The dual / 3-class scripts seed torch themselves
    (`torch.manual_seed(args.seed)`, train_dual_head_final.py:40-45); for object / spatial the caller seeds."""
    if kind == OBJECT:
        train_ids, val_ids, keep, pos_w = prepare_object(cache)
    elif kind == SPATIAL:
        train_ids, val_ids, keep, pos_w = prepare_spatial(cache, all_label_mats)
    elif kind == DUAL:
        random.seed(seed); np.random.seed(seed); torch.manual_seed(seed)
        torch.Generator().manual_seed(seed)
        train_ids, val_ids, keep, pos_w = prepare_dual(cache, seed)
    else:
        random.seed(seed); np.random.seed(seed); torch.manual_seed(seed)
        torch.Generator().manual_seed(seed)
        train_ids, val_ids, keep, pos_w = prepare_3class(cache, seed)
    per_epoch_val = kind in (DUAL, THREE)
    drop_last = kind in (DUAL, THREE)
    out = dict(kind=kind, train_ids=train_ids, val_ids=val_ids, keep=keep, pos_weight=pos_w, layers={})
    for L in layers:
        Xtr, Ytr = layer_samples(cache, train_ids, L, dual_rule=per_epoch_val)
        Xva, Yva = layer_samples(cache, val_ids, L, dual_rule=per_epoch_val)
        if Xtr.shape[0] == 0 or Xva.shape[0] == 0:
            continue
        Ytr_k, Yva_k = Ytr[:, keep], Yva[:, keep]
        dl_tr = DataLoader(_IndexDS(Xtr.shape[0]), batch_size=batch, shuffle=True, drop_last=drop_last)
        dl_va = DataLoader(_IndexDS(Xva.shape[0]), batch_size=batch, shuffle=False)
        D, K = Xtr.shape[1], len(keep)
        if kind == DUAL:
            probe = DualHeadProbe(D, K)
        elif kind == THREE:
            probe = nn.Linear(D, 3 * K)
        else:
            probe = nn.Linear(D, K)
        opt = optim.AdamW(probe.parameters(), lr=1e-3, weight_decay=1e-4)
        rec = dict(init={k: v.detach().clone() for k, v in probe.state_dict().items()}, orders=[], losses=[],
                   n_train=Xtr.shape[0], n_val=Xva.shape[0])
        metrics = None
        for _ in range(epochs):
            probe.train(True)
            ep_orders = []
            for idx in dl_tr:
                x, y = Xtr[idx], Ytr_k[idx]
                loss, _ = loss_fn(kind, probe, x, y, pos_w)
                opt.zero_grad()
                loss.backward()
                opt.step()
                if record:
                    ep_orders.append(idx.clone())
                    rec["losses"].append(float(loss.detach()))
            rec["orders"].append(ep_orders)
            if per_epoch_val:
                for _ in dl_va:          # the script's per-epoch validation pass draws one loader seed from the global RNG
                    pass
        probe.train(False)
        if not per_epoch_val:
            for _ in dl_va:
                pass
        metrics = val_metrics(kind, probe, Xva, Yva_k, batch)
        rec["final"] = {k: v.detach().clone() for k, v in probe.state_dict().items()}
        rec["metrics"] = metrics
        out["layers"][L] = rec
    return out
