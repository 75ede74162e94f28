"""Probe training on the captured 4096-d layer features -- B200-native counterpart of the reference's
`experiment_utils/train_object_probes.py`, `train_spatial_probes.py` and `train_dual_head_final.py`.

Same data contract (`episode_*.pt` dicts), same split / label filtering / pos_weight rules, same
loss definitions, AdamW hyper-parameters and output files (`linear_probe_L{L:02d}.pth`, `*_dual_head_final_L{L:02d}.pth`,
CSV).  What changes is the execution: the whole layer dataset stays resident in HBM (fp32 [N, D] + a per-epoch
shuffled row-major / transposed copy), and one optimisation step is five kernels behind the C ABI:

    logits = X_b W^T + b        tcgen05 TF32 GEMM (fp32 accumulate)          ovla_gemm(kind=TF32)
    dZ^T, loss, counts          fused BCE-with-logits gradient, transposed    ovla_probe_bce_grad
    db                          row sums of dZ^T                              ovla_probe_rowsum
    dW = dZ^T X_b               tcgen05 TF32 GEMM on the transposed copies    ovla_gemm(kind=TF32)
    AdamW                       fused update, normalisers read on device      ovla_probe_adamw

Precision: the reference runs true-fp32 matmuls (torch default allow_tf32=False); here the two GEMMs multiply in TF32
(10-bit mantissa) with fp32 accumulation -- tests/test_gpu_probes.py states and checks the tolerance (rel-L2 <= 2e-3 on
logits / dW per step against fp32 autograd).

Multi-GPU (`torch.distributed`, NCCL over NVLink): every global batch is split evenly over the ranks; gradients are
kept UN-normalised, so ONE allreduce(sum) of the flat [dW | db | loss, count] buffer per step yields the global
gradient and the global `mask.sum()` normaliser together; every rank then applies the identical AdamW update.
"""
from __future__ import annotations

import argparse
import ast
import ctypes as C
import glob
import os
import random
from dataclasses import dataclass
from pathlib import Path
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

KIND_OBJECT, KIND_SPATIAL, KIND_DUAL, KIND_3CLASS = "object", "spatial", "dual", "3class"
_KIND0 = {KIND_OBJECT: 0, KIND_SPATIAL: 1, KIND_DUAL: 2}


# ----------------------------------------------------------------------------------------------- episode files
class EpisodeWriter:
    """Accumulates one episode exactly as run_libero_eval_object.py:254-256,309-312,357-366 does and writes
    `episode_{n}.pt` = {"visual_semantic_encoding": {L: float32 [T, D]}, "symbolic_state_object_relations": int8
    [T, n_rel], "symbolic_state_action_subgoals": int8 [T, n_act]} (dict keys are Python ints)."""

    def __init__(self, layers: Sequence[int] = tuple(range(33))):
        self.layers = list(layers)
        self.embeds: Dict[int, List[np.ndarray]] = {L: [] for L in self.layers}
        self.rel: List[np.ndarray] = []
        self.act: List[np.ndarray] = []

    def append(self, embeds: Dict[int, np.ndarray], rel_vec: np.ndarray, act_vec: np.ndarray) -> None:
        assert set(embeds) == set(self.layers)                          # run_libero_eval_object.py:292
        for L in self.layers:
            self.embeds[L].append(np.asarray(embeds[L], dtype=np.float32))
        rel_vec, act_vec = np.asarray(rel_vec, dtype=np.int8), np.asarray(act_vec, dtype=np.int8)
        assert set(np.unique(rel_vec)).issubset({-1, 0, 1}) and set(np.unique(act_vec)).issubset({-1, 0, 1})
        self.rel.append(rel_vec)
        self.act.append(act_vec)

    def append_batch(self, pooled: np.ndarray, rel: np.ndarray, act: np.ndarray) -> None:
        """pooled fp32 [n_layers, B, D] straight from the engine; rel/act int8 [B, *]."""
        for b in range(pooled.shape[1]):
            self.append({L: pooled[L, b] for L in self.layers}, rel[b], act[b])

    def save(self, path: str) -> None:
        d = {
            "visual_semantic_encoding": {L: torch.from_numpy(np.stack(v)) for L, v in self.embeds.items()},
            "symbolic_state_object_relations": torch.from_numpy(np.stack(self.rel)),
            "symbolic_state_action_subgoals": torch.from_numpy(np.stack(self.act)),
        }
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        torch.save(d, path)


def load_episodes(log_dir: str, exclude: Sequence[int] = ()) -> Dict[int, dict]:
    """train_object_probes.py:59-69."""
    files = sorted(glob.glob(os.path.join(log_dir, "episode_*.pt")))
    files = [fp for fp in files if int(Path(fp).stem.split("_")[1]) not in set(exclude)]
    if not files:
        raise FileNotFoundError("No episode_*.pt after applying exclusions")
    return {i: torch.load(fp, map_location="cpu", weights_only=False) for i, fp in enumerate(files)}


def parse_exclusions(spec: str) -> set:
    """train_object_probes.py:35-45."""
    out = set()
    if spec.strip():
        for tok in spec.split(","):
            tok = tok.strip()
            if "-" in tok:
                a, b = map(int, tok.split("-"))
                out.update(range(a, b + 1))
            else:
                out.add(int(tok))
    return out


# ----------------------------------------------------------------------------------------------- dataset prep (host)
@dataclass
class ProbeSplit:
    train_ids: List[int]
    val_ids: List[int]
    keep: torch.Tensor            # kept label columns (int64)
    pos_weight: torch.Tensor      # vector [K] (object / spatial) or scalar tensor (dual)
    num_labels: int


def _stack_labels(cache, ids) -> torch.Tensor:
    return torch.cat([torch.cat([cache[i]["symbolic_state_object_relations"],
                                 cache[i]["symbolic_state_action_subgoals"]], 1) for i in ids], 0)


def split_episodes(cache: Dict[int, dict], seed: int = 0) -> Tuple[List[int], List[int]]:
    """Episode-level 90/10 split with random.Random(seed).shuffle (train_object_probes.py:50,72-75)."""
    rng = random.Random(seed)
    ep_ids = list(cache.keys())
    rng.shuffle(ep_ids)
    val_len = max(1, int(0.10 * len(ep_ids)))
    return ep_ids[val_len:], ep_ids[:val_len]


def prepare_object(cache: Dict[int, dict]) -> ProbeSplit:
    """train_object_probes.py:72-102: keep = labels that take both 0 and 1 somewhere (train U val);
    pos_weight = clamp((neg+1)/(pos+1), max=20) on the TRAIN split."""
    train_ids, val_ids = split_episodes(cache)
    Y_full = _stack_labels(cache, train_ids + val_ids)
    mask_full = Y_full != -1
    keep = (((Y_full == 1) & mask_full).any(0) & ((Y_full == 0) & mask_full).any(0)).nonzero(as_tuple=True)[0]
    if len(keep) == 0:
        raise RuntimeError("No label flips value across remaining episodes.")
    Y_tr = _stack_labels(cache, train_ids)
    m = Y_tr != -1
    pos = ((Y_tr == 1) & m).sum(0).float()
    neg = ((Y_tr == 0) & m).sum(0).float()
    pw = ((neg + 1.0) / (pos + 1.0))[keep].clamp(max=20)
    return ProbeSplit(train_ids, val_ids, keep, pw, Y_full.shape[1])


def prepare_spatial(cache: Dict[int, dict], all_label_mats: Optional[Sequence[torch.Tensor]] = None) -> ProbeSplit:
    """train_spatial_probes.py:96-131.  keep = columns that show both a 0 and a 1 over ALL episode files -- the script
    stacks `all_files`, i.e. also the rollouts dropped by --exclude_eps (pass their label matrices as `all_label_mats`;
    default: the cached episodes); pos_cnt = Y_train.sum(0) (labels assumed {0, 1}: the script does not mask -1, SURVEY
    Appendix B), neg_cnt = N_train - pos_cnt, pos_weight = clamp((neg + 1) / (pos + 1), max=20)[keep]."""
    train_ids, val_ids = split_episodes(cache)
    Y_full = torch.cat(list(all_label_mats), 0) if all_label_mats is not None else _stack_labels(cache, sorted(cache))
    keep = ((Y_full == 1).any(0) & (Y_full == 0).any(0)).nonzero(as_tuple=True)[0]
    if len(keep) == 0:
        raise RuntimeError("Even across all episodes no label flips value.")
    Y_tr = _stack_labels(cache, train_ids)
    pos = Y_tr.sum(0).float()
    neg = Y_tr.shape[0] - pos
    pw = ((neg + 1.0) / (pos + 1.0))[keep].clamp(max=20)
    return ProbeSplit(train_ids, val_ids, keep, pw, Y_tr.shape[1])


def prepare_dual(cache: Dict[int, dict], seed: int = 0) -> ProbeSplit:
    """train_dual_head_final.py:100-127: keep labels whose TRAIN 0/1 frequency is in (1 %, 99 %);
    presence pos_weight = #absent / (#present + 1e-9) over the kept training labels (a scalar)."""
    train_ids, val_ids = split_episodes(cache, seed)
    Y_tr = _stack_labels(cache, train_ids)
    m = Y_tr != -1
    cnt = m.sum(0)
    freq = torch.full((Y_tr.shape[1],), -1.0)
    ok = cnt > 0
    freq[ok] = ((Y_tr == 1) & m).sum(0)[ok].float() / cnt[ok]
    keep = ((freq > 0.01) & (freq < 0.99)).nonzero(as_tuple=True)[0]
    if len(keep) == 0:
        keep = torch.arange(Y_tr.shape[1])
    present = (Y_tr[:, keep] != -1)
    n_present = int(present.sum())
    pw = torch.tensor((present.numel() - n_present) / (n_present + 1e-9))
    return ProbeSplit(train_ids, val_ids, keep, pw, Y_tr.shape[1])


def prepare_3class(cache: Dict[int, dict], seed: int = 0) -> ProbeSplit:
    """train_3class_direct.py:82-131: same 1 %-99 % keep rule as the dual-head script; inverse-frequency class weights
    for {N/A, False, True} over the kept TRAIN labels, normalised to sum to 3 (`pos_weight` holds the 3 weights)."""
    train_ids, val_ids = split_episodes(cache, seed)
    Y_tr = _stack_labels(cache, train_ids)
    m = Y_tr != -1
    cnt = m.sum(0)
    freq = torch.full((Y_tr.shape[1],), -1.0)
    ok = cnt > 0
    freq[ok] = ((Y_tr == 1) & m).sum(0)[ok].float() / cnt[ok]
    keep = ((freq > 0.01) & (freq < 0.99)).nonzero(as_tuple=True)[0]
    if len(keep) == 0:
        keep = torch.arange(Y_tr.shape[1])
    Yk = Y_tr[:, keep]
    total = Yk.numel()
    if total == 0:
        w = torch.ones(3)
    else:
        w = torch.tensor([total / (3 * (float((Yk == v).sum()) + 1e-6)) for v in (-1, 0, 1)], dtype=torch.float32)
        w = w / w.sum() * 3
    return ProbeSplit(train_ids, val_ids, keep, w, Y_tr.shape[1])


def layer_matrix(cache: Dict[int, dict], ids: Sequence[int], layer: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """The samples StepDS enumerates (train_object_probes.py:129-145), as dense tensors: X fp32 [N, D], Y int8 [N, L]."""
    xs, ys = [], []
    for i in ids:
        enc = cache[i]["visual_semantic_encoding"]
        if layer not in enc:
            continue
        x = enc[layer].float()
        y = torch.cat([cache[i]["symbolic_state_object_relations"], cache[i]["symbolic_state_action_subgoals"]], 1)
        n = min(x.shape[0], y.shape[0])
        xs.append(x[:n])
        ys.append(y[:n].to(torch.int8))
    if not xs:
        return torch.empty(0, 0), torch.empty(0, 0, dtype=torch.int8)
    return torch.cat(xs), torch.cat(ys)


# ----------------------------------------------------------------------------------------------- sharding (host logic)
def shard_batches(n: int, batch: int, world: int, rank: int, drop_last: bool) -> List[Tuple[int, int]]:
    """Row ranges [lo, hi) of the shuffled order that `rank` processes, one per optimisation step: every global
    batch [s*batch, min((s+1)*batch, n)) is cut into `world` near-equal contiguous pieces (sizes differ by <= 1)."""
    out = []
    n_steps = n // batch if drop_last else (n + batch - 1) // batch
    for s in range(n_steps):
        lo, hi = s * batch, min((s + 1) * batch, n)
        m = hi - lo
        base, extra = divmod(m, world)
        a = lo + rank * base + min(rank, extra)
        b = a + base + (1 if rank < extra else 0)
        out.append((a, b))
    return out


def allreduce_flat(buf: torch.Tensor, group=None) -> None:
    """One collective per step: [dW | db | loss_h0, count_h0, loss_h1, count_h1] summed over ranks."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)


def adamw_reference_step(p, g, m, v, step, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, wd=1e-4):
    """torch.optim.AdamW update rule in plain tensor ops (host-side twin of ovla_probe_adamw, used by CPU tests)."""
    p.mul_(1 - lr * wd)
    m.lerp_(g, 1 - betas[0])
    v.mul_(betas[1]).addcmul_(g, g, value=1 - betas[1])
    bc1, bc2 = 1 - betas[0] ** step, 1 - betas[1] ** step
    p.addcdiv_(m, (v.sqrt() / (bc2 ** 0.5)).add_(eps), value=-lr / bc1)


# ----------------------------------------------------------------------------------------------- device trainer
class ProbeTrainer:
    """One probe (object / spatial: a single nn.Linear(D, K); dual: presence + truth heads) trained on one layer."""

    def __init__(self, kind: str, D: int, K: int, pos_weight: torch.Tensor, batch: int = 4096, lr: float = 1e-3,
                 weight_decay: float = 1e-4, device: int = 0, group=None, init_state: Optional[Dict[str, torch.Tensor]] = None):
        from . import _lib

        self._lib_mod = _lib
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.OvlaError("probe training needs a CUDA device (sm_100a); there is no CPU fallback")
        if D % 4:
            raise ValueError("feature dim must be a multiple of 4")
        self.kind, self.D, self.K = kind, D, K
        self.heads = 2 if kind == KIND_DUAL else 1
        self.Kpad = (K + 7) // 8 * 8                      # label columns of the gathered int8 label matrix
        self.n_out = 3 * K if kind == KIND_3CLASS else K  # real output rows per head
        self.rows_per_head = (3 * K + 7) // 8 * 8 if kind == KIND_3CLASS else self.Kpad
        self.rows = self.heads * self.rows_per_head
        self.batch, self.lr, self.wd = batch, lr, weight_decay
        self.dev = torch.device("cuda", device)
        self.group = group
        import torch.distributed as dist

        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        n_w = self.rows * D
        self.n_w, self.n_total = n_w, n_w + self.rows
        self.P = torch.zeros(self.n_total, dtype=torch.float32, device=self.dev)
        self.G = torch.zeros(self.n_total + 4, dtype=torch.float32, device=self.dev)     # [dW | db | stats(4)]
        self.M = torch.zeros_like(self.P)
        self.V = torch.zeros_like(self.P)
        self.step_count = 0
        self.class_w = None
        if kind == KIND_3CLASS:
            self.pw_vec, self.pw_scalar = None, 1.0
            self.class_w = (C.c_float * 3)(*[float(v) for v in pos_weight])
        elif kind == KIND_DUAL:
            self.pw_vec, self.pw_scalar = None, float(pos_weight)
        else:
            pw = torch.ones(self.Kpad, dtype=torch.float32)
            pw[:K] = pos_weight.float()
            self.pw_vec, self.pw_scalar = pw.to(self.dev), 1.0
        self._init_params(init_state)
        self.last_loss = None

    # -- parameters ---------------------------------------------------------------------------------
    def _init_params(self, init_state):
        """nn.Linear default init (kaiming-uniform a=sqrt(5)), drawn with torch on the host exactly as the reference's
        `nn.Linear(4096, K)` would (train_object_probes.py:219), or a given state dict."""
        W = self.P[: self.n_w].view(self.rows, self.D)
        b = self.P[self.n_w:]
        names = ["presence_head", "truth_head"] if self.heads == 2 else [None]
        for h, nm in enumerate(names):
            if init_state is None:
                lin = torch.nn.Linear(self.D, self.n_out)
                w0, b0 = lin.weight.detach(), lin.bias.detach()
            else:
                pre = f"{nm}." if nm else ""
                w0, b0 = init_state[pre + "weight"].float(), init_state[pre + "bias"].float()
            W[h * self.rows_per_head: h * self.rows_per_head + self.n_out].copy_(w0)
            b[h * self.rows_per_head: h * self.rows_per_head + self.n_out].copy_(b0)

    def state_dict(self) -> Dict[str, torch.Tensor]:
        W = self.P[: self.n_w].view(self.rows, self.D)
        b = self.P[self.n_w:]
        if self.heads == 1:
            return {"weight": W[: self.n_out].cpu().clone(), "bias": b[: self.n_out].cpu().clone()}
        return {"presence_head.weight": W[: self.K].cpu().clone(), "presence_head.bias": b[: self.K].cpu().clone(),
                "truth_head.weight": W[self.Kpad: self.Kpad + self.K].cpu().clone(),
                "truth_head.bias": b[self.Kpad: self.Kpad + self.K].cpu().clone()}

    # -- native calls --------------------------------------------------------------------------------
    def _gemm_tf32(self, A, lda, Wt, ldw, M, N, K, out, ldo, bias=None):
        epi = self._lib_mod.GemmEpilogue()
        if bias is not None:
            epi.bias_f32 = bias.data_ptr()
        self._lib_mod.check(self.lib.ovla_gemm(C.c_void_p(A), C.c_longlong(lda), C.c_void_p(Wt), C.c_longlong(ldw),
                                               M, N, K, 2, 1, C.c_void_p(out), C.c_longlong(ldo), C.byref(epi), 0, 0,
                                               self._lib_mod.stream_ptr()))

    def logits(self, X: torch.Tensor) -> torch.Tensor:
        """X fp32 [n, D] on the device -> logits fp32 [n, heads*Kpad]."""
        n = X.shape[0]
        Z = torch.empty(n, self.rows, dtype=torch.float32, device=self.dev)
        if n:
            self._gemm_tf32(X.data_ptr(), self.D, self.P.data_ptr(), self.D, n, self.rows, self.D, Z.data_ptr(), self.rows,
                            bias=self.P[self.n_w:])
        return Z

    def load_epoch(self, X: torch.Tensor, Y: torch.Tensor, keep: torch.Tensor, perm: torch.Tensor, drop_last: bool):
        """Gather this rank's rows of the shuffled epoch into contiguous row-major + transposed copies."""
        n = X.shape[0]
        ranges = shard_batches(n, self.batch, self.world, self.rank, drop_last)
        idx = torch.cat([perm[a:b] for a, b in ranges]) if ranges else perm[:0]
        n_loc = idx.numel()
        ldt = (n_loc + 3) // 4 * 4
        self.Xp = torch.empty(max(n_loc, 1), self.D, dtype=torch.float32, device=self.dev)
        self.XpT = torch.zeros(self.D, max(ldt, 4), dtype=torch.float32, device=self.dev)
        self.Yp = torch.empty(max(n_loc, 1), self.Kpad, dtype=torch.int8, device=self.dev)
        idx_d = idx.to(self.dev, torch.int64)
        keep_d = keep.to(self.dev, torch.int32)
        st = self._lib_mod.stream_ptr()
        ck = self._lib_mod.check
        if n_loc:
            ck(self.lib.ovla_probe_gather(C.c_void_p(X.data_ptr()), C.c_longlong(X.stride(0)), C.c_void_p(idx_d.data_ptr()),
                                          n_loc, self.D, C.c_void_p(self.Xp.data_ptr()), C.c_longlong(self.D),
                                          C.c_void_p(self.XpT.data_ptr()), C.c_longlong(self.XpT.stride(0)), st))
            ck(self.lib.ovla_probe_gather_labels(C.c_void_p(Y.data_ptr()), C.c_longlong(Y.stride(0)),
                                                 C.c_void_p(idx_d.data_ptr()), C.c_void_p(keep_d.data_ptr()), n_loc,
                                                 self.K, self.Kpad, C.c_void_p(self.Yp.data_ptr()), st))
        # local [lo, hi) of every step inside the gathered copy
        self.steps, off = [], 0
        for a, b in ranges:
            self.steps.append((off, off + (b - a)))
            off += b - a
        bmax = max([hi - lo for lo, hi in self.steps], default=0)
        self.ldz_t = max((bmax + 3) // 4 * 4, 4)
        self.Z = torch.empty(max(bmax, 1), self.rows, dtype=torch.float32, device=self.dev)
        self.dZT = torch.zeros(self.rows, self.ldz_t, dtype=torch.float32, device=self.dev)

    def train_step(self, s: int) -> None:
        """One optimisation step on local rows self.steps[s] (all ranks call this in lock-step)."""
        lo, hi = self.steps[s]
        n = hi - lo
        st = self._lib_mod.stream_ptr()
        ck = self._lib_mod.check
        self.G.zero_()
        stats = self.G[self.n_total:]
        if n > 0:
            xp = self.Xp.data_ptr() + lo * self.D * 4
            self._gemm_tf32(xp, self.D, self.P.data_ptr(), self.D, n, self.rows, self.D, self.Z.data_ptr(), self.rows,
                            bias=self.P[self.n_w:])
            if self.kind == KIND_3CLASS:
                ck(self.lib.ovla_probe_ce3_grad(C.c_void_p(self.Z.data_ptr()), C.c_longlong(self.rows),
                                                C.c_void_p(self.Yp.data_ptr() + lo * self.Kpad), C.c_longlong(self.Kpad), n,
                                                self.K, self.rows, self.class_w, C.c_void_p(self.dZT.data_ptr()),
                                                C.c_longlong(self.ldz_t), C.c_void_p(stats.data_ptr()), st))
            else:
                ck(self.lib.ovla_probe_bce_grad(C.c_void_p(self.Z.data_ptr()), C.c_longlong(self.rows),
                                                C.c_void_p(self.Yp.data_ptr() + lo * self.Kpad), n, self.K, self.Kpad,
                                                _KIND0[self.kind], self.heads,
                                                C.c_void_p(self.pw_vec.data_ptr()) if self.pw_vec is not None else None,
                                                C.c_float(self.pw_scalar), C.c_void_p(self.dZT.data_ptr()),
                                                C.c_longlong(self.ldz_t), C.c_void_p(stats.data_ptr()), st))
            ck(self.lib.ovla_probe_rowsum(C.c_void_p(self.dZT.data_ptr()), C.c_longlong(self.ldz_t), self.rows, n,
                                          C.c_void_p(self.G.data_ptr() + self.n_w * 4), st))
            # dW[rows, D] = dZT[rows, n] . (XpT[:, lo:hi])^T      (contraction over the local batch)
            self._gemm_tf32(self.dZT.data_ptr(), self.ldz_t, self.XpT.data_ptr() + lo * 4, self.XpT.stride(0),
                            self.rows, self.D, n, self.G.data_ptr(), self.D)
        allreduce_flat(self.G, self.group)
        self.step_count += 1
        ck(self.lib.ovla_probe_adamw(C.c_void_p(self.P.data_ptr()), C.c_void_p(self.G.data_ptr()),
                                     C.c_void_p(self.M.data_ptr()), C.c_void_p(self.V.data_ptr()),
                                     C.c_longlong(self.n_w), self.D, self.rows_per_head, C.c_longlong(self.n_total),
                                     C.c_void_p(stats.data_ptr()), C.c_float(self.lr), C.c_float(0.9), C.c_float(0.999),
                                     C.c_float(1e-8), C.c_float(self.wd), self.step_count, st))

    def step_loss(self) -> float:
        """Loss of the last step (host sync): loss_h0/count_h0 (+ loss_h1/count_h1)."""
        s = self.G[self.n_total:].cpu()
        loss = float(s[0] / s[1]) if s[1] > 0 else 0.0
        if self.heads == 2 and s[3] > 0:
            loss += float(s[2] / s[3])
        return loss

    def fit(self, X: torch.Tensor, Y: torch.Tensor, keep: torch.Tensor, epochs: int, seed: int = 0,
            drop_last: bool = False, on_epoch: Optional[Callable[[int], None]] = None) -> None:
        """X fp32 [N, D], Y int8 [N, n_labels] (host or device); shuffles every epoch like DataLoader(shuffle=True)."""
        X = X.to(self.dev, torch.float32).contiguous()
        Y = Y.to(self.dev, torch.int8).contiguous()
        g = torch.Generator().manual_seed(seed)
        for e in range(epochs):
            perm = torch.randperm(X.shape[0], generator=g)
            self.load_epoch(X, Y, keep, perm, drop_last)
            for s in range(len(self.steps)):
                self.train_step(s)
            if on_epoch:
                on_epoch(e)


class MultiLayerProbeTrainer:
    """The probes of `G` captured layers trained CONCURRENTLY on the same shuffled batches (the reference loops over the
    layers one at a time: train_object_probes.py:208-232; every layer sees the same labels, only the features differ).

    One optimisation step of all G probes is four grouped launches (object / spatial / dual kinds):

        logits_g = X_g W_g^T + b_g     one persistent tcgen05 TF32 GEMM over G x tiles        ovla_gemm_grouped
        dZ^T_g, db_g, loss/count_g     fused BCE gradient + bias gradient + statistics,       ovla_probe_bce_grad_grouped
                                       deterministic (no floating-point atomics)
        dW_g = dZ^T_g X_g              grouped TF32 GEMM straight into the flat gradient      ovla_gemm_grouped
        AdamW of all G probes          one launch                                             ovla_probe_adamw_grouped

    A single probe has only 64 output tiles for 148 SMs; G x 64 tiles fill the machine.

    Multi-GPU: the flat gradient buffer [G][dW | db | stats] is cut into `chunks` layer ranges; the all-reduce of chunk c
    (NCCL, its own stream) runs while chunk c+1 computes, and AdamW of chunk c is issued one chunk later, so the
    collective is hidden behind compute instead of trailing every step.  With world > 1 the persistent GEMMs leave
    `comm_sms` SMs to the collective (a persistent grid that cannot get all its CTAs resident runs a second wave).
    Parameters / moments / gradients of layer g live at [g] of P / M / V / G (so `trainer(g)`-style views are free)."""

    def __init__(self, kind: str, G: int, D: int, K: int, pos_weight: torch.Tensor, batch: int = 4096, lr: float = 1e-3,
                 weight_decay: float = 1e-4, device: int = 0, group=None, init_states: Optional[Sequence[Dict[str, torch.Tensor]]] = None,
                 chunks: int = 0, comm_sms: int = 16, shard: str = "split"):
        from . import _lib

        if kind not in _KIND0:
            raise ValueError(f"grouped training covers object / spatial / dual probes, not {kind!r}")
        self._lib_mod, self.lib = _lib, _lib.load()
        if not torch.cuda.is_available():
            raise _lib.OvlaError("probe training needs a CUDA device (sm_100a); there is no CPU fallback")
        if D % 4:
            raise ValueError("feature dim must be a multiple of 4")
        import torch.distributed as dist

        self.kind, self.G, self.D, self.K = kind, G, D, K
        self.heads = 2 if kind == KIND_DUAL else 1
        self.Kpad = (K + 7) // 8 * 8
        self.n_out = K
        self.rows_per_head = self.Kpad
        self.rows = self.heads * self.Kpad
        self.batch, self.lr, self.wd = batch, lr, weight_decay
        self.dev = torch.device("cuda", device)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        # "split": every global batch of `batch` rows is cut over the ranks (the reference's trajectory at any N);
        # "whole": rank r takes global batches r, r + world, ... whole (global batch = world x batch);
        # "local": X / Y handed to load_epoch are already this rank's own rows (global batch = world x batch)
        if shard not in ("split", "whole", "local"):
            raise ValueError(f"unknown shard mode {shard!r}")
        self.shard = shard
        self.n_w = self.rows * D
        self.n_total = self.n_w + self.rows
        self.gs = self.n_total + 4                # floats between the layers' [dW | db | stats] blocks (16-byte multiple)
        self.P = torch.zeros(G, self.n_total, dtype=torch.float32, device=self.dev)
        self.M = torch.zeros_like(self.P)
        self.V = torch.zeros_like(self.P)
        self.Gbuf = torch.zeros(G, self.gs, dtype=torch.float32, device=self.dev)
        self.step_count = 0
        if kind == KIND_DUAL:
            self.pw_vec, self.pw_scalar = None, float(pos_weight)
        else:
            pw = torch.ones(self.Kpad, dtype=torch.float32)
            pw[:K] = pos_weight.float()
            self.pw_vec, self.pw_scalar = pw.to(self.dev), 1.0
        if chunks <= 0:
            chunks = 1 if self.world == 1 else min(G, 3)
        chunks = max(1, min(chunks, G))
        cuts = [round(i * G / chunks) for i in range(chunks + 1)]
        self.chunks = [(cuts[i], cuts[i + 1]) for i in range(chunks) if cuts[i + 1] > cuts[i]]
        self.sm_limit = 0
        if self.world > 1 and comm_sms > 0:
            n_sm = torch.cuda.get_device_properties(self.dev).multi_processor_count
            self.sm_limit = max(2, (n_sm - comm_sms) // 2 * 2)
        self.isplits = 8
        self.ktiles = (self.Kpad + 31) // 32
        self.part = torch.zeros(G * self.isplits * (self.rows + 4 * self.ktiles), dtype=torch.float32, device=self.dev)
        self.ticket = torch.zeros(G, dtype=torch.int32, device=self.dev)
        self._pending = None                      # (chunk, all-reduce work handle) whose AdamW has not been issued yet
        for g in range(G):
            self._init_params(g, init_states[g] if init_states is not None else None)

    def _init_params(self, g: int, init_state):
        """nn.Linear default init drawn on the host in layer order (as the per-layer loop of the reference would)."""
        W = self.P[g, : self.n_w].view(self.rows, self.D)
        b = self.P[g, self.n_w:]
        names = ["presence_head", "truth_head"] if self.heads == 2 else [None]
        for h, nm in enumerate(names):
            if init_state is None:
                lin = torch.nn.Linear(self.D, self.K)
                w0, b0 = lin.weight.detach(), lin.bias.detach()
            else:
                pre = f"{nm}." if nm else ""
                w0, b0 = init_state[pre + "weight"].float(), init_state[pre + "bias"].float()
            W[h * self.Kpad: h * self.Kpad + self.K].copy_(w0)
            b[h * self.Kpad: h * self.Kpad + self.K].copy_(b0)

    def state_dict(self, g: int) -> Dict[str, torch.Tensor]:
        W = self.P[g, : self.n_w].view(self.rows, self.D)
        b = self.P[g, self.n_w:]
        if self.heads == 1:
            return {"weight": W[: self.K].cpu().clone(), "bias": b[: self.K].cpu().clone()}
        return {"presence_head.weight": W[: self.K].cpu().clone(), "presence_head.bias": b[: self.K].cpu().clone(),
                "truth_head.weight": W[self.Kpad: self.Kpad + self.K].cpu().clone(),
                "truth_head.bias": b[self.Kpad: self.Kpad + self.K].cpu().clone()}

    # -- native calls --------------------------------------------------------------------------------
    def _grouped_gemm(self, A, lda, a_gs, W, ldw, w_gs, groups, M, N, K, out, ldo, out_gs, bias=0, bias_gs=0):
        self._lib_mod.check(self.lib.ovla_gemm_grouped(A, lda, a_gs, W, ldw, w_gs, groups, M, N, K, 1, out, ldo, out_gs,
                                                       bias or None, bias_gs, 0, 0, self.sm_limit, self._lib_mod.stream_ptr()))

    def logits(self, X: torch.Tensor) -> torch.Tensor:
        """X fp32 [G, n, D] on the device -> logits fp32 [G, n, rows]."""
        G, n = X.shape[0], X.shape[1]
        assert G == self.G and X.is_contiguous()
        Z = torch.empty(G, n, self.rows, dtype=torch.float32, device=self.dev)
        if n:
            self._grouped_gemm(X.data_ptr(), self.D, n * self.D, self.P.data_ptr(), self.D, self.n_total, G, n, self.rows,
                               self.D, Z.data_ptr(), self.rows, n * self.rows, self.P.data_ptr() + self.n_w * 4, self.n_total)
        return Z

    def load_epoch(self, X: torch.Tensor, Y: torch.Tensor, keep: torch.Tensor, perm: torch.Tensor, drop_last: bool):
        """X fp32 [G, N, D] (device), Y int8 [N, n_labels]: gather this rank's rows of the shuffled epoch into row-major
        and transposed copies; every step starts at a row offset that is a multiple of 4 (16-byte aligned TMA bases)."""
        n = X.shape[1]
        if self.shard == "split":
            ranges = shard_batches(n, self.batch, self.world, self.rank, drop_last)
        elif self.shard == "local":
            ranges = shard_batches(n, self.batch, 1, 0, drop_last)
        else:
            full = shard_batches(n, self.batch, 1, 0, drop_last)
            n_steps = len(full) // self.world
            ranges = [full[s * self.world + self.rank] for s in range(n_steps)]
        pieces, self.steps, off = [], [], 0
        for a, b in ranges:
            m = b - a
            pad = (-m) % 4
            pieces.append(perm[a:b])
            if pad:
                pieces.append(perm[:1].expand(pad))
            self.steps.append((off, off + m))
            off += m + pad
        idx = torch.cat(pieces) if pieces else perm[:0]
        n_loc = idx.numel()
        self.n_alloc = max(n_loc, 4)
        self.Xp = torch.empty(self.G, self.n_alloc, self.D, dtype=torch.float32, device=self.dev)
        self.XpT = torch.zeros(self.G, self.D, self.n_alloc, dtype=torch.float32, device=self.dev)
        self.Yp = torch.empty(self.n_alloc, self.Kpad, dtype=torch.int8, device=self.dev)
        idx_d = idx.to(self.dev, torch.int64)
        keep_d = keep.to(self.dev, torch.int32)
        st, ck = self._lib_mod.stream_ptr(), self._lib_mod.check
        if n_loc:
            for g in range(self.G):
                ck(self.lib.ovla_probe_gather(X[g].data_ptr(), X.stride(1), idx_d.data_ptr(), n_loc, self.D,
                                              self.Xp[g].data_ptr(), self.D, self.XpT[g].data_ptr(), self.n_alloc, st))
            ck(self.lib.ovla_probe_gather_labels(Y.data_ptr(), Y.stride(0), idx_d.data_ptr(), keep_d.data_ptr(), n_loc,
                                                 self.K, self.Kpad, self.Yp.data_ptr(), st))
        bmax = max([hi - lo for lo, hi in self.steps], default=0)
        self.bmax = max(bmax, 1)
        self.ldz_t = max((bmax + 3) // 4 * 4, 4)
        self.Z = torch.empty(self.G, self.bmax, self.rows, dtype=torch.float32, device=self.dev)
        self.dZT = torch.zeros(self.G, self.rows, self.ldz_t, dtype=torch.float32, device=self.dev)

    def _compute_chunk(self, g0: int, g1: int, lo: int, n: int) -> None:
        ng = g1 - g0
        f4 = 4
        if n <= 0:
            self.Gbuf[g0:g1].zero_()
            return
        P0 = self.P.data_ptr() + g0 * self.n_total * f4
        Z0 = self.Z.data_ptr() + g0 * self.bmax * self.rows * f4
        T0 = self.dZT.data_ptr() + g0 * self.rows * self.ldz_t * f4
        G0 = self.Gbuf.data_ptr() + g0 * self.gs * f4
        self._grouped_gemm(self.Xp.data_ptr() + (g0 * self.n_alloc + lo) * self.D * f4, self.D, self.n_alloc * self.D,
                           P0, self.D, self.n_total, ng, n, self.rows, self.D, Z0, self.rows, self.bmax * self.rows,
                           P0 + self.n_w * f4, self.n_total)
        self._lib_mod.check(self.lib.ovla_probe_bce_grad_grouped(
            Z0, self.rows, self.bmax * self.rows, self.Yp.data_ptr() + lo * self.Kpad, n, self.K, self.Kpad,
            _KIND0[self.kind], self.heads, self.pw_vec.data_ptr() if self.pw_vec is not None else None, self.pw_scalar,
            T0, self.ldz_t, self.rows * self.ldz_t, ng, G0, self.gs, self.n_w, self.n_total,
            self.part.data_ptr() + g0 * self.isplits * (self.rows + 4 * self.ktiles) * f4, self.isplits,
            self.ticket.data_ptr() + g0 * 4, self._lib_mod.stream_ptr()))
        # dW_g[rows, D] = dZT_g[rows, n] . (XpT_g[:, lo:lo+n])^T
        self._grouped_gemm(T0, self.ldz_t, self.rows * self.ldz_t,
                           self.XpT.data_ptr() + (g0 * self.D * self.n_alloc + lo) * f4, self.n_alloc, self.D * self.n_alloc,
                           ng, self.rows, self.D, n, G0, self.D, self.gs)

    def _adamw_chunk(self, g0: int, g1: int, step: int) -> None:
        f4 = 4
        off = g0 * self.n_total * f4
        G0 = self.Gbuf.data_ptr() + g0 * self.gs * f4
        self._lib_mod.check(self.lib.ovla_probe_adamw_grouped(
            self.P.data_ptr() + off, G0, self.M.data_ptr() + off, self.V.data_ptr() + off, g1 - g0, self.n_w, self.D,
            self.rows_per_head, self.n_total, self.gs, G0 + self.n_total * f4, self.gs, self.lr, 0.9, 0.999, 1e-8, self.wd,
            step, self._lib_mod.stream_ptr()))

    def _flush_pending(self) -> None:
        if self._pending is not None:
            (g0, g1), work, step = self._pending
            if work is not None:
                work.wait()                       # the compute stream waits for the collective, not the host
            self._adamw_chunk(g0, g1, step)
            self._pending = None

    def train_step(self, s: int) -> None:
        """One optimisation step of all G probes on local rows self.steps[s] (all ranks in lock-step)."""
        import torch.distributed as dist

        lo, hi = self.steps[s]
        self.step_count += 1
        for (g0, g1) in self.chunks:
            self._compute_chunk(g0, g1, lo, hi - lo)
            work = None
            if self.world > 1:
                work = dist.all_reduce(self.Gbuf[g0:g1], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            self._flush_pending()                 # AdamW of the PREVIOUS chunk, whose all-reduce ran beside this compute
            self._pending = ((g0, g1), work, self.step_count)
            if len(self.chunks) == 1:
                self._flush_pending()             # a single chunk has nothing to hide behind

    def finish(self) -> None:
        """Issue the AdamW update still in flight (call before reading parameters or statistics)."""
        self._flush_pending()

    def step_losses(self) -> List[float]:
        """Loss of the last step for every layer (host sync)."""
        self.finish()
        st = self.Gbuf[:, self.n_total:].cpu()
        out = []
        for g in range(self.G):
            loss = float(st[g, 0] / st[g, 1]) if st[g, 1] > 0 else 0.0
            if self.heads == 2 and st[g, 3] > 0:
                loss += float(st[g, 2] / st[g, 3])
            out.append(loss)
        return out

    def fit(self, X: torch.Tensor, Y: torch.Tensor, keep: torch.Tensor, epochs: int, seed: int = 0,
            drop_last: bool = False, on_epoch: Optional[Callable[[int], None]] = None) -> None:
        X = X.to(self.dev, torch.float32).contiguous()
        Y = Y.to(self.dev, torch.int8).contiguous()
        g = torch.Generator().manual_seed(seed)
        for e in range(epochs):
            perm = torch.randperm(X.shape[1], generator=g)
            self.load_epoch(X, Y, keep, perm, drop_last)
            for s in range(len(self.steps)):
                self.train_step(s)
            self.finish()
            if on_epoch:
                on_epoch(e)


# ----------------------------------------------------------------------------------------------- evaluation (host, as reference)
def _f1_binary_macro(tp: int, fp: int, fn: int, tn: int) -> float:
    """sklearn.metrics.f1_score(y_true, y_pred, average="macro", zero_division=0) from the confusion counts: per-class F1
    over the classes that occur in y_true or y_pred (sklearn's default `labels`), undefined F1 counted as 0."""
    f = []
    if tp + fn + fp > 0:                                  # class 1 occurs in truth or prediction
        f.append(2 * tp / (2 * tp + fp + fn))
    if tn + fp + fn > 0:                                  # class 0 occurs
        f.append(2 * tn / (2 * tn + fn + fp))
    return float(sum(f) / len(f)) if f else 0.0


def metrics_from_counts(kind: str, c: Sequence[int]) -> Dict[str, float]:
    """Validation accuracy / F1 of the reference scripts from on-device confusion counts (ovla_probe_confusion)."""
    c = [int(v) for v in c]
    if kind == KIND_3CLASS:                               # train_3class_direct.py:196-207, labels=[0,1,2] macro
        total = sum(c)
        acc = (c[0] + c[4] + c[8]) / total if total else 0.0
        f = []
        for k in range(3):
            tp = c[3 * k + k]
            fp = sum(c[3 * t + k] for t in range(3)) - tp
            fn = sum(c[3 * k + p] for p in range(3)) - tp
            f.append(2 * tp / (2 * tp + fp + fn) if 2 * tp + fp + fn else 0.0)
        return dict(val_acc=float(acc), val_f1=float(sum(f) / 3) if total else 0.0)
    if kind == KIND_DUAL:                                 # train_dual_head_final.py:196-232
        ptp, pfp, pfn, ptn, ttp, tfp, tfn, ttn = c[:8]
        pn, tn_ = ptp + pfp + pfn + ptn, ttp + tfp + tfn + ttn
        pres_f1 = 2 * ptp / (2 * ptp + pfp + pfn) if 2 * ptp + pfp + pfn else 0.0       # average="binary", pos_label=1
        tf = [2 * ttn / (2 * ttn + tfn + tfp) if 2 * ttn + tfn + tfp else 0.0,          # labels=[0, 1], macro
              2 * ttp / (2 * ttp + tfp + tfn) if 2 * ttp + tfp + tfn else 0.0]
        return dict(pres_acc_va=(ptp + ptn) / pn if pn else 0.0, truth_acc_va=(ttp + ttn) / tn_ if tn_ else 0.0,
                    pres_f1_va=float(pres_f1), truth_f1_va=float(sum(tf) / 2) if tn_ else 0.0)
    tp, fp, fn, tn = c[:4]
    n = tp + fp + fn + tn
    if not n:
        return dict(val_acc=0.0, val_f1=0.0)
    return dict(val_acc=(tp + tn) / n, val_f1=_f1_binary_macro(tp, fp, fn, tn))


def confusion_counts(kind: str, trainer: "ProbeTrainer", Z: torch.Tensor, Y: torch.Tensor, keep: torch.Tensor,
                     thresh: float = 0.5) -> List[int]:
    """Confusion counts of device logits Z [n, rows] against int8 labels Y [n, *] (columns `keep`), on the device."""
    dev = trainer.dev
    Yd = Y.to(dev, torch.int8).contiguous()
    kd = keep.to(dev, torch.int32).contiguous()
    counts = torch.zeros(9, dtype=torch.int64, device=dev)
    code = {KIND_OBJECT: 0, KIND_SPATIAL: 1, KIND_DUAL: 2, KIND_3CLASS: 3}[kind]
    trainer._lib_mod.check(trainer.lib.ovla_probe_confusion(
        C.c_void_p(Z.data_ptr()), C.c_longlong(Z.stride(0)), C.c_void_p(Yd.data_ptr()), C.c_longlong(Yd.stride(0)),
        C.c_void_p(kd.data_ptr()), Z.shape[0], trainer.K, trainer.Kpad, code, C.c_float(thresh),
        C.c_void_p(counts.data_ptr()), trainer._lib_mod.stream_ptr()))
    return counts.cpu().tolist()


def per_label_counts(trainer: "ProbeTrainer", Z: torch.Tensor, Y: torch.Tensor, keep: torch.Tensor,
                     thresh: float = 0.5) -> np.ndarray:
    """int64 [K, 4] (tp, fp, fn, tn) per kept label, counted on the device (mask y != -1, target y == 1)."""
    dev = trainer.dev
    Yd = Y.to(dev, torch.int8).contiguous()
    kd = keep.to(dev, torch.int32).contiguous()
    counts = torch.zeros(trainer.K, 4, dtype=torch.int64, device=dev)
    trainer._lib_mod.check(trainer.lib.ovla_probe_confusion_per_label(
        C.c_void_p(Z.data_ptr()), C.c_longlong(Z.stride(0)), C.c_void_p(Yd.data_ptr()), C.c_longlong(Yd.stride(0)),
        C.c_void_p(kd.data_ptr()), Z.shape[0], trainer.K, C.c_float(thresh), C.c_void_p(counts.data_ptr()),
        trainer._lib_mod.stream_ptr()))
    return counts.cpu().numpy()


def per_label_metrics(counts: np.ndarray, keep: Sequence[int]) -> List[Dict[str, float]]:
    """eval_probes_per_label.py:77-96 from per-label confusion counts: sklearn's precision / recall / F1
    (average="binary", zero_division=0), matthews_corrcoef and balanced_accuracy_score (NaN when the masked target holds
    a single class); labels whose mask is empty are skipped, as in the reference."""
    out = []
    for k, (tp, fp, fn, tn) in enumerate(np.asarray(counts, dtype=np.int64).tolist()):
        n = tp + fp + fn + tn
        if n == 0:
            continue
        prec = tp / (tp + fp) if tp + fp else 0.0
        rec = tp / (tp + fn) if tp + fn else 0.0
        f1 = 2 * tp / (2 * tp + fp + fn) if 2 * tp + fp + fn else 0.0
        both = (tp + fn) > 0 and (tn + fp) > 0                        # len(np.unique(targ)) > 1
        if both:
            den = float(tp + fp) * float(tp + fn) * float(tn + fp) * float(tn + fn)
            mcc = (float(tp) * tn - float(fp) * fn) / den ** 0.5 if den > 0 else 0.0
            bal = 0.5 * (tp / (tp + fn) + tn / (tn + fp))
        else:
            mcc = bal = float("nan")
        out.append(dict(label_idx=int(keep[k]), prec=float(prec), recall=float(rec), f1=float(f1), mcc=float(mcc),
                        bal_acc=float(bal)))
    return out


def evaluate(kind: str, trainer: ProbeTrainer, X: torch.Tensor, Y: torch.Tensor, keep: torch.Tensor, thresh: float = 0.5,
             on_device: bool = False):
    if on_device:
        # accuracy / F1 from integer confusion counts computed next to the logits: only 9 integers leave the GPU.
        # Average precision needs the sorted scores and stays on the sklearn path (on_device=False).
        Z = trainer.logits(X.to(trainer.dev, torch.float32).contiguous())
        if kind == KIND_SPATIAL:
            # train_spatial_probes.py:165-176 hands sklearn the 2-D indicator matrices: "macro" then averages the
            # positive-class F1 of every label column (not the two classes of the flattened problem)
            c = per_label_counts(trainer, Z, Y, keep, thresh).astype(np.int64)
            tp, fp, fn, tn = c[:, 0], c[:, 1], c[:, 2], c[:, 3]
            den = 2 * tp + fp + fn
            f1 = np.where(den > 0, 2.0 * tp / np.maximum(den, 1), 0.0)
            n = int(c.sum())
            return dict(val_acc=float((tp + tn).sum() / n) if n else 0.0, val_f1=float(f1.mean()) if len(f1) else 0.0)
        return metrics_from_counts(kind, confusion_counts(kind, trainer, Z, Y, keep, thresh))
    return _evaluate_host(kind, trainer, X, Y, keep, thresh)


def _evaluate_host(kind: str, trainer: ProbeTrainer, X: torch.Tensor, Y: torch.Tensor, keep: torch.Tensor, thresh: float = 0.5):
    """Validation metrics exactly as the reference computes them on the host with sklearn
    (train_object_probes.py:190-206; train_dual_head_final.py:196-232); logits come from the device GEMM."""
    from sklearn.metrics import average_precision_score, f1_score

    Z = trainer.logits(X.to(trainer.dev, torch.float32).contiguous()).cpu()
    y = Y[:, keep].long().cpu()
    K, Kp = trainer.K, trainer.Kpad
    if kind == KIND_3CLASS:                       # train_3class_direct.py:196-207: argmax over the 3 logits of each label
        tgt = (y + 1).view(-1)
        pred = Z[:, :3 * K].reshape(-1, 3).argmax(1)
        acc = float((pred == tgt).float().mean()) if tgt.numel() else 0.0
        f1 = f1_score(tgt.numpy(), pred.numpy(), labels=[0, 1, 2], average="macro", zero_division=0) if tgt.numel() else 0.0
        return dict(val_acc=acc, val_f1=f1)
    if kind == KIND_DUAL:
        pres_t, truth_t, mask = (y != -1).long(), (y == 1).long(), (y != -1)
        pres_p = (Z[:, :K].sigmoid() > 0.5).long()
        truth_p = (Z[:, Kp:Kp + K].sigmoid() > 0.5).long()
        pres_acc = float((pres_p == pres_t).float().mean()) if y.numel() else 0.0
        truth_acc = float((truth_p == truth_t)[mask].float().mean()) if mask.any() else 0.0
        pres_f1 = f1_score(pres_t.view(-1).numpy(), pres_p.view(-1).numpy(), average="binary", pos_label=1, zero_division=0)
        truth_f1 = f1_score(truth_t[mask].numpy(), truth_p[mask].numpy(), labels=[0, 1], average="macro",
                            zero_division=0) if mask.any() else 0.0
        return dict(pres_acc_va=pres_acc, truth_acc_va=truth_acc, pres_f1_va=pres_f1, truth_f1_va=truth_f1)
    probs = Z[:, :K].sigmoid()
    if kind == KIND_SPATIAL:                      # train_spatial_probes.py:165-176: no mask, 2-D multilabel sklearn inputs
        target = y.float()
        pred = (probs > thresh).float()
        if not y.numel():
            return dict(val_acc=0.0, val_f1=0.0, val_ap=0.0)
        return dict(val_acc=float((pred == target).float().mean()),
                    val_f1=f1_score(target.numpy(), pred.numpy(), average="macro", zero_division=0),
                    val_ap=average_precision_score(target.numpy(), probs.numpy(), average="macro"))
    mask, target = (y != -1), (y == 1).float()
    if not mask.any():
        return dict(val_acc=0.0, val_f1=0.0, val_ap=0.0)
    pred = (probs > thresh).long()
    acc = float((pred[mask] == target[mask]).float().mean())
    f1 = f1_score(target[mask].numpy(), pred[mask].numpy(), average="macro", zero_division=0)
    ap = average_precision_score(target[mask].numpy(), probs[mask].numpy(), average="macro")
    return dict(val_acc=acc, val_f1=f1, val_ap=ap)


# ----------------------------------------------------------------------------------------------- per-layer driver
def train_probes(kind: str, log_dir: str, layers: Sequence[int], epochs: int = 20, batch: int = 4096, device: int = 0,
                 out_dir: str = ".", exclude: Sequence[int] = (), seed: int = 0, group=None, verbose: bool = True,
                 concurrent: bool = True, hbm_budget_gb: float = 120.0):
    """Per-layer loop of the three reference scripts (train_object_probes.py:208-232, train_spatial_probes.py:179-204,
    train_dual_head_final.py:236-290).  Rank 0 writes the `.pth` files and the CSV.  `concurrent`: the probes of all
    requested layers train at the same time on shared shuffled batches (MultiLayerProbeTrainer) instead of one after the
    other; the per-layer results are the same training procedure, only the shuffle is common to the layers."""
    import pandas as pd

    cache = load_episodes(log_dir, exclude)
    if kind == KIND_SPATIAL:      # the spatial script decides `keep` over every episode file, excluded ones too
        every = cache if not exclude else load_episodes(log_dir, ())
        split = prepare_spatial(cache, [torch.cat([every[i]["symbolic_state_object_relations"],
                                                   every[i]["symbolic_state_action_subgoals"]], 1) for i in sorted(every)])
    else:
        split = {KIND_OBJECT: prepare_object, KIND_DUAL: prepare_dual, KIND_3CLASS: prepare_3class}[kind](cache)
    if kind in (KIND_DUAL, KIND_3CLASS):
        torch.manual_seed(seed)
    records = []
    rank0 = True

    def _save(L, tr_state, D_, rec):
        os.makedirs(out_dir, exist_ok=True)
        K_ = len(split.keep)
        if kind == KIND_3CLASS:
            torch.save({"model_type": "Direct3ClassProbe", "state_dict": tr_state, "layer": L,
                        "kept_indices": split.keep.tolist(), "input_dim": D_, "num_output_labels_probed": K_,
                        "num_classes_per_label": 3, "class_weights_used": [float(v) for v in split.pos_weight],
                        "class_mapping": {"NA(-1)": 0, "False(0)": 1, "True(1)": 2}},
                       os.path.join(out_dir, f"linear_probe_3class_direct_L{L:02d}.pth"))
        elif kind == KIND_DUAL:
            torch.save({"model_type": "DualHeadProbe", "state_dict": tr_state, "layer": L,
                        "kept_indices": split.keep.tolist(), "input_dim": D_, "num_output_labels": K_,
                        "presence_pos_weight_used": float(split.pos_weight)},
                       os.path.join(out_dir, f"linear_probe_dual_head_final_L{L:02d}.pth"))
        else:
            torch.save({"state_dict": tr_state, "layer": L, "kept": split.keep.tolist()},
                       os.path.join(out_dir, f"linear_probe_L{L:02d}.pth"))
        if verbose:
            print(f"L{L:02d}  " + "  ".join(f"{k}={v:.3f}" for k, v in rec.items() if k != "layer"))

    drop_last = kind in (KIND_DUAL, KIND_3CLASS)
    mats = {L: (layer_matrix(cache, split.train_ids, L), layer_matrix(cache, split.val_ids, L)) for L in layers}
    live = [L for L in layers if mats[L][0][0].shape[0] and mats[L][1][0].shape[0]]
    same = len({(tuple(mats[L][0][0].shape), tuple(mats[L][1][0].shape)) for L in live}) == 1
    if concurrent and kind != KIND_3CLASS and len(live) > 1 and same:
        # all layers at once (MultiLayerProbeTrainer): same labels, same shuffled batches, grouped kernels.  Layers are
        # processed in groups that keep features + the two shuffled copies within `hbm_budget_gb`.
        N, D_ = mats[live[0]][0][0].shape
        per_layer = 3.0 * N * D_ * 4
        gmax = max(1, int(hbm_budget_gb * 1e9 // per_layer))
        done = {}
        for i in range(0, len(live), gmax):
            grp = live[i:i + gmax]
            X = torch.stack([mats[L][0][0] for L in grp])
            Ytr = mats[grp[0]][0][1]
            tr = MultiLayerProbeTrainer(kind, len(grp), D_, len(split.keep), split.pos_weight, batch=batch, device=device,
                                        group=group)
            rank0 = tr.rank == 0
            tr.fit(X, Ytr, split.keep, epochs, seed=seed, drop_last=drop_last)
            for gi, L in enumerate(grp):
                sd = tr.state_dict(gi)
                ev = ProbeTrainer(kind, D_, len(split.keep), split.pos_weight, batch=batch, device=device, init_state=sd)
                done[L] = (dict(layer=L, **evaluate(kind, ev, mats[L][1][0], mats[L][1][1], split.keep)), sd)
            del tr, X
            torch.cuda.empty_cache()
        for L in layers:
            if L not in done:
                records.append(dict(layer=L, status="skipped_empty_data"))
                continue
            records.append(done[L][0])
            if rank0:
                _save(L, done[L][1], D_, done[L][0])
    else:
        for L in layers:
            (Xtr, Ytr), (Xva, Yva) = mats[L]
            if Xtr.shape[0] == 0 or Xva.shape[0] == 0:
                records.append(dict(layer=L, status="skipped_empty_data"))
                continue
            tr = ProbeTrainer(kind, Xtr.shape[1], len(split.keep), split.pos_weight, batch=batch, device=device, group=group)
            rank0 = tr.rank == 0
            tr.fit(Xtr, Ytr, split.keep, epochs, seed=seed + L, drop_last=drop_last)
            rec = dict(layer=L, **evaluate(kind, tr, Xva, Yva, split.keep))
            records.append(rec)
            if rank0:
                _save(L, tr.state_dict(), tr.D, rec)
    if rank0:
        name = {KIND_OBJECT: "probe_metrics_object.csv", KIND_SPATIAL: "probe_metrics_spatial.csv",
                KIND_DUAL: "probe_metrics_dual_head_final.csv", KIND_3CLASS: "probe_metrics_3class_direct.csv"}[kind]
        pd.DataFrame(records).to_csv(os.path.join(out_dir, name), index=False)
    return records


def main():
    cli = argparse.ArgumentParser(description="B200-native probe training (object / spatial / dual-head)")
    cli.add_argument("--kind", default=KIND_OBJECT, choices=[KIND_OBJECT, KIND_SPATIAL, KIND_DUAL, KIND_3CLASS])
    cli.add_argument("--log_dir", default="experiments/logs", help="folder containing episode_*.pt")
    cli.add_argument("--epochs", type=int, default=20)
    cli.add_argument("--batch", type=int, default=4096)
    cli.add_argument("--layers", default="all", help="comma-sep list, e.g. 32 or 0,8,16,32; 'all' = 0-32")
    cli.add_argument("--exclude_eps", default="")
    cli.add_argument("--out_dir", default=".")
    cli.add_argument("--sequential", action="store_true", help="one layer at a time, as the reference loops")
    args = cli.parse_args()
    layers = list(range(33)) if args.layers.strip().lower() == "all" else [int(x) for x in args.layers.split(",")]
    assert all(0 <= L <= 32 for L in layers), "layer index must be 0-32"
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist

        torch.cuda.set_device(local)
        os.environ.setdefault("NCCL_MAX_CTAS", "16")     # the all-reduce runs beside GEMMs that leave it 16 SMs
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    train_probes(args.kind, args.log_dir, layers, args.epochs, args.batch, device=local, out_dir=args.out_dir,
                 exclude=sorted(parse_exclusions(args.exclude_eps)), concurrent=not args.sequential)
    if world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()


if __name__ == "__main__":
    main()


# ----------------------------------------------------------------------------------------------- packed episode store
class PackedEpisodeStore:
    """Streaming alternative to one `episode_N.pt` pickle per rollout: a directory with
    `features.npy` = float32 [n_layers, N, D] (memory-mapped, appended step by step), `relations.npy` / `subgoals.npy` =
    int8 [N, *] and `episodes.json` = [(first_row, n_rows)] per episode.  The probe trainer can map one layer slab
    without unpickling everything; `export_reference_episodes` converts back to the reference's dict layout
    (`run_libero_eval_object.py:357-366`) so the unmodified reference scripts can still read it."""

    def __init__(self, root: str, n_layers: int = 33, dim: int = 4096, n_rel: int = 461, n_act: int = 20,
                 capacity: int = 65536, mode: str = "w+"):
        import json

        self.root, self.n_layers, self.dim = root, n_layers, dim
        os.makedirs(root, exist_ok=True)
        meta_path = os.path.join(root, "episodes.json")
        if mode == "r":
            meta = json.load(open(meta_path))
            self.n_layers, self.dim, n_rel, n_act, capacity = meta["n_layers"], meta["dim"], meta["n_rel"], meta["n_act"], meta["capacity"]
            self.episodes, self.n_rows = [tuple(e) for e in meta["episodes"]], meta["n_rows"]
        else:
            self.episodes, self.n_rows = [], 0
        self.n_rel, self.n_act, self.capacity = n_rel, n_act, capacity
        mm = "r" if mode == "r" else mode
        self.features = np.lib.format.open_memmap(os.path.join(root, "features.npy"), mode=mm, dtype=np.float32,
                                                  shape=(self.n_layers, capacity, self.dim)) if mode != "r" else \
            np.load(os.path.join(root, "features.npy"), mmap_mode="r")
        self.relations = np.lib.format.open_memmap(os.path.join(root, "relations.npy"), mode=mm, dtype=np.int8,
                                                   shape=(capacity, n_rel)) if mode != "r" else \
            np.load(os.path.join(root, "relations.npy"), mmap_mode="r")
        self.subgoals = np.lib.format.open_memmap(os.path.join(root, "subgoals.npy"), mode=mm, dtype=np.int8,
                                                  shape=(capacity, n_act)) if mode != "r" else \
            np.load(os.path.join(root, "subgoals.npy"), mmap_mode="r")
        self._ep_start = self.n_rows

    def append_batch(self, pooled: np.ndarray, rel: np.ndarray, act: np.ndarray) -> None:
        """pooled fp32 [n_layers, B, D] straight from `predict_action_and_capture`; rel / act int8 [B, *]."""
        B = pooled.shape[1]
        if self.n_rows + B > self.capacity:
            raise ValueError(f"store capacity {self.capacity} exceeded")
        self.features[:, self.n_rows:self.n_rows + B] = pooled
        self.relations[self.n_rows:self.n_rows + B] = rel
        self.subgoals[self.n_rows:self.n_rows + B] = act
        self.n_rows += B

    def end_episode(self) -> None:
        self.episodes.append((self._ep_start, self.n_rows - self._ep_start))
        self._ep_start = self.n_rows

    def flush(self) -> None:
        import json

        for a in (self.features, self.relations, self.subgoals):
            if hasattr(a, "flush"):
                a.flush()
        json.dump({"n_layers": self.n_layers, "dim": self.dim, "n_rel": self.n_rel, "n_act": self.n_act,
                   "capacity": self.capacity, "n_rows": self.n_rows, "episodes": self.episodes},
                  open(os.path.join(self.root, "episodes.json"), "w"))

    def layer(self, L: int) -> np.ndarray:
        return self.features[L, : self.n_rows]

    def as_cache(self) -> Dict[int, dict]:
        """The `cache` dict of the trainers (train_object_probes.py:59-69) backed by the memory maps (zero copy)."""
        out = {}
        for i, (s, n) in enumerate(self.episodes):
            out[i] = {"visual_semantic_encoding": {L: torch.from_numpy(np.asarray(self.features[L, s:s + n])) for L in range(self.n_layers)},
                      "symbolic_state_object_relations": torch.from_numpy(np.asarray(self.relations[s:s + n])),
                      "symbolic_state_action_subgoals": torch.from_numpy(np.asarray(self.subgoals[s:s + n]))}
        return out

    def export_reference_episodes(self, log_dir: str) -> None:
        os.makedirs(log_dir, exist_ok=True)
        for i, d in self.as_cache().items():
            torch.save({"visual_semantic_encoding": {L: v.clone() for L, v in d["visual_semantic_encoding"].items()},
                        "symbolic_state_object_relations": d["symbolic_state_object_relations"].clone(),
                        "symbolic_state_action_subgoals": d["symbolic_state_action_subgoals"].clone()},
                       os.path.join(log_dir, f"episode_{i + 1}.pt"))
