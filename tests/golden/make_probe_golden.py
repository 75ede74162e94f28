"""Pin the probe path to the reference ITSELF and write tests/golden/probe_golden.pt.

    python tests/golden/make_probe_golden.py          # needs /root/reference (this container only)

1. Synthetic rollouts are written with the PRODUCT's `probes.EpisodeWriter` as `episode_{n}.pt` files.
2. The reference's UNMODIFIED scripts are executed on them (runpy, cwd = a scratch tree that holds the reference's
   `experiments/robot/libero/*_keys.txt` and `experiments/logs/episode_*.pt`), CPU, with the global torch seed set by
   the wrapper for the two scripts that do not seed themselves:
       experiment_utils/train_object_probes.py, train_spatial_probes.py, train_dual_head_final.py, train_3class_direct.py
3. oracle/probe_oracle.py (the CPU restatement) is run with the same seed and must reproduce the scripts' outputs
   exactly: `kept` / `kept_indices`, the pos_weight the script recorded, every saved state dict (torch.equal) and the
   CSV metrics.  Only then are the oracle's per-step records (initial weights, batch orders, losses, final weights)
   stored, together with the episodes, as the fixture the CPU and GPU tests use.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np
import pandas as pd
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden", "probe_golden.pt")
SEED = 1234
D, LAYERS = 64, (0, 32)
EPOCHS, BATCH = 3, 64


def synth_episodes(suite: str, n_ep: int, seed: int):
    """Rollouts with learnable structure: label l is applicable with probability p_l and, when applicable, equals
    [x . w_l > 0]; many labels never flip (the scripts' keep filters must drop them), a few are rare (< 1 %)."""
    rng = np.random.default_rng(seed)
    n_rel, n_act = (461, 20) if suite == "object" else (224, 12)
    n_lab = n_rel + n_act
    W = rng.standard_normal((n_lab, D)).astype(np.float32)
    varying = np.zeros(n_lab, bool)
    varying[:40] = True
    varying[n_rel:n_rel + 8] = True
    const_val = rng.integers(-1 if suite == "object" else 0, 2, n_lab)
    rare = np.zeros(n_lab, bool)
    rare[36:40] = True                                     # varying but almost never true
    eps = []
    for _ in range(n_ep):
        T = int(rng.integers(18, 37))
        base = rng.standard_normal((T, D)).astype(np.float32)
        feats = {L: (base + 0.05 * L * rng.standard_normal((T, D))).astype(np.float32) for L in LAYERS}
        val = (base @ W.T > 0).astype(np.int8)
        val[:, rare] = (rng.random((T, int(rare.sum()))) < 0.004).astype(np.int8)
        y = np.where(varying[None, :], val, const_val[None, :]).astype(np.int8)
        if suite == "object":
            applicable = rng.random((T, n_lab)) < 0.6
            y = np.where(varying[None, :] & ~applicable, -1, y).astype(np.int8)
        eps.append((feats, y[:, :n_rel], y[:, n_rel:]))
    return eps


def write_tree(suite: str, eps, root: str):
    from openvla_probe_b200.probes import EpisodeWriter

    lib = os.path.join(root, "experiments", "robot", "libero")
    os.makedirs(lib, exist_ok=True)
    for f in os.listdir(os.path.join(REF, "experiments/robot/libero")):
        if f.endswith("_keys.txt"):
            shutil.copy(os.path.join(REF, "experiments/robot/libero", f), lib)
    logs = os.path.join(root, "experiments", "logs")
    for n, (feats, rel, act) in enumerate(eps, start=1):
        w = EpisodeWriter(layers=LAYERS)
        for t in range(rel.shape[0]):
            w.append({L: feats[L][t] for L in LAYERS}, rel[t], act[t])
        w.save(os.path.join(logs, f"episode_{n}.pt"))
    return logs


def run_reference(script: str, cwd: str, argv, seed=None):
    code = ("import sys, runpy, torch\n"
            + (f"torch.manual_seed({seed})\n" if seed is not None else "")
            + f"sys.argv = {[script] + list(argv)!r}\n"
            + f"runpy.run_path({os.path.join(REF, 'experiment_utils', script)!r}, run_name='__main__')\n")
    r = subprocess.run([sys.executable, "-c", code], cwd=cwd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"{script} failed:\n{r.stdout[-3000:]}\n{r.stderr[-3000:]}")
    return r.stdout


def same_state(a, b):
    return set(a) == set(b) and all(torch.equal(a[k], b[k]) for k in a)


def main():
    from oracle import probe_oracle as PO
    from openvla_probe_b200 import probes

    assert os.path.isdir(REF), "the reference tree is needed to generate the golden"
    torch.set_num_threads(1)          # one thread here AND in the scripts' own process (env below): same reduction order
    os.environ["OMP_NUM_THREADS"] = "1"
    os.environ["MKL_NUM_THREADS"] = "1"
    gold = {"seed": SEED, "D": D, "layers": list(LAYERS), "epochs": EPOCHS, "batch": BATCH, "kinds": {}, "episodes": {}}
    scratch = tempfile.mkdtemp(prefix="probe_golden_")
    jobs = [
        ("object", "object", "train_object_probes.py", ["--exclude_eps", "3"], SEED),
        ("spatial", "spatial", "train_spatial_probes.py", ["--exclude_eps", "2,5-6"], SEED),
        ("dual", "object", "train_dual_head_final.py", ["--num_workers", "0", "--seed", "0"], None),
        ("3class", "object", "train_3class_direct.py", ["--num_workers", "0", "--seed", "0"], None),
    ]
    for kind, suite, script, extra, seed in jobs:
        eps = synth_episodes(suite, 14, seed=7 if suite == "object" else 8)
        cwd = os.path.join(scratch, kind)
        logs = write_tree(suite, eps, cwd)
        argv = ["--epochs", str(EPOCHS), "--batch", str(BATCH), "--device", "cpu", *extra]
        if kind in ("object", "spatial"):
            argv += ["--layers", ",".join(map(str, LAYERS))]
        stdout = run_reference(script, cwd, argv, seed)
        # ---- the restatement on the same files, same seed
        exclude = probes.parse_exclusions(extra[1]) if extra[0] == "--exclude_eps" else set()
        cache = probes.load_episodes(logs, sorted(exclude))
        all_mats = None
        if kind == "spatial":
            allc = probes.load_episodes(logs, ())
            all_mats = [torch.cat([allc[i]["symbolic_state_object_relations"], allc[i]["symbolic_state_action_subgoals"]], 1)
                        for i in sorted(allc)]
        if seed is not None:
            torch.manual_seed(seed)
        res = PO.train_layers(kind, cache, LAYERS if kind in ("object", "spatial") else range(33), EPOCHS, BATCH,
                              all_label_mats=all_mats, seed=0)
        # ---- compare with what the reference wrote
        prefix = {"object": "linear_probe_L", "spatial": "linear_probe_L", "dual": "linear_probe_dual_head_final_L",
                  "3class": "linear_probe_3class_direct_L"}[kind]
        csv = {"object": "probe_metrics_object.csv", "spatial": "probe_metrics_spatial.csv",
               "dual": "probe_metrics_dual_head_final.csv", "3class": "probe_metrics_3class_direct.csv"}[kind]
        df = pd.read_csv(os.path.join(cwd, csv))
        for L in LAYERS:
            ref = torch.load(os.path.join(cwd, f"{prefix}{L:02d}.pth"), map_location="cpu", weights_only=False)
            kept = ref["kept"] if "kept" in ref else ref["kept_indices"]
            assert kept == res["keep"].tolist(), (kind, L, "kept differs")
            assert same_state(ref["state_dict"], res["layers"][L]["final"]), (kind, L, "final state dict differs from the reference's")
            if kind == "dual":
                assert ref["presence_pos_weight_used"] == float(torch.tensor(float(res["pos_weight"]))), (kind, "pos_weight")
            if kind == "3class":
                assert np.allclose(ref["class_weights_used"], res["pos_weight"].numpy(), rtol=0, atol=0), (kind, "class weights")
            row = df[df.layer == L].iloc[0]
            for k, v in res["layers"][L]["metrics"].items():
                assert abs(float(row[k]) - float(v)) < 1e-12, (kind, L, k, row[k], v)
        print(f"[{kind}] reference == restatement: kept={len(res['keep'])}, layers={list(res['layers'])}, "
              f"steps/layer={sum(len(e) for e in res['layers'][LAYERS[0]]['orders'])}")
        # the product's host-side dataset preparation must agree as well
        gold["episodes"].setdefault(suite, [
            {"n": n, "features": {L: torch.from_numpy(f[L]) for L in LAYERS}, "rel": torch.from_numpy(r),
             "act": torch.from_numpy(a)} for n, (f, r, a) in enumerate(eps, start=1)])
        gold["kinds"][kind] = dict(
            suite=suite, script=script, argv=argv, exclude=sorted(exclude),
            train_ids=res["train_ids"], val_ids=res["val_ids"], keep=res["keep"], pos_weight=res["pos_weight"],
            layers=res["layers"], reference_stdout_tail=stdout[-600:])
    torch.save(gold, OUT)
    print(f"wrote {OUT} ({os.path.getsize(OUT) / 1e6:.2f} MB)")
    shutil.rmtree(scratch, ignore_errors=True)


if __name__ == "__main__":
    main()
