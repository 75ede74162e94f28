"""Probe path pinned to the reference's own scripts (tests/golden/probe_golden.pt).

The fixture was produced by tests/golden/make_probe_golden.py: the reference's UNMODIFIED
`experiment_utils/train_{object,spatial}_probes.py`, `train_dual_head_final.py`, `train_3class_direct.py` were run on
episode files written by `probes.EpisodeWriter`, and the CPU restatement (oracle/probe_oracle.py) reproduced their saved
state dicts bit for bit; the fixture holds the episodes, `kept`, pos_weight, initial weights, per-step batch orders /
losses and the final weights of that run.

CPU tests: the restatement still reproduces the fixture; the PRODUCT's host-side dataset preparation equals the
reference's; episode files load in the reference trainers (run directly when /root/reference exists).
GPU tests: the device trainer, started from the reference's initial weights and fed the reference's batch order, follows
the reference's loss trajectory and lands on its final weights (TF32 GEMMs; tolerances stated below).
"""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import probe_oracle as PO
from openvla_probe_b200 import probes

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "probe_golden.pt")
REF = "/root/reference"
KINDS = ["object", "spatial", "dual", "3class"]


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLD, map_location="cpu", weights_only=False)


def _cache(gold, kind, with_excluded=False):
    """The trainers' `cache` dict: episode files in glob-sorted order (episode_1, episode_10, ...), exclusions dropped."""
    g = gold["kinds"][kind]
    eps = sorted(gold["episodes"][g["suite"]], key=lambda e: f"episode_{e['n']}.pt")
    if not with_excluded:
        eps = [e for e in eps if e["n"] not in set(g["exclude"])]
    return {i: {"visual_semantic_encoding": dict(e["features"]), "symbolic_state_object_relations": e["rel"],
                "symbolic_state_action_subgoals": e["act"]} for i, e in enumerate(eps)}


def _all_mats(gold, kind):
    c = _cache(gold, kind, with_excluded=True)
    return [torch.cat([c[i]["symbolic_state_object_relations"], c[i]["symbolic_state_action_subgoals"]], 1) for i in sorted(c)]


@pytest.mark.parametrize("kind", KINDS)
def test_restatement_reproduces_the_reference_run(gold, kind):
    g = gold["kinds"][kind]
    cache = _cache(gold, kind)
    if kind in ("object", "spatial"):
        torch.manual_seed(gold["seed"])
    res = PO.train_layers(kind, cache, gold["layers"] if kind in ("object", "spatial") else range(33), gold["epochs"],
                          gold["batch"], all_label_mats=_all_mats(gold, kind) if kind == "spatial" else None, seed=0)
    assert res["keep"].tolist() == g["keep"].tolist()
    assert torch.equal(torch.as_tensor(res["pos_weight"]), torch.as_tensor(g["pos_weight"]))
    assert res["train_ids"] == g["train_ids"] and res["val_ids"] == g["val_ids"]
    for L in gold["layers"]:
        a, b = res["layers"][L], g["layers"][L]
        for k in b["init"]:
            assert torch.equal(a["init"][k], b["init"][k])                       # same RNG consumption as the script
        assert all(torch.equal(x, y) for ea, eb in zip(a["orders"], b["orders"]) for x, y in zip(ea, eb))
        np.testing.assert_allclose(a["losses"], b["losses"], rtol=1e-5)        # thread count may change the reduction order
        for k in b["final"]:
            torch.testing.assert_close(a["final"][k], b["final"][k], rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("kind", KINDS)
def test_product_dataset_preparation_equals_the_reference(gold, kind):
    """probes.prepare_* (split, keep filter, pos_weight / class weights) and probes.layer_matrix against the values the
    reference scripts computed (train_object_probes.py:72-102, train_spatial_probes.py:90-131,
    train_dual_head_final.py:76-127, train_3class_direct.py:75-131)."""
    g = gold["kinds"][kind]
    cache = _cache(gold, kind)
    if kind == "spatial":
        sp = probes.prepare_spatial(cache, _all_mats(gold, kind))
    else:
        sp = {"object": probes.prepare_object, "dual": probes.prepare_dual, "3class": probes.prepare_3class}[kind](cache)
    assert sp.train_ids == g["train_ids"] and sp.val_ids == g["val_ids"]
    assert sp.keep.tolist() == g["keep"].tolist()
    assert torch.equal(sp.pos_weight.float(), torch.as_tensor(g["pos_weight"]).float())
    for L in gold["layers"]:
        X, Y = probes.layer_matrix(cache, sp.train_ids, L)
        Xo, Yo = PO.layer_samples(cache, g["train_ids"], L, dual_rule=kind in ("dual", "3class"))
        assert torch.equal(X, Xo) and torch.equal(Y, Yo.to(torch.int8))
        assert X.shape[0] == g["layers"][L]["n_train"]
    assert probes.layer_matrix(cache, sp.train_ids, 5)[0].numel() == 0            # layer absent from the files


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree exists only in the build container")
def test_episode_layout_roundtrip_in_reference_trainers(tmp_path):
    """files written by the new path load in the UNMODIFIED reference trainers
    (run_libero_eval_object.py:357-366 layout -> train_object_probes.py:61-69,129-145; train_dual_head_final.py:60-75),
    directly and through PackedEpisodeStore.export_reference_episodes."""
    import shutil

    rng = np.random.default_rng(0)
    D, layers = 32, (0, 32)
    lib = tmp_path / "experiments" / "robot" / "libero"
    lib.mkdir(parents=True)
    for f in os.listdir(os.path.join(REF, "experiments/robot/libero")):
        if f.endswith("_keys.txt"):
            shutil.copy(os.path.join(REF, "experiments/robot/libero", f), lib)
    logs = tmp_path / "experiments" / "logs"
    store = probes.PackedEpisodeStore(str(tmp_path / "packed"), n_layers=33, dim=D, capacity=256)
    for n in range(1, 7):
        T = int(rng.integers(8, 14))
        pooled = rng.standard_normal((33, T, D)).astype(np.float32)
        rel = rng.integers(-1, 2, (T, 461)).astype(np.int8)
        act = rng.integers(-1, 2, (T, 20)).astype(np.int8)
        if n <= 3:
            w = probes.EpisodeWriter(layers=layers)
            w.append_batch(pooled, rel, act)
            w.save(str(logs / f"episode_{n}.pt"))
        else:
            store.append_batch(pooled, rel, act)
            store.end_episode()
    store.flush()
    exported = tmp_path / "exported"
    probes.PackedEpisodeStore(str(tmp_path / "packed"), mode="r").export_reference_episodes(str(exported))
    for i, f in enumerate(sorted(os.listdir(exported)), start=4):
        shutil.move(str(exported / f), str(logs / f"episode_{i}.pt"))

    def run(script, argv):
        code = (f"import sys, runpy\nsys.argv = {[script] + argv!r}\n"
                f"runpy.run_path({os.path.join(REF, 'experiment_utils', script)!r}, run_name='__main__')\n")
        r = subprocess.run([sys.executable, "-c", code], cwd=tmp_path, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]

    run("train_object_probes.py", ["--epochs", "1", "--batch", "32", "--device", "cpu", "--layers", "0,32"])
    run("train_dual_head_final.py", ["--epochs", "1", "--batch", "16", "--device", "cpu", "--num_workers", "0"])
    ck = torch.load(tmp_path / "linear_probe_L32.pth", map_location="cpu", weights_only=False)
    assert set(ck) == {"state_dict", "layer", "kept"} and ck["state_dict"]["weight"].shape == (len(ck["kept"]), D)
    cache = probes.load_episodes(str(logs))
    assert ck["kept"] == probes.prepare_object(cache).keep.tolist()
    dk = torch.load(tmp_path / "linear_probe_dual_head_final_L00.pth", map_location="cpu", weights_only=False)
    sp = probes.prepare_dual(cache)
    assert dk["kept_indices"] == sp.keep.tolist() and dk["presence_pos_weight_used"] == float(sp.pos_weight.float())


# ----------------------------------------------------------------------------------------------- GPU
def _gpu_run(gold, kind, L):
    g = gold["kinds"][kind]
    cache = _cache(gold, kind)
    rec = g["layers"][L]
    X, Y = probes.layer_matrix(cache, g["train_ids"], L)
    tr = probes.ProbeTrainer(kind, X.shape[1], len(g["keep"]), g["pos_weight"], batch=gold["batch"], init_state=rec["init"])
    Xd, Yd = X.cuda(), Y.cuda()
    losses = []
    for ep in rec["orders"]:
        perm = torch.cat(ep)
        tr.load_epoch(Xd, Yd, g["keep"], perm, drop_last=kind in ("dual", "3class"))
        assert len(tr.steps) == len(ep)
        for s in range(len(ep)):
            tr.train_step(s)
            losses.append(tr.step_loss())
    return tr, losses, rec


@pytest.mark.gpu
@pytest.mark.parametrize("kind", KINDS)
def test_device_trainer_follows_the_reference_trajectory(gold, kind):
    """Same initial weights, same batch order as the reference script's run -> same losses and final weights.
    Stated tolerance: the two GEMMs of a step multiply in TF32 (the reference: fp32), so per-step loss within 2e-3
    relative, final weights within rel-L2 2e-3, and the accumulated update (final - init) within rel-L2 5e-2 (AdamW
    normalises each element's step, so elements whose gradient is near zero may move differently)."""
    for L in gold["layers"]:
        tr, losses, rec = _gpu_run(gold, kind, L)
        np.testing.assert_allclose(losses, rec["losses"], rtol=2e-3)
        sd = tr.state_dict()
        for k, ref in rec["final"].items():
            got, init = sd[k].float(), rec["init"][k].float()
            assert float((got - ref).norm() / ref.norm()) < 2e-3, (kind, L, k)
            assert float(((got - init) - (ref - init)).norm() / (ref - init).norm()) < 5e-2, (kind, L, k)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", KINDS)
def test_device_validation_metrics_equal_the_reference_csv(gold, kind):
    """Validation metrics of the reference's FINAL weights, evaluated by the device path (TF32 logits + on-device
    confusion counters), against the values in the reference's CSV (sklearn on the host): a logit within TF32 error of
    the threshold may flip one prediction, so accuracies agree to 2 counts and F1 to 5e-3."""
    g = gold["kinds"][kind]
    cache = _cache(gold, kind)
    for L in gold["layers"]:
        rec = g["layers"][L]
        Xva, Yva = probes.layer_matrix(cache, g["val_ids"], L)
        tr = probes.ProbeTrainer(kind, Xva.shape[1], len(g["keep"]), g["pos_weight"], batch=gold["batch"], init_state=rec["final"])
        for on_device in (False, True):
            m = probes.evaluate(kind, tr, Xva, Yva, g["keep"], on_device=on_device)
            n = Yva[:, g["keep"]].numel()
            for k, v in rec["metrics"].items():
                if k == "val_ap" and on_device:
                    continue
                tol = 2.0 / n + 1e-9 if "acc" in k else 5e-3
                assert abs(m[k] - v) <= tol, (kind, L, on_device, k, m[k], v)
