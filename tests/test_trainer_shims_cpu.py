"""Host shims, trainer part: the torchpack stand-in (u2mkd_b200/shims/torchpack_shim.py), the synthetic
nuScenes dataset adapter and — when the reference checkout is present — the reference's OWN NuScenesTrainer / MeanIoU classes
running unchanged over them, driven as train_spformer.py drives them."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def test_trainer_hook_order_summary_and_savers(tmp_path):
    from u2mkd_b200.shims import torchpack_shim as tp
    tp.set_run_dir(str(tmp_path))
    log = []

    class T(tp.Trainer):
        def _before_epoch(self): log.append("T.before_epoch")
        def _after_epoch(self): log.append("T.after_epoch")
        def _trigger_epoch(self): log.append("T.trigger_epoch")

        def _run_step(self, feed_dict):
            self.summary.add_scalar("loss", 10.0 - self.global_step)
            self.summary.add_scalar("acc", [1.0, 3.0, 2.0][self.epoch_num - 1])
            return {"y": feed_dict * 2}

        def _state_dict(self): return {"w": torch.ones(2) * self.global_step}
        def _load_state_dict(self, sd): log.append(("loaded", float(sd["w"][0])))

    seen = []
    cb = tp.LambdaCallback(before_epoch=lambda c: log.append("C.before_epoch"), after_epoch=lambda c: log.append("C.after_epoch"),
                           trigger_epoch=lambda c: log.append("C.trigger_epoch"),
                           after_step=lambda c, out: seen.append(out["y"]))
    inf_seen = []
    runner = tp.InferenceRunner([5, 6], callbacks=[tp.LambdaCallback(after_step=lambda c, out: inf_seen.append(out["y"]))])
    t = T()
    t.train_with_defaults([1, 2, 3], num_epochs=3, callbacks=[cb, runner, tp.MaxSaver("acc"), tp.MinSaver("loss"), tp.Saver(max_to_keep=2)])
    assert log[:6] == ["T.before_epoch", "C.before_epoch", "C.after_epoch", "T.after_epoch", "C.trigger_epoch", "T.trigger_epoch"]
    assert seen == [2, 4, 6] * 3 and inf_seen == [10, 12] * 3
    assert t.global_step == 9 and t.epoch_num == 3 and t.steps_per_epoch == 3
    # (the inference runs also call run_step, which logs scalars at the same global_step: 5 entries per epoch)
    assert [s for s, _ in t.summary["loss"]][:3] == [1, 2, 3] and "acc" in t.summary and "nope" not in t.summary
    ck = sorted(os.listdir(tmp_path / "checkpoints"))
    assert ck == ["max-acc.pt", "min-loss.pt", "step-6.pt", "step-9.pt"], ck          # max_to_keep = 2
    best = tp.io.load(str(tmp_path / "checkpoints" / "max-acc.pt"))
    assert best["epoch_num"] == 2 and best["global_step"] == 6                        # acc peaked in epoch 2
    assert t.summary["acc/max"][-1][1] == 3.0
    t.load_state_dict(tp.io.load(str(tmp_path / "checkpoints" / "step-9.pt")))
    assert log[-1] == ("loaded", 9.0) and t.global_step == 9
    rows = [json.loads(l) for l in open(tmp_path / "summary" / "scalars.jsonl")]
    assert len(rows) == 3 and rows[-1]["epoch_num"] == 3 and "loss" in rows[-1]


def test_small_stand_ins():
    from u2mkd_b200.shims import torchpack_shim as tp
    pt = tp.PrettyTable()
    pt.field_names = ["Item", "a", "Mean"]
    pt.add_row(["IoU", 12.5, 50])
    s = str(pt)
    assert s.count("\n") == 4 and "| IoU  | 12.5 |  50  |" in s
    cm = tp.ConfusionMatrix(3, ignore_idx=0)
    cm.update(np.array([1, 1, 2, 2, 0]), np.array([1, 2, 2, 2, 1]))
    iou = cm.get_per_class_iou()
    assert np.isnan(iou[0]) and iou[1] == pytest.approx(0.5) and iou[2] == pytest.approx(2 / 3)
    assert cm.get_mean_iou() == pytest.approx((0.5 + 2 / 3) / 2)
    d = tp.distributed
    assert d.size() == 1 and d.rank() == 0 and d.is_master() and d.allreduce(3.5, reduction="sum") == 3.5 and d.allgather("x") == ["x"]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _dist_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from u2mkd_b200.shims import torchpack_shim as tp
    d = tp.distributed
    d.init()                                         # env:// rendezvous, gloo without CUDA
    out = {"size": d.size(), "rank": d.rank(), "local_rank": d.local_rank(), "master": d.is_master(),
           "sum": d.allreduce(float(rank + 1), reduction="sum"), "max": d.allreduce(rank, reduction="max"),
           "gather": d.allgather([rank, rank * 2]), "bcast": d.broadcast("from%d" % rank, src=0),
           "saver_is_inert": type(tp.Saver(save_dir="/tmp/u2_never_written")).__name__}
    d.barrier()
    q.put(out)
    import torch.distributed as td
    td.destroy_process_group()


@pytest.mark.timeout(300)
def test_torchpack_distributed_two_ranks_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dist_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = sorted([q.get(timeout=240) for _ in procs], key=lambda o: o["rank"])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert [o["size"] for o in outs] == [2, 2] and [o["master"] for o in outs] == [True, False]
    assert all(o["sum"] == 3.0 and o["max"] == 1 and o["gather"] == [[0, 0], [1, 2]] and o["bcast"] == "from0" for o in outs)
    assert outs[0]["saver_is_inert"] == "Saver" and outs[1]["saver_is_inert"] == "LambdaCallback"   # master_only


def test_synthetic_nuscenes_feed_dict_contract(oracle):
    """Keys, types and invariants of semantic_nusc.py:338-351 (inverse map reconstructs every point's voxel, first-point
    features, keyframe masks, collate appends the batch index)."""
    saved = {k: sys.modules.get(k) for k in list(sys.modules) if k == "torchsparse" or k.startswith("torchsparse.")}
    oracle.install_as_torchsparse()
    try:
        from u2mkd_b200.shims.synthetic_nusc import SyntheticNuScenes
        ds = SyntheticNuScenes(voxel_size=0.2, num_train=3, num_val=2, multisweeps=2, max_points=4000)
        assert set(ds) == {"train", "val"} and len(ds["train"]) == 3
        a = ds["val"][1]
        assert set(a) == {"lidar", "targets", "targets_mapped", "inverse_map", "lidar_token", "num_vox", "keyframe_mask", "keyframe_mask_full"}
        lidar, inv, full = a["lidar"], a["inverse_map"], a["targets_mapped"]
        assert lidar.F.dtype == np.float32 and lidar.F.shape == (a["num_vox"], 4) and lidar.C.dtype == np.int32 and lidar.C.min() == 0
        assert np.array_equal(lidar.C[inv.F], inv.C) and np.array_equal(full.C, inv.C)            # voxel of every point
        assert len(np.unique(lidar.C, axis=0)) == a["num_vox"]
        km, kmf = a["keyframe_mask"].F, a["keyframe_mask_full"].F
        assert kmf.dtype == bool and 0 < kmf.sum() < kmf.size and bool((full.F[~kmf] == 0).all()) and bool((full.F[kmf] > 0).all())
        assert km.shape == (a["num_vox"],)
        b = ds["val"][1]
        assert np.array_equal(a["lidar"].C, b["lidar"].C)                                          # val split: deterministic
        batch = ds["train"].collate_fn([ds["train"][0], ds["train"][1]])
        assert batch["lidar"].C.shape[1] == 4 and set(batch["lidar"].C[:, 3].tolist()) == {0, 1}
        assert batch["num_vox"] == [ds["train"][0]["num_vox"], ds["train"][1]["num_vox"]] or len(batch["num_vox"]) == 2
        assert batch["lidar"].F.shape[0] == batch["targets"].F.shape[0] == batch["keyframe_mask"].F.shape[0]
    finally:
        for k in [k for k in sys.modules if k == "torchsparse" or k.startswith("torchsparse.")]:
            del sys.modules[k]
        sys.modules.update({k: v for k, v in saved.items() if v is not None})


@pytest.mark.timeout(900)
@pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "core", "spformer_trainer.py")), reason="reference tree not present (GPU box)")
def test_unmodified_reference_trainer_runs_on_the_shims():
    """core/spformer_trainer.py NuScenesTrainer + core/callbacks.py MeanIoU, unchanged, through train_with_defaults with
    InferenceRunner / MaxSaver / Saver as in train_spformer.py:98-115 — 2 epochs on the synthetic dataset (CPU: the oracle's
    torchsparse namespace, `.cuda()` = identity; tests/trainer_shim_run.py in a fresh interpreter)."""
    r = subprocess.run([sys.executable, os.path.join(HERE, "trainer_shim_run.py")], capture_output=True, text=True, timeout=850)
    assert r.returncode == 0, r.stderr[-3000:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert "torchpack" in out["shims"] and "visualize_utils" in out["shims"]
    assert out["global_step"] == 4 and out["epoch_num"] == 2 and len(out["losses"]) == 4
    assert all(np.isfinite(out["losses"])) and out["losses"][-1] < out["losses"][0] and out["moved"]
    assert len(out["miou"]) == 2 and all(0.0 <= v <= 100.0 for v in out["miou"])
    # the reference trainer's _after_epoch switches to eval AFTER the callbacks' after_epoch and before trigger_epoch,
    # where InferenceRunner validates (hook order of the stand-in = torchpack's)
    assert out["order"] == [["before_epoch", True], ["after_epoch", True], ["trigger_epoch", False]] * 2
    assert out["checkpoints"] == ["max-iou-val-vox.pt", "step-4.pt"]                               # Saver(max_to_keep=1)
    assert {"model", "optimizer", "scheduler", "scaler", "epoch_num", "global_step"} <= set(out["state_keys"])
    assert out["state_steps"] == [2, 4]


def test_config_recursive_load_and_command_line_overrides(tmp_path):
    """torchpack.utils.config.Config as train_spformer.py:32-33 uses it: load(path, recursive=True) merges every default.yaml
    on the way down to the file, update([...]) takes `--a.b value` / `--a.b=value` strings."""
    from u2mkd_b200.shims import Config
    (tmp_path / "cfg" / "nusc" / "train").mkdir(parents=True)
    (tmp_path / "cfg" / "default.yaml").write_text("workers: 4\namp: false\n")
    (tmp_path / "cfg" / "nusc" / "default.yaml").write_text("data:\n  classes: 17\n  root: /data\nepochs: 25\n")
    (tmp_path / "cfg" / "nusc" / "train" / "x.yaml").write_text("data:\n  root: /other\nmodel:\n  name: spvcnn\n  cr: 1.0\nepochs: 3\n")
    c = Config()
    c.load(str(tmp_path / "cfg" / "nusc" / "train" / "x.yaml"), recursive=True)
    assert c.workers == 4 and c.data.classes == 17 and c.data.root == "/other" and c.epochs == 3 and c.model.cr == 1.0
    c.update(["--model.cr", "0.5", "--data.root=/tmp/d", "--new.a.b", "[1, 2]", "--amp", "true", "--model.name", "spvcnn_spformer"])
    assert c.model.cr == 0.5 and c.data.root == "/tmp/d" and c.new.a.b == [1, 2] and c.amp is True and c["model"]["name"] == "spvcnn_spformer"
    assert "cr: 0.5" in str(c)
    with pytest.raises(ValueError):
        c.update(["--dangling"])
    d = Config()
    d.load(str(tmp_path / "cfg" / "nusc" / "train" / "x.yaml"))
    assert "workers" not in d and d.epochs == 3


@pytest.mark.timeout(900)
@pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "train_spformer.py")), reason="reference tree not present (GPU box)")
def test_unmodified_train_spformer_script_runs_through_the_launcher():
    """train_spformer.py itself (argument parsing, recursive configs + overrides, builder.make_model / criterion / optimizer /
    scheduler, samplers, DataLoaders, NuScenesTrainer.train_with_defaults + InferenceRunner / MeanIoU / MaxSaver / Saver) via
    u2mkd_b200.shims.launch.run_script with the synthetic dataset and configs/nuscenes/train/spformer.yaml as is apart from sizes:
    one epoch of the reference's own SPVCNN_SPFORMER with its lovasz criterion on CPU (tests/trainer_script_run.py)."""
    r = subprocess.run([sys.executable, os.path.join(HERE, "trainer_script_run.py"), "spformer"], capture_output=True, text=True, timeout=850)
    assert r.returncode == 0, r.stderr[-3000:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert out["checkpoints"] == ["max-iou-val-vox.pt", "step-2.pt"] and out["metainfo"] == ["args.txt", "configs.json"]
    row = out["rows"][-1]
    assert row["epoch_num"] == 1 and row["global_step"] == 2 and np.isfinite(row["total_loss"]) and 0 <= row["iou/val/vox"] <= 100


@pytest.mark.timeout(1500)
@pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "train_lc_nusc_tsd_full.py")), reason="reference tree not present (GPU box)")
def test_unmodified_student_distillation_script_runs_through_the_launcher():
    """train_lc_nusc_tsd_full.py (BASELINE configs[4]: the cross-modal teacher-student run) unchanged: the reference's own
    SPVCNN_SWIFTNET18_SPFORMER_TSD_FULL (SwiftNet image branch, SphereFormer blocks, point<->pixel loops), its
    NuScenesLCTSDFullTrainer (lovasz + KL + feature MSE, teacher logits mapped to the student's voxels through inverse_map ->
    keyframe_mask_full -> inds) and three MeanIoU metrics, over the torchpack stand-in and the synthetic LiDAR + six-camera
    dataset (student / teacher feed_dicts of lc_semantic_nusc_tsd_full.py:436-462).  One epoch on CPU."""
    r = subprocess.run([sys.executable, os.path.join(HERE, "trainer_script_run.py"), "student"], capture_output=True, text=True, timeout=1400)
    assert r.returncode == 0, r.stderr[-3000:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert out["checkpoints"] == ["max-iou-pix-val.pt", "max-iou-vox-val.pt", "step-2.pt"]
    row = out["rows"][-1]
    for k in ("ce/vox", "ce/pix", "ce/kl", "mse/feat", "mse/layer0", "mse/layer3", "total_loss"):
        assert np.isfinite(row[k]), (k, row)
    for k in ("iou-vox/val", "iou-pix/val", "iou-vox-t/val"):
        assert 0 <= row[k] <= 100, (k, row)
    assert row["epoch_num"] == 1 and row["global_step"] == 2


def test_synthetic_camera_dataset_contract(oracle):
    """feed_dict_s / feed_dict_t of lc_semantic_nusc_tsd_full.py:436-462: shapes, the camera masks, and the index chain the
    distillation trainer relies on (teacher voxels -> all teacher points -> keyframe points -> the student's voxels)."""
    saved = {k: sys.modules.get(k) for k in list(sys.modules) if k == "torchsparse" or k.startswith("torchsparse.")}
    oracle.install_as_torchsparse()
    try:
        from u2mkd_b200.shims.synthetic_nusc import SyntheticNuScenesCameras
        ds = SyntheticNuScenesCameras(voxel_size=0.2, num_train=2, num_val=2, multisweeps=2, max_points=4000, image_size=(36, 64), im_drop=3)
        tr, va = ds["train"][0], ds["val"][0]
        assert set(tr) == {"feed_dict_s", "feed_dict_t", "lidar_token"}
        s, t = va["feed_dict_s"], va["feed_dict_t"]
        n = s["num_vox"]
        assert s["images"].shape == (6, 36, 64, 3) and tr["feed_dict_s"]["images"].shape == (3, 36, 64, 3)      # im_drop on train
        assert s["pixel_coordinates"].shape == (6, n, 2) and s["masks"].shape == (6, n) and s["masks"].dtype == bool
        pc, m = s["pixel_coordinates"], s["masks"]
        assert bool((np.abs(pc[m]) < 1).all()) and 0.2 < m.any(0).mean() <= 1.0
        assert np.array_equal(s["fov_mask"].F, m.any(0)) and len(s["inds"]) == 1 and s["inds"][0].shape == (n,)
        assert bool((s["label_fov"].F[s["label_fov"].F > 0] == s["targets_mapped"].F[s["label_fov"].F > 0]).all())
        kf = t["keyframe_mask_full"].F
        n_key = s["targets_mapped"].F.shape[0]
        assert t["num_pts"] == kf.shape[0] and int(kf.sum()) == n_key and bool(kf[:n_key].all())              # keyframe first
        # val split (no augmentation): teacher voxel features reach the student's voxels through the trainer's index chain
        feats_t = t["lidar"].F[t["inverse_map"].F][kf][s["inds"][0]]
        vs_t, vs_s = np.round(feats_t[:, :3] / 0.2), np.round(s["lidar"].F[:, :3] / 0.2)
        assert np.array_equal(vs_t, vs_s)                                                                     # same voxel
        batch = ds["val"].collate_fn([ds["val"][0], ds["val"][1]])
        assert isinstance(batch["feed_dict_s"]["masks"], list) and batch["feed_dict_s"]["images"].shape[:2] == (2, 6)
        assert batch["feed_dict_t"]["num_pts"] == [ds["val"][0]["feed_dict_t"]["num_pts"], ds["val"][1]["feed_dict_t"]["num_pts"]]
    finally:
        for k in [k for k in sys.modules if k == "torchsparse" or k.startswith("torchsparse.")]:
            del sys.modules[k]
        sys.modules.update({k: v for k, v in saved.items() if v is not None})
