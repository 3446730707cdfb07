"""Run by tests/test_trainer_shims_cpu.py in a fresh interpreter: one of the reference's training scripts itself, UNCHANGED,
through u2mkd_b200.shims.launch.run_script — argument parsing, recursive YAML configs + command-line overrides, seeding,
builder.make_* (dataset patched to the synthetic adapter; model / criterion / optimizer / scheduler the reference's own),
samplers and DataLoaders, the reference's Trainer subclass under train_with_defaults with InferenceRunner / MeanIoU / MaxSaver /
Saver.
    argv[1] = "spformer": train_spformer.py with configs/nuscenes/train/spformer.yaml as is, sizes aside (LiDAR only; the reference's
                          SPVCNN_SPFORMER, lovasz criterion, sgd + cosine warm-up schedule)
    argv[1] = "student" : train_lc_nusc_tsd_full.py (teacher-student distillation, LiDAR + six synthetic cameras; the
                          reference's SPVCNN_SWIFTNET18_SPFORMER_TSD_FULL with its SwiftNet image branch, SphereFormer blocks,
                          point<->pixel loops and NuScenesLCTSDFullTrainer)
CPU box: the oracle's torchsparse / sptr namespaces are registered before the launcher runs (it keeps an existing
`torchsparse`), `.cuda()` / set_device are the identity, and --non-dist skips SyncBatchNorm + DDP (CUDA-only).
Prints one JSON line."""
import json
import os
import sys
import tempfile

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, REPO)
which = sys.argv[1] if len(sys.argv) > 1 else "spformer"

import torch  # noqa: E402

from oracle import sptr_oracle, ts_oracle  # noqa: E402  (test infrastructure)

ts = ts_oracle.install_as_torchsparse()
sp = sptr_oracle.as_sptr_module()
sp.__u2_keep__ = True
sys.path.insert(0, REF)
import third_party.SparseTransformer  # noqa: E402,F401  (the checkout's package; its sptr leaf needs the absent sptr_cuda)
sys.modules["third_party.SparseTransformer.sptr"] = sp
torch.Tensor.cuda = lambda self, *a, **k: self
torch.nn.Module.cuda = lambda self, *a, **k: self
ts.SparseTensor.cuda = lambda self, *a, **k: self
torch.cuda.set_device = lambda *a, **k: None
torch.cuda.manual_seed = lambda *a, **k: None

from u2mkd_b200.shims import launch  # noqa: E402

run_dir = tempfile.mkdtemp(prefix=f"u2_train_{which}_")
os.chdir(REF)   # the scripts are run from the checkout root, config paths are relative (README.md:89,101)
common = ["--run-dir", run_dir, "--non-dist", "--dataset.voxel_size", "0.4", "--num_epochs", "1", "--batch_size", "2",
          "--workers_per_gpu", "0", "--optimizer.lr", "0.02", "--data.training_size", "4"]
if which == "spformer":
    launch.run_script(os.path.join(REF, "train_spformer.py"),
                      ["configs/nuscenes/train/spformer.yaml"] + common +
                      [],   # the config's own model (spvcnn_spformer, cr 1.0) and criterion (lovasz + cross entropy)
                      synthetic=(4, 2, 4000))
else:
    launch.run_script(os.path.join(REF, "train_lc_nusc_tsd_full.py"),
                      ["configs/nuscenes/train/spformer_tsd_full_ours_star.yaml"] + common +
                      ["--model.cr", "1.0", "--model.cr_t", "1.0", "--model.in_channel_t", "4", "--model.imagenet_pretrain", "None",
                       "--model.teacher_pretrain", "None", "--model.window_size", "6"],
                      synthetic=(4, 2, 3000, 48, 80))
ck = sorted(os.listdir(os.path.join(run_dir, "checkpoints")))
rows = [json.loads(l) for l in open(os.path.join(run_dir, "summary", "scalars.jsonl"))]
print(json.dumps({"checkpoints": ck, "rows": rows, "metainfo": sorted(os.listdir(os.path.join(run_dir, "metainfo")))}))
