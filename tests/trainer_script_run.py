"""Run by tests/test_trainer_shims_cpu.py in a fresh interpreter: the reference's train_spformer.py itself, UNCHANGED, through
u2mkd_b200.shims.launch.run_script — argument parsing, recursive YAML configs + command-line overrides, seeding, builder.make_*
(dataset patched to the synthetic adapter, model / criterion / optimizer / scheduler the reference's own), samplers and
DataLoaders, NuScenesTrainer.train_with_defaults with InferenceRunner / MeanIoU / MaxSaver / Saver.  CPU box: the oracle's
torchsparse namespace is registered before the launcher runs (it keeps an existing `torchsparse`), `.cuda()` / set_device are
the identity, and --non-dist skips SyncBatchNorm + DDP (CUDA-only).  Prints one JSON line."""
import json
import os
import sys
import tempfile

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, REPO)

import torch  # noqa: E402

from oracle import ts_oracle  # noqa: E402  (test infrastructure)

ts = ts_oracle.install_as_torchsparse()
torch.Tensor.cuda = lambda self, *a, **k: self
torch.nn.Module.cuda = lambda self, *a, **k: self
ts.SparseTensor.cuda = lambda self, *a, **k: self
torch.cuda.set_device = lambda *a, **k: None
torch.cuda.manual_seed = lambda *a, **k: None

from u2mkd_b200.shims import launch  # noqa: E402

run_dir = tempfile.mkdtemp(prefix="u2_train_spformer_")
os.chdir(REF)   # the script is run from its checkout root, config paths are relative (README.md:89)
launch.run_script(os.path.join(REF, "train_spformer.py"),
                  ["configs/nuscenes/train/spformer.yaml", "--run-dir", run_dir, "--non-dist",
                   "--model.name", "spvcnn", "--model.cr", "0.25", "--dataset.voxel_size", "0.4", "--criterion.name", "cross_entropy",
                   "--num_epochs", "1", "--batch_size", "2", "--workers_per_gpu", "0", "--optimizer.lr", "0.05",
                   "--data.training_size", "4"],
                  synthetic=(4, 2, 5000))
ck = sorted(os.listdir(os.path.join(run_dir, "checkpoints")))
rows = [json.loads(l) for l in open(os.path.join(run_dir, "summary", "scalars.jsonl"))]
print(json.dumps({"checkpoints": ck, "rows": rows, "metainfo": sorted(os.listdir(os.path.join(run_dir, "metainfo")))}))
