"""Run by tests/test_trainer_shims_cpu.py in a fresh interpreter (the reference modules bind `torchsparse` at import time, so
this must not share sys.modules with tests that install the product namespace): the reference's OWN trainer and metric
classes — core/spformer_trainer.py NuScenesTrainer, core/callbacks.py MeanIoU, unmodified — driven exactly as
train_spformer.py:57-115 drives them, over the torchpack stand-in, the synthetic dataset adapter, and (CPU box: no CUDA)
the oracle's torchsparse namespace with `.cuda()` turned into the identity.  Prints one JSON line."""
import json
import os
import sys
import tempfile

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, REPO)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import ts_oracle  # noqa: E402  (test infrastructure)

ts = ts_oracle.install_as_torchsparse()
import u2mkd_b200  # noqa: E402

sys.path.insert(0, REF)   # before the shims: a real `third_party` package (the checkout) must win over the stand-in
names = u2mkd_b200.install_reference_shims()

# a CPU run of code that says .cuda(): identity
torch.Tensor.cuda = lambda self, *a, **k: self
ts.SparseTensor.cuda = lambda self, *a, **k: self

from torchpack import distributed as dist  # noqa: E402
from torchpack.callbacks import InferenceRunner, MaxSaver, Saver  # noqa: E402
from torchpack.environ import get_run_dir, set_run_dir  # noqa: E402
from torchpack.utils.config import configs  # noqa: E402

run_dir = tempfile.mkdtemp(prefix="u2_trainer_shim_")
set_run_dir(run_dir)
configs.update({"criterion": {"name": "cross_entropy", "ignore_index": 0},
                "dataset": {"name": "semantic_nusc", "multisweeps": {"num_sweeps": 2}},
                "data": {"num_classes": 17, "ignore_label": 0},
                "model": {"cr": 0.25, "in_channel": 4}, "workers_per_gpu": 0, "num_epochs": 2, "batch_size": 2,
                "amp_enabled": False, "train": {"seed": 7}})

from core.callbacks import MeanIoU  # noqa: E402   (reference file, unchanged)
from core.spformer_trainer import NuScenesTrainer  # noqa: E402   (reference file, unchanged)

from u2mkd_b200 import models  # noqa: E402
from u2mkd_b200.shims.synthetic_nusc import SyntheticNuScenes  # noqa: E402

seed = configs.train.seed + dist.rank() * configs.workers_per_gpu * configs.num_epochs       # train_spformer.py:54-58
np.random.seed(seed)
torch.manual_seed(seed)

dataset = SyntheticNuScenes(voxel_size=0.4, num_train=4, num_val=2, multisweeps=configs.dataset.multisweeps.num_sweeps,
                            max_points=5000)
dataflow = {}
for split in dataset:                                                                      # train_spformer.py:61-75
    sampler = torch.utils.data.distributed.DistributedSampler(dataset[split], num_replicas=dist.size(), rank=dist.rank(),
                                                              shuffle=(split == "train"))
    dataflow[split] = torch.utils.data.DataLoader(dataset[split], batch_size=configs.batch_size, sampler=sampler,
                                                  num_workers=configs.workers_per_gpu, collate_fn=dataset[split].collate_fn)

fam = models.build_family(ts)
model = fam.SPVCNN(cr=configs.model.cr, pres=0.4, vres=0.4, num_classes=configs.data.num_classes)
criterion = torch.nn.CrossEntropyLoss(ignore_index=configs.criterion.ignore_index)
optimizer = torch.optim.SGD(model.parameters(), lr=0.05, momentum=0.9, nesterov=True, weight_decay=1e-4)
scheduler = torch.optim.lr_scheduler.LambdaLR(optimizer, lambda it: 1.0)
w0 = model.classifier_vox[0].weight.detach().clone() if hasattr(model.classifier_vox, "__getitem__") else None

trainer = NuScenesTrainer(model=model, criterion=criterion, optimizer=optimizer, scheduler=scheduler,
                          num_workers=configs.workers_per_gpu, seed=seed, weight_path=None, amp_enabled=configs.amp_enabled)
order = []
from torchpack.callbacks import LambdaCallback  # noqa: E402
probe = LambdaCallback(before_epoch=lambda cb: order.append(("before_epoch", cb.trainer.model.training)),
                       after_epoch=lambda cb: order.append(("after_epoch", cb.trainer.model.training)),
                       trigger_epoch=lambda cb: order.append(("trigger_epoch", cb.trainer.model.training)))
trainer.train_with_defaults(                                                               # train_spformer.py:98-115
    dataflow["train"], num_epochs=configs.num_epochs,
    callbacks=[probe] + [InferenceRunner(dataflow[split], callbacks=[MeanIoU(
        name=f"iou/{split}/vox", num_classes=configs.data.num_classes, ignore_label=configs.data.ignore_label,
        output_tensor="outputs_vox", target_tensor="targets")]) for split in ["val"]] + [MaxSaver("iou/val/vox"), Saver(max_to_keep=1)])

ckpt_dir = os.path.join(get_run_dir(), "checkpoints")
ckpts = sorted(os.listdir(ckpt_dir))
from torchpack.utils import io  # noqa: E402
state = io.load(os.path.join(ckpt_dir, [c for c in ckpts if c.startswith("step-")][-1]))
print(json.dumps({
    "shims": names,
    "losses": [v for _, v in trainer.summary["total_loss"]],
    "miou": [v for _, v in trainer.summary["iou/val/vox"]],
    "global_step": trainer.global_step, "epoch_num": trainer.epoch_num,
    "order": order,
    "checkpoints": ckpts,
    "state_keys": sorted(state.keys()),
    "state_steps": [state["epoch_num"], state["global_step"]],
    "summary_files": sorted(os.listdir(os.path.join(get_run_dir(), "summary"))) if os.path.isdir(os.path.join(get_run_dir(), "summary")) else [],
    "moved": bool(w0 is not None and not torch.equal(w0, model.classifier_vox[0].weight.detach())),
}))
