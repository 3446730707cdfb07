"""Trainer-level stand-in for `torchpack` (github.com/zhijian-liu/torchpack; the reference installs it from git HEAD,
README.md:35 — no version pin, not vendored, absent from this image): SURVEY.md §8(f) row 3.

What train_spformer.py:12-16,57-115, core/spformer_trainer.py:10-14,20-139, core/nusc_trainers.py, core/callbacks.py:8-13
and core/builder.py / core/schedulers.py touch, restated from torchpack's published behaviour:

  torchpack.distributed      init() (env:// rendezvous instead of MPI), size / rank / local_size / local_rank / is_master,
                             barrier, allreduce(data, reduction), allgather(data), broadcast
  torchpack.environ          set_run_dir / auto_set_run_dir / get_run_dir
  torchpack.utils.logging    logger (stdlib logging with .success)
  torchpack.utils.typing     Dataset / Optimizer / Scheduler / Trainer aliases
  torchpack.utils.fs / io    normpath, makedir, remove, io.save / io.load (torch files by extension .pt / .pth, JSON otherwise)
  torchpack.train            Trainer (train / train_with_defaults, the before/after/trigger hook order, state_dict), Summary
  torchpack.callbacks        Callback, Callbacks, LambdaCallback, ProgressBar, EstimatedTimeLeft, InferenceRunner,
                             MaxSaver / MinSaver / Saver, SummaryWriter / ConsoleWriter / TFEventWriter / JSONLWriter,
                             MetaInfoSaver

Host orchestration only: nothing here touches the hot path.  Hook order (the part trainers depend on):
    before_train -> { before_epoch -> [ before_step, run_step, after_step, trigger_step ]* -> after_epoch -> trigger_epoch }*
    -> after_train; the trainer's own _hook runs BEFORE the callbacks' in before_* and AFTER them in after_* / trigger_*.
core/spformer_trainer.py relies on exactly that: its _after_epoch puts the model in eval mode and InferenceRunner then
validates from trigger_epoch."""
from __future__ import annotations

import json
import logging
import os
import sys
import time
from collections import deque
from typing import Any, Dict, List, Optional

import numpy as np
import torch

# ------------------------------------------------------------------ torchpack.utils.logging
logger = logging.getLogger("torchpack")
if not logger.handlers:
    _h = logging.StreamHandler(sys.stderr)
    _h.setFormatter(logging.Formatter("[%(asctime)s] %(message)s", "%H:%M:%S"))
    logger.addHandler(_h)
    logger.setLevel(logging.INFO)
    logger.propagate = False
if not hasattr(logger, "success"):
    logger.success = logger.info  # loguru-style level used by torchpack


# ------------------------------------------------------------------ torchpack.distributed
class _Dist:
    """torchpack.distributed over torch.distributed.  torchpack reads OpenMPI's environment; this reads torchrun's
    (RANK / WORLD_SIZE / LOCAL_RANK, env:// rendezvous) and works un-initialised as a single process."""

    def init(self) -> None:
        import torch.distributed as td
        if td.is_available() and not td.is_initialized() and int(os.environ.get("WORLD_SIZE", "1")) > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29500")
            td.init_process_group("nccl" if torch.cuda.is_available() else "gloo", init_method="env://")

    @staticmethod
    def _on() -> bool:
        import torch.distributed as td
        return td.is_available() and td.is_initialized()

    def size(self) -> int:
        import torch.distributed as td
        return td.get_world_size() if self._on() else int(os.environ.get("WORLD_SIZE", "1"))

    def rank(self) -> int:
        import torch.distributed as td
        return td.get_rank() if self._on() else int(os.environ.get("RANK", "0"))

    def local_size(self) -> int:
        return int(os.environ.get("LOCAL_WORLD_SIZE", str(max(1, torch.cuda.device_count() if torch.cuda.is_available() else 1))))

    def local_rank(self) -> int:
        return int(os.environ.get("LOCAL_RANK", "0"))

    def is_master(self) -> bool:
        return self.rank() == 0

    def barrier(self) -> None:
        import torch.distributed as td
        if self._on() and td.get_world_size() > 1:
            td.barrier()

    def allgather(self, data: Any) -> List[Any]:
        import torch.distributed as td
        if not self._on() or td.get_world_size() == 1:
            return [data]
        out = [None] * td.get_world_size()
        td.all_gather_object(out, data)
        return out

    def allreduce(self, data: Any, reduction: str = "sum") -> Any:
        parts = self.allgather(data)
        if reduction == "sum":
            total = parts[0]
            for p in parts[1:]:
                total = total + p
            return total
        if reduction == "max":
            return max(parts)
        if reduction == "min":
            return min(parts)
        if reduction == "mean":
            return sum(parts) / len(parts)
        raise ValueError(f"unknown reduction {reduction!r}")

    def broadcast(self, data: Any, src: int = 0) -> Any:
        import torch.distributed as td
        if not self._on() or td.get_world_size() == 1:
            return data
        box = [data]
        td.broadcast_object_list(box, src=src)
        return box[0]


distributed = _Dist()

# ------------------------------------------------------------------ torchpack.environ
_run_dir: Optional[str] = None


def get_run_dir() -> str:
    global _run_dir
    if _run_dir is None:
        auto_set_run_dir()
    return _run_dir


def set_run_dir(dirpath: str) -> None:
    global _run_dir
    _run_dir = os.path.normpath(dirpath)
    os.makedirs(_run_dir, exist_ok=True)


def auto_set_run_dir() -> str:
    tags = ["run"]
    if len(sys.argv) > 1 and sys.argv[1].endswith((".yaml", ".yml")):
        tags = [os.path.splitext(os.path.basename(sys.argv[1]))[0]]
    run_dir = os.path.join("runs", "-".join(tags) + time.strftime("-%y%m%d-%H%M%S"))
    set_run_dir(run_dir)
    return run_dir


# ------------------------------------------------------------------ torchpack.utils.fs / io
class fs:  # noqa: N801 - module-like namespace
    @staticmethod
    def normpath(path: str) -> str:
        return os.path.normpath(os.path.realpath(os.path.abspath(os.path.expanduser(path))))

    @staticmethod
    def makedir(path: str) -> None:
        os.makedirs(fs.normpath(path), exist_ok=True)

    @staticmethod
    def remove(path: str) -> None:
        if os.path.exists(path):
            os.remove(path)

    exists = staticmethod(os.path.exists)


class io:  # noqa: N801
    @staticmethod
    def save(fpath: str, obj: Any, **kwargs) -> None:
        fpath = fs.normpath(fpath)
        os.makedirs(os.path.dirname(fpath), exist_ok=True)
        if fpath.endswith((".pt", ".pth", ".pth.tar")):
            torch.save(obj, fpath, **kwargs)
        elif fpath.endswith(".json"):
            with open(fpath, "w") as f:
                json.dump(obj, f, **kwargs)
        elif fpath.endswith(".npy"):
            np.save(fpath, obj)
        else:
            raise NotImplementedError(f"io.save: unsupported extension of {fpath!r}")

    @staticmethod
    def load(fpath: str, **kwargs) -> Any:
        fpath = fs.normpath(fpath)
        if fpath.endswith((".pt", ".pth", ".pth.tar")):
            kwargs.setdefault("map_location", "cpu")
            kwargs.setdefault("weights_only", False)
            return torch.load(fpath, **kwargs)
        if fpath.endswith(".json"):
            with open(fpath) as f:
                return json.load(f)
        if fpath.endswith(".npy"):
            return np.load(fpath)
        raise NotImplementedError(f"io.load: unsupported extension of {fpath!r}")


# ------------------------------------------------------------------ torchpack.callbacks
class Callback:
    """Base class.  `master_only` callbacks become no-ops on the other ranks.  Public hooks call the underscore hooks that
    subclasses override; `trainer` is set by set_trainer before training starts."""
    master_only: bool = False

    def __new__(cls, *args, **kwargs):
        if cls.master_only and not distributed.is_master():
            return object.__new__(LambdaCallback)
        return object.__new__(cls)

    def set_trainer(self, trainer) -> None:
        self.trainer = trainer
        self._set_trainer(trainer)

    def _set_trainer(self, trainer) -> None: ...

    def before_train(self) -> None: self._before_train()
    def _before_train(self) -> None: ...
    def before_epoch(self) -> None: self._before_epoch()
    def _before_epoch(self) -> None: ...
    def before_step(self, feed_dict: Dict[str, Any]) -> None: self._before_step(feed_dict)
    def _before_step(self, feed_dict: Dict[str, Any]) -> None: ...
    def after_step(self, output_dict: Dict[str, Any]) -> None: self._after_step(output_dict)
    def _after_step(self, output_dict: Dict[str, Any]) -> None: ...
    def trigger_step(self) -> None: self._trigger_step()
    def _trigger_step(self) -> None: ...
    def after_epoch(self) -> None: self._after_epoch()
    def _after_epoch(self) -> None: ...
    def trigger_epoch(self) -> None: self._trigger_epoch()
    def _trigger_epoch(self) -> None: ...
    def trigger(self) -> None: self._trigger()
    def _trigger(self) -> None: ...
    def after_train(self) -> None: self._after_train()
    def _after_train(self) -> None: ...

    def state_dict(self) -> Optional[Dict[str, Any]]: return self._state_dict()
    def _state_dict(self) -> Optional[Dict[str, Any]]: return None
    def load_state_dict(self, state_dict: Dict[str, Any]) -> None: self._load_state_dict(state_dict)
    def _load_state_dict(self, state_dict: Dict[str, Any]) -> None: ...

    def __str__(self) -> str:
        return type(self).__name__


class LambdaCallback(Callback):
    """Callback from plain functions (and the inert replacement of master_only callbacks on non-master ranks)."""

    def __init__(self, **fns) -> None:
        self._fns = fns

    def _call(self, name, *args):
        fn = getattr(self, "_fns", {}).get(name)
        if fn is not None:
            fn(self, *args)

    def _before_train(self): self._call("before_train")
    def _before_epoch(self): self._call("before_epoch")
    def _before_step(self, feed_dict): self._call("before_step", feed_dict)
    def _after_step(self, output_dict): self._call("after_step", output_dict)
    def _trigger_step(self): self._call("trigger_step")
    def _after_epoch(self): self._call("after_epoch")
    def _trigger_epoch(self): self._call("trigger_epoch")
    def _trigger(self): self._call("trigger")
    def _after_train(self): self._call("after_train")


class Callbacks(Callback):
    """A list of callbacks behaving as one."""

    def __init__(self, callbacks: List[Callback]) -> None:
        for cb in callbacks:
            assert isinstance(cb, Callback), type(cb)
        self.callbacks = list(callbacks)

    def _set_trainer(self, trainer) -> None:
        for cb in self.callbacks:
            cb.set_trainer(trainer)

    def _before_train(self):
        for cb in self.callbacks: cb.before_train()

    def _before_epoch(self):
        for cb in self.callbacks: cb.before_epoch()

    def _before_step(self, feed_dict):
        for cb in self.callbacks: cb.before_step(feed_dict)

    def _after_step(self, output_dict):
        for cb in self.callbacks: cb.after_step(output_dict)

    def _trigger_step(self):
        for cb in self.callbacks: cb.trigger_step()

    def _after_epoch(self):
        for cb in self.callbacks: cb.after_epoch()

    def _trigger_epoch(self):
        for cb in self.callbacks: cb.trigger_epoch()

    def _trigger(self):
        for cb in self.callbacks: cb.trigger()

    def _after_train(self):
        for cb in self.callbacks: cb.after_train()

    def _state_dict(self):
        out = {}
        for k, cb in enumerate(self.callbacks):
            sd = cb.state_dict()
            if sd:
                out[f"{type(cb).__name__}.{k}"] = sd
        return out

    def _load_state_dict(self, state_dict):
        for k, cb in enumerate(self.callbacks):
            sd = state_dict.get(f"{type(cb).__name__}.{k}")
            if sd:
                cb.load_state_dict(sd)

    def __getitem__(self, i): return self.callbacks[i]
    def __len__(self): return len(self.callbacks)


class ProgressBar(Callback):
    """tqdm bar over the steps of an epoch (master only; silent when stderr is not a terminal)."""
    master_only = True

    def __init__(self, scalars: Optional[Any] = None) -> None:
        self.scalars = scalars
        self.pbar = None

    def _before_epoch(self) -> None:
        if sys.stderr.isatty():
            import tqdm
            self.pbar = tqdm.trange(self.trainer.steps_per_epoch, ncols=0)

    def _trigger_step(self) -> None:
        if self.pbar is not None:
            self.pbar.update()

    def _after_epoch(self) -> None:
        if self.pbar is not None:
            self.pbar.close()
            self.pbar = None


class EstimatedTimeLeft(Callback):
    master_only = True

    def _before_train(self) -> None:
        self.times = deque(maxlen=8)
        self.last = time.perf_counter()

    def _trigger_epoch(self) -> None:
        now = time.perf_counter()
        self.times.append(now - self.last)
        self.last = now
        left = (self.trainer.num_epochs - self.trainer.epoch_num) * float(np.mean(self.times))
        if left > 0:
            logger.info(f"Estimated time left: {left / 60:.1f} min.")


class InferenceRunner(Callback):
    """After every training epoch (trigger_epoch): run `dataflow` through trainer.run_step under no_grad and feed the
    outputs to its own callbacks (train_spformer.py:101-111)."""

    def __init__(self, dataflow, *, callbacks: List[Callback]) -> None:
        self.dataflow = dataflow
        self.callbacks = Callbacks(callbacks)

    def _set_trainer(self, trainer) -> None:
        self.callbacks.set_trainer(trainer)

    def _trigger_epoch(self) -> None:
        self._trigger()

    def _trigger(self) -> None:
        t0 = time.perf_counter()
        self.callbacks.before_epoch()
        with torch.no_grad():
            for feed_dict in self.dataflow:
                self.callbacks.before_step(feed_dict)
                output_dict = self.trainer.run_step(feed_dict)
                self.callbacks.after_step(output_dict)
        self.callbacks.after_epoch()
        logger.info(f"Inference finished in {time.perf_counter() - t0:.1f} s.")


class Saver(Callback):
    """Checkpoint `trainer.state_dict()` to <run_dir>/checkpoints/step-<global_step>.pt every epoch, keeping the last
    `max_to_keep`."""
    master_only = True

    def __init__(self, *, max_to_keep: int = 4, save_dir: Optional[str] = None) -> None:
        self.max_to_keep = max_to_keep
        self.save_dir = fs.normpath(save_dir if save_dir is not None else os.path.join(get_run_dir(), "checkpoints"))
        self.checkpoints = deque()

    def _trigger_epoch(self) -> None:
        self._trigger()

    def _trigger(self) -> None:
        path = os.path.join(self.save_dir, f"step-{self.trainer.global_step}.pt")
        try:
            io.save(path, self.trainer.state_dict())
        except OSError:
            logger.exception(f'Error occurred when saving checkpoint "{path}".')
            return
        logger.info(f'Checkpoint saved: "{path}".')
        self.checkpoints.append(path)
        while self.max_to_keep is not None and len(self.checkpoints) > self.max_to_keep:
            fs.remove(self.checkpoints.popleft())


class _BestSaver(Callback):
    master_only = True
    extreme = "max"

    def __init__(self, scalar: str, *, name: Optional[str] = None, save_dir: Optional[str] = None) -> None:
        self.scalar = scalar
        self.name = name if name is not None else scalar.replace("/", "-")
        self.save_dir = fs.normpath(save_dir if save_dir is not None else os.path.join(get_run_dir(), "checkpoints"))
        self.step = None
        self.best = None

    def _trigger_epoch(self) -> None:
        self._trigger()

    def _trigger(self) -> None:
        if self.scalar not in self.trainer.summary:
            logger.warning(f'`{self.scalar}` has not been added to `trainer.summary`.')
            return
        step, value = self.trainer.summary[self.scalar][-1]
        if self.step is not None and step <= self.step:
            logger.warning(f'`{self.scalar}` has not been updated since the last trigger.')
            return
        self.step = step
        better = self.best is None or (value > self.best[1] if self.extreme == "max" else value < self.best[1])
        if better:
            self.best = (step, value)
            path = os.path.join(self.save_dir, f"{self.extreme}-{self.name}.pt")
            try:
                io.save(path, self.trainer.state_dict())
            except OSError:
                logger.exception(f'Error occurred when saving checkpoint "{path}".')
            else:
                logger.info(f'Checkpoint saved: "{path}" ({value:.5g}).')
        if self.best is not None:
            self.trainer.summary.add_scalar(self.scalar + "/" + self.extreme, self.best[1])

    def _state_dict(self):
        return {"step": self.step, "best": self.best}

    def _load_state_dict(self, state_dict):
        self.step, self.best = state_dict["step"], state_dict["best"]


class MaxSaver(_BestSaver):
    extreme = "max"


class MinSaver(_BestSaver):
    extreme = "min"


class SummaryWriter(Callback):
    """Receives every scalar added to trainer.summary."""
    master_only = True

    def add_scalar(self, name: str, scalar: float) -> None:
        self._add_scalar(name, scalar)

    def _add_scalar(self, name: str, scalar: float) -> None: ...

    def add_image(self, name: str, tensor) -> None: ...


class ConsoleWriter(SummaryWriter):
    """Prints the scalars of an epoch when it ends."""

    def __init__(self, scalars="*") -> None:
        self.scalars_filter = scalars

    def _set_trainer(self, trainer) -> None:
        self.scalars = {}

    def _add_scalar(self, name, scalar) -> None:
        self.scalars[name] = scalar

    def _trigger_epoch(self) -> None:
        self._trigger()

    def _trigger(self) -> None:
        if self.scalars:
            logger.info("\n+ " + "\n+ ".join(f"[{k}] = {v:.5g}" for k, v in sorted(self.scalars.items())))
            self.scalars.clear()


class TFEventWriter(SummaryWriter):
    """torchpack writes TensorBoard event files; here the scalars go to <run_dir>/summary/scalars.jsonl (no tensorboard
    dependency).  core/callbacks.py:150-153 only uses add_scalar and the isinstance check."""

    def __init__(self, *, save_dir: Optional[str] = None) -> None:
        self.save_dir = save_dir

    def _set_trainer(self, trainer) -> None:
        d = fs.normpath(self.save_dir if self.save_dir is not None else os.path.join(get_run_dir(), "tensorboard"))
        os.makedirs(d, exist_ok=True)
        self._path = os.path.join(d, "scalars.jsonl")

    def _add_scalar(self, name, scalar) -> None:
        with open(self._path, "a") as f:
            f.write(json.dumps({"step": getattr(self.trainer, "global_step", 0), "name": name, "value": float(scalar)}) + "\n")


class JSONLWriter(SummaryWriter):
    def __init__(self, save_dir: Optional[str] = None) -> None:
        self.save_dir = save_dir

    def _set_trainer(self, trainer) -> None:
        d = fs.normpath(self.save_dir if self.save_dir is not None else os.path.join(get_run_dir(), "summary"))
        os.makedirs(d, exist_ok=True)
        self._path = os.path.join(d, "scalars.jsonl")
        self._row = {}

    def _add_scalar(self, name, scalar) -> None:
        self._row[name] = float(scalar)

    def _trigger_epoch(self) -> None:
        if self._row:
            with open(self._path, "a") as f:
                f.write(json.dumps(dict(self._row, epoch_num=self.trainer.epoch_num, global_step=self.trainer.global_step)) + "\n")
            self._row = {}


class MetaInfoSaver(Callback):
    """<run_dir>/metainfo/{configs.json, args.txt}."""
    master_only = True

    def _before_train(self) -> None:
        d = os.path.join(get_run_dir(), "metainfo")
        os.makedirs(d, exist_ok=True)
        try:
            from . import configs as cfg
            with open(os.path.join(d, "configs.json"), "w") as f:
                json.dump(cfg, f, indent=1, default=str)
        except Exception:  # noqa: BLE001 - metadata only
            pass
        with open(os.path.join(d, "args.txt"), "w") as f:
            f.write(" ".join(sys.argv) + "\n")


# ------------------------------------------------------------------ torchpack.train
class Summary:
    """trainer.summary: per-name history of (global_step, value) + fan-out to the SummaryWriter callbacks."""

    def __init__(self) -> None:
        self.history: Dict[str, deque] = {}
        self.writers: List[SummaryWriter] = []
        self.trainer = None

    def set_trainer(self, trainer) -> None:
        self.trainer = trainer
        self.writers = []
        stack = list(trainer.callbacks.callbacks)
        while stack:
            cb = stack.pop()
            if isinstance(cb, SummaryWriter):
                self.writers.append(cb)
            elif isinstance(cb, Callbacks):
                stack.extend(cb.callbacks)

    def add_scalar(self, name: str, scalar: Any, *, max_to_keep: Optional[int] = None) -> None:
        if isinstance(scalar, torch.Tensor):
            scalar = scalar.item()
        scalar = float(scalar)
        hist = self.history.get(name)
        if hist is None:
            hist = self.history[name] = deque(maxlen=max_to_keep)
        hist.append((getattr(self.trainer, "global_step", 0), scalar))
        for w in self.writers:
            w.add_scalar(name, scalar)

    def add_image(self, name: str, tensor, *, max_to_keep: Optional[int] = None) -> None:
        for w in self.writers:
            w.add_image(name, tensor)

    def get(self, name, default=None):
        return self.history.get(name, default)

    def __getitem__(self, name): return self.history[name]
    def keys(self): return self.history.keys()
    def __contains__(self, name): return name in self.history


class Trainer:
    """torchpack.train.Trainer: the epoch / step loop around the subclass's _run_step."""

    def train_with_defaults(self, dataflow, *, num_epochs: int = 9999999, callbacks: Optional[List[Callback]] = None) -> None:
        callbacks = list(callbacks or [])
        callbacks += [MetaInfoSaver(), ConsoleWriter(), TFEventWriter(), JSONLWriter(), ProgressBar(), EstimatedTimeLeft()]
        self.train(dataflow=dataflow, num_epochs=num_epochs, callbacks=callbacks)

    def train(self, dataflow, *, num_epochs: int = 9999999, callbacks: Optional[List[Callback]] = None) -> None:
        self.dataflow = dataflow
        self.steps_per_epoch = len(dataflow)
        self.num_epochs = num_epochs
        self.callbacks = Callbacks(list(callbacks or []))
        self.summary = Summary()
        try:
            self.callbacks.set_trainer(self)
            self.summary.set_trainer(self)
            self.epoch_num = 0
            self.global_step = 0
            t_train = time.perf_counter()
            self.before_train()
            while self.epoch_num < self.num_epochs:
                self.epoch_num += 1
                self.local_step = 0
                logger.info(f"Epoch {self.epoch_num}/{self.num_epochs} started.")
                t_epoch = time.perf_counter()
                self.before_epoch()
                for feed_dict in self.dataflow:
                    self.local_step += 1
                    self.global_step += 1
                    self.before_step(feed_dict)
                    output_dict = self.run_step(feed_dict)
                    self.after_step(output_dict)
                    self.trigger_step()
                self.after_epoch()
                logger.info(f"Training finished in {time.perf_counter() - t_epoch:.1f} s.")
                self.trigger_epoch()
                logger.info(f"Epoch finished in {time.perf_counter() - t_epoch:.1f} s.")
            logger.success(f"{self.num_epochs} epochs of training finished in {time.perf_counter() - t_train:.1f} s.")
        except StopTraining as e:
            logger.info(f"Training was stopped by {e}.")
        finally:
            self.after_train()

    # trainer hook first in before_*, callbacks first in after_* / trigger_*
    def before_train(self) -> None:
        self._before_train()
        self.callbacks.before_train()

    def _before_train(self) -> None: ...

    def before_epoch(self) -> None:
        self._before_epoch()
        self.callbacks.before_epoch()

    def _before_epoch(self) -> None: ...

    def before_step(self, feed_dict: Dict[str, Any]) -> None:
        self._before_step(feed_dict)
        self.callbacks.before_step(feed_dict)

    def _before_step(self, feed_dict: Dict[str, Any]) -> None: ...

    def run_step(self, feed_dict: Dict[str, Any]) -> Dict[str, Any]:
        return self._run_step(feed_dict)

    def _run_step(self, feed_dict: Dict[str, Any]) -> Dict[str, Any]:
        raise NotImplementedError

    def after_step(self, output_dict: Dict[str, Any]) -> None:
        self.callbacks.after_step(output_dict)
        self._after_step(output_dict)

    def _after_step(self, output_dict: Dict[str, Any]) -> None: ...

    def trigger_step(self) -> None:
        self.callbacks.trigger_step()
        self._trigger_step()

    def _trigger_step(self) -> None: ...

    def after_epoch(self) -> None:
        self.callbacks.after_epoch()
        self._after_epoch()

    def _after_epoch(self) -> None: ...

    def trigger_epoch(self) -> None:
        self.callbacks.trigger_epoch()
        self._trigger_epoch()

    def _trigger_epoch(self) -> None: ...

    def trigger(self) -> None:
        self.callbacks.trigger()
        self._trigger()

    def _trigger(self) -> None: ...

    def after_train(self) -> None:
        self.callbacks.after_train()
        self._after_train()

    def _after_train(self) -> None: ...

    def state_dict(self) -> Dict[str, Any]:
        sd = self._state_dict()
        sd["callbacks"] = self.callbacks.state_dict()
        sd["epoch_num"] = self.epoch_num
        sd["local_step"] = self.local_step
        sd["global_step"] = self.global_step
        return sd

    def _state_dict(self) -> Dict[str, Any]:
        return {}

    def load_state_dict(self, state_dict: Dict[str, Any]) -> None:
        self.epoch_num = state_dict.pop("epoch_num", getattr(self, "epoch_num", 0))
        self.local_step = state_dict.pop("local_step", 0)
        self.global_step = state_dict.pop("global_step", getattr(self, "global_step", 0))
        cbs = state_dict.pop("callbacks", None)
        if cbs and hasattr(self, "callbacks"):
            self.callbacks.load_state_dict(cbs)
        self._load_state_dict(state_dict)

    def _load_state_dict(self, state_dict: Dict[str, Any]) -> None: ...


class StopTraining(Exception):
    pass


# ------------------------------------------------------------------ small third-party stand-ins core/callbacks.py imports
class PrettyTable:
    """prettytable.PrettyTable in the small (core/callbacks.py:154-157): field_names, add_row, str()."""

    def __init__(self, field_names: Optional[List[str]] = None) -> None:
        self.field_names = list(field_names or [])
        self.rows: List[List[Any]] = []

    def add_row(self, row) -> None:
        self.rows.append(list(row))

    def __str__(self) -> str:
        cells = [[str(c) for c in self.field_names]] + [[str(c) for c in r] for r in self.rows]
        n = max(len(r) for r in cells) if cells else 0
        cells = [r + [""] * (n - len(r)) for r in cells]
        w = [max(len(r[i]) for r in cells) for i in range(n)]
        bar = "+" + "+".join("-" * (x + 2) for x in w) + "+"
        out = [bar]
        for k, r in enumerate(cells):
            out.append("|" + "|".join(" " + c.center(x) + " " for c, x in zip(r, w)) + "|")
            if k == 0:
                out.append(bar)
        out.append(bar)
        return "\n".join(out)


class ConfusionMatrix:
    """nuscenes.eval.lidarseg.utils.ConfusionMatrix in the small (imported by core/callbacks.py:14, unused by MeanIoU):
    num_classes x num_classes counts, per-class IoU with an ignored index."""

    def __init__(self, num_classes: int, ignore_idx: Optional[int] = None) -> None:
        self.num_classes = num_classes
        self.ignore_idx = ignore_idx
        self.global_cm = None

    def update(self, gt_array, pred_array) -> np.ndarray:
        gt, pred = np.asarray(gt_array).astype(np.int64), np.asarray(pred_array).astype(np.int64)
        cm = np.bincount(self.num_classes * gt + pred, minlength=self.num_classes ** 2).reshape(self.num_classes, self.num_classes)
        self.global_cm = cm if self.global_cm is None else self.global_cm + cm
        return cm

    def get_per_class_iou(self) -> List[float]:
        conf = self.global_cm.copy().astype(np.float64)
        if self.ignore_idx is not None:
            conf[self.ignore_idx, :] = 0
            conf[:, self.ignore_idx] = 0
        tp = np.diagonal(conf)
        denom = conf.sum(1) + conf.sum(0) - tp
        with np.errstate(divide="ignore", invalid="ignore"):
            iou = np.where(denom > 0, tp / denom, np.nan)
        if self.ignore_idx is not None:
            iou[self.ignore_idx] = np.nan
        return iou.tolist()

    def get_mean_iou(self) -> float:
        return float(np.nanmean(np.array(self.get_per_class_iou())))
