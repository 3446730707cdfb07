"""Host shims (the part the model files need): minimal stand-ins for third-party Python packages that
core/models/nuscenes/spvcnn_spformer.py and core/models/sphereformer/*.py import but this path never depended on for
arithmetic — timm.models.layers (DropPath, trunc_normal_), torch_scatter (scatter_mean / scatter_add / scatter_max),
torchpack.utils.config (the global `configs` mapping) — and the import path third_party.SparseTransformer.sptr.

    import u2mkd_b200
    u2mkd_b200.install_as_torchsparse(); u2mkd_b200.install_reference_shims()
    from core.models.nuscenes.spvcnn_spformer import SPVCNN_SPFORMER      # reference file, unchanged

A shim is only registered when the real package is NOT importable.  Trainer-level torchpack (Trainer, callbacks,
distributed, environ, logging, fs / io: torchpack_shim.py), prettytable, the one nuscenes-devkit class core/callbacks.py
imports and visualize_utils (open3d) come with it, so that train_spformer.py's own classes — core.spformer_trainer.
NuScenesTrainer, core.callbacks.MeanIoU — run unchanged; synthetic_nusc.py supplies the dataset they consume."""
import importlib
import importlib.util
import sys
import types

import torch
import torch.nn as nn


# ------------------------------------------------------------------ timm.models.layers
class DropPath(nn.Module):
    """Stochastic depth per sample (timm.models.layers.DropPath): rows of the batch are dropped with prob drop_prob and
    the survivors rescaled by 1 / keep_prob; identity in eval mode."""

    def __init__(self, drop_prob: float = 0., scale_by_keep: bool = True):
        super().__init__()
        self.drop_prob = drop_prob
        self.scale_by_keep = scale_by_keep

    def forward(self, x):
        if self.drop_prob == 0. or not self.training:
            return x
        keep_prob = 1 - self.drop_prob
        shape = (x.shape[0],) + (1,) * (x.ndim - 1)
        mask = x.new_empty(shape).bernoulli_(keep_prob)
        if keep_prob > 0.0 and self.scale_by_keep:
            mask.div_(keep_prob)
        return x * mask

    def extra_repr(self):
        return f"drop_prob={round(self.drop_prob, 3):0.3f}"


def trunc_normal_(tensor, mean=0., std=1., a=-2., b=2.):
    return nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)


# ------------------------------------------------------------------ torch_scatter (the three reductions the models touch)
def _out(src, index, dim, dim_size, fill=0.0):
    if dim < 0:
        dim += src.dim()
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() else 0
    shape = list(src.shape)
    shape[dim] = dim_size
    return src.new_full(shape, fill), dim


def _bcast(index, src, dim):
    if index.dim() == src.dim():
        return index
    view = [1] * src.dim()
    view[dim] = -1
    return index.view(view).expand_as(src)


def scatter_add(src, index, dim=-1, out=None, dim_size=None):
    o, dim = _out(src, index, dim, dim_size) if out is None else (out, dim if dim >= 0 else dim + src.dim())
    return o.scatter_add_(dim, _bcast(index.long(), src, dim), src)


scatter_sum = scatter_add


def scatter_mean(src, index, dim=-1, out=None, dim_size=None):
    o = scatter_add(src, index, dim, out, dim_size)
    d = dim if dim >= 0 else dim + src.dim()
    cnt = scatter_add(torch.ones_like(src), index, dim, None, o.shape[d])
    return o / cnt.clamp_min(1)


def scatter_max(src, index, dim=-1, out=None, dim_size=None):
    o, dim = _out(src, index, dim, dim_size, fill=float("-inf"))
    idx = _bcast(index.long(), src, dim)
    o = o.scatter_reduce(dim, idx, src, "amax", include_self=True)
    # argmax: first position attaining the maximum
    pos = torch.arange(src.shape[dim], device=src.device)
    view = [1] * src.dim()
    view[dim] = -1
    pos = pos.view(view).expand_as(src)
    hit = src == o.gather(dim, idx)
    arg = torch.full_like(o, src.shape[dim], dtype=torch.long).scatter_reduce(dim, idx, torch.where(hit, pos, src.shape[dim]), "amin",
                                                                              include_self=True)
    o = torch.where(torch.isinf(o) & (o < 0), torch.zeros_like(o), o)
    return o, arg


# ------------------------------------------------------------------ torchpack.utils.config
class Config(dict):
    """torchpack.utils.config.Config in the small: a dict with attribute access, nested dicts converted on the way in, and
    `load` / `update` for YAML files and overrides (what `configs['model']['cr']` / `configs.model.cr` need)."""

    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError as e:
            raise AttributeError(key) from e

    def __setattr__(self, key, value):
        self[key] = value

    def __setitem__(self, key, value):
        super().__setitem__(key, Config(value) if isinstance(value, dict) and not isinstance(value, Config) else value)

    def __init__(self, *args, **kwargs):
        super().__init__()
        for k, v in dict(*args, **kwargs).items():
            self[k] = v

    def update(self, other=(), **kwargs):
        """dict: merged key by key (nested dicts recursively).  list of command-line strings (train_spformer.py:33-34,
        `configs.update(opts)`): `--a.b.c value` / `--a.b.c=value` pairs, values read as Python literals when they parse."""
        if isinstance(other, (list, tuple)) and all(isinstance(o, str) for o in other):
            self._update_from_args(list(other))
            other = ()
        for k, v in dict(other, **kwargs).items():
            if isinstance(v, dict) and isinstance(self.get(k), Config):
                self[k].update(v)
            else:
                self[k] = v

    def _update_from_args(self, opts):
        import ast
        i = 0
        while i < len(opts):
            opt = opts[i]
            if not opt.startswith("--"):
                raise ValueError(f"config override {opt!r}: expected --key[.subkey] value")
            if "=" in opt:
                key, raw = opt[2:].split("=", 1)
                i += 1
            else:
                if i + 1 >= len(opts):
                    raise ValueError(f"config override {opt!r} has no value")
                key, raw = opt[2:], opts[i + 1]
                i += 2
            try:
                value = ast.literal_eval(raw)
            except (ValueError, SyntaxError):
                value = {"true": True, "false": False, "none": None, "null": None}.get(raw.lower(), raw)
            node = self
            parts = key.split(".")
            for part in parts[:-1]:
                if not isinstance(node.get(part), Config):
                    node[part] = Config()
                node = node[part]
            node[parts[-1]] = value

    def load(self, fpath, recursive=False):
        """YAML file into this mapping.  recursive=True (train_spformer.py:32): first every `default.yaml` found on the way
        from the top-most directory of the path down to the file's own directory, then the file itself."""
        import os
        import yaml
        if not os.path.exists(fpath):
            raise FileNotFoundError(fpath)
        fpaths = [fpath]
        if recursive:
            extension = os.path.splitext(fpath)[1]
            d = fpath
            while os.path.dirname(d) != d:
                d = os.path.dirname(d)
                fpaths.append(os.path.join(d, "default" + extension))
            fpaths = [fp for fp in reversed(fpaths) if os.path.exists(fp)]
        for fp in fpaths:
            with open(fp) as f:
                self.update(yaml.safe_load(f) or {})

    def __str__(self):
        def lines(node, indent):
            out = []
            for k, v in node.items():
                if isinstance(v, dict):
                    out.append(" " * indent + f"{k}:")
                    out += lines(v, indent + 2)
                else:
                    out.append(" " * indent + f"{k}: {v}")
            return out
        return "\n".join(lines(self, 0))


configs = Config()


def _missing(name: str) -> bool:
    if name in sys.modules:
        return False
    try:
        return importlib.util.find_spec(name) is None
    except (ImportError, ValueError):
        return True


def _module(name: str, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__u2_shim__ = True
    sys.modules[name] = m
    return m


def install_reference_shims() -> list:
    """Register the shims (only for packages that are not importable) and alias third_party.SparseTransformer.sptr to
    u2mkd_b200.sptr.  Returns the names registered."""
    done = []
    if _missing("timm"):
        layers = _module("timm.models.layers", DropPath=DropPath, trunc_normal_=trunc_normal_)
        models = _module("timm.models", layers=layers)
        _module("timm", models=models)
        done.append("timm.models.layers")
    if _missing("torch_scatter"):
        _module("torch_scatter", scatter_mean=scatter_mean, scatter_add=scatter_add, scatter_sum=scatter_sum, scatter_max=scatter_max)
        done.append("torch_scatter")
    if _missing("torchpack"):
        from . import torchpack_shim as tp
        cfg = _module("torchpack.utils.config", Config=Config, configs=configs)
        log = _module("torchpack.utils.logging", logger=tp.logger)
        typ = _module("torchpack.utils.typing", Dataset=torch.utils.data.Dataset, Optimizer=torch.optim.Optimizer,
                      Scheduler=torch.optim.lr_scheduler.LRScheduler, Trainer=tp.Trainer)
        fsm = _module("torchpack.utils.fs", **{k: getattr(tp.fs, k) for k in ("normpath", "makedir", "remove", "exists")})
        iom = _module("torchpack.utils.io", save=tp.io.save, load=tp.io.load)
        utils = _module("torchpack.utils", config=cfg, logging=log, typing=typ, fs=fsm, io=iom)
        d = tp.distributed
        dist = _module("torchpack.distributed", **{k: getattr(d, k) for k in (
            "init", "size", "rank", "local_size", "local_rank", "is_master", "barrier", "allgather", "allreduce", "broadcast")})
        env = _module("torchpack.environ", get_run_dir=tp.get_run_dir, set_run_dir=tp.set_run_dir, auto_set_run_dir=tp.auto_set_run_dir)
        cb_names = ("Callback", "Callbacks", "LambdaCallback", "ProgressBar", "EstimatedTimeLeft", "InferenceRunner", "Saver",
                    "MaxSaver", "MinSaver", "SummaryWriter", "ConsoleWriter", "TFEventWriter", "JSONLWriter", "MetaInfoSaver")
        cbk = {k: getattr(tp, k) for k in cb_names}
        cbmod = _module("torchpack.callbacks.callback", Callback=tp.Callback, Callbacks=tp.Callbacks, LambdaCallback=tp.LambdaCallback)
        cbs = _module("torchpack.callbacks", callback=cbmod, **cbk)
        cbs.__path__ = []   # `from torchpack.callbacks.callback import Callback` treats it as a package
        summ = _module("torchpack.train.summary", Summary=tp.Summary)
        exc = _module("torchpack.train.exception", StopTraining=tp.StopTraining)
        train = _module("torchpack.train", Trainer=tp.Trainer, Summary=tp.Summary, StopTraining=tp.StopTraining, summary=summ,
                        exception=exc)
        train.__path__ = []
        utils.__path__ = []
        top = _module("torchpack", utils=utils, distributed=dist, environ=env, callbacks=cbs, train=train)
        top.__path__ = []
        done.append("torchpack")
    if _missing("prettytable"):
        from . import torchpack_shim as tp
        _module("prettytable", PrettyTable=tp.PrettyTable)
        done.append("prettytable")
    if _missing("nuscenes"):
        from . import torchpack_shim as tp
        u = _module("nuscenes.eval.lidarseg.utils", ConfusionMatrix=tp.ConfusionMatrix)
        ls = _module("nuscenes.eval.lidarseg", utils=u)
        ev = _module("nuscenes.eval", lidarseg=ls)
        top = _module("nuscenes", eval=ev)
        for m in (ls, ev, top):
            m.__path__ = []
        done.append("nuscenes.eval.lidarseg.utils")
    if _missing("open3d") and "visualize_utils" not in sys.modules:
        # visualize_utils.py (repo root of the reference) needs open3d at import time; the trainers import two drawing helpers
        # from it and never call them on the training path (core/spformer_trainer.py:17,67-73 — commented out)
        def _no_display(*_a, **_k):
            raise RuntimeError("visualize_utils: no display / open3d in this environment")
        _module("visualize_utils", visualize_img=_no_display, visualize_pcd=_no_display)
        done.append("visualize_utils")
    sptr = importlib.import_module("u2mkd_b200.sptr")
    # the import path the reference uses (spherical_transformer.py:7).  A real `third_party` package on sys.path (the
    # reference checkout) is left alone — only the `sptr` leaf, whose own __init__ would import the absent `sptr_cuda`
    # extension, is pre-seeded in sys.modules; without a checkout the two parent packages are stand-ins
    if "third_party" not in sys.modules and _missing("third_party"):
        _module("third_party")
    if "third_party.SparseTransformer" not in sys.modules and _missing("third_party.SparseTransformer"):
        st = _module("third_party.SparseTransformer", sptr=sptr)
        if getattr(sys.modules.get("third_party"), "__u2_shim__", False):
            sys.modules["third_party"].SparseTransformer = st
    cur = sys.modules.get("third_party.SparseTransformer.sptr")
    if cur is None or not getattr(cur, "__u2_keep__", False):   # (tests pre-register the CPU oracle's namespace and mark it)
        sys.modules["third_party.SparseTransformer.sptr"] = sptr
    sys.modules.setdefault("sptr", sptr)
    done.append("third_party.SparseTransformer.sptr")
    return done
