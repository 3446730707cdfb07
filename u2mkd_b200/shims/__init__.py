"""Host shims (the part the model files need): minimal stand-ins for third-party Python packages that
core/models/nuscenes/spvcnn_spformer.py and core/models/sphereformer/*.py import but this path never depended on for
arithmetic — timm.models.layers (DropPath, trunc_normal_), torch_scatter (scatter_mean / scatter_add / scatter_max),
torchpack.utils.config (the global `configs` mapping) — and the import path third_party.SparseTransformer.sptr.

    import u2mkd_b200
    u2mkd_b200.install_as_torchsparse(); u2mkd_b200.install_reference_shims()
    from core.models.nuscenes.spvcnn_spformer import SPVCNN_SPFORMER      # reference file, unchanged

A shim is only registered when the real package is NOT importable.  Trainer-level torchpack (Trainer, callbacks,
distributed launch) is out of scope."""
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
        for k, v in dict(other, **kwargs).items():
            if isinstance(v, dict) and isinstance(self.get(k), Config):
                self[k].update(v)
            else:
                self[k] = v

    def load(self, fpath, recursive=False):
        import yaml
        with open(fpath) as f:
            self.update(yaml.safe_load(f) or {})


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
        cfg = _module("torchpack.utils.config", Config=Config, configs=configs)
        utils = _module("torchpack.utils", config=cfg)
        _module("torchpack", utils=utils)
        done.append("torchpack.utils.config")
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
    sys.modules["third_party.SparseTransformer.sptr"] = sptr
    sys.modules.setdefault("sptr", sptr)
    done.append("third_party.SparseTransformer.sptr")
    return done
