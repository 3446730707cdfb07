"""SparseTensor / PointTensor containers [TS v1.4.0 torchsparse/tensor.py]; used at
core/models/utils.py:28-33,59-61,100-108 and core/models/semantickitti/spvcnn.py:94."""
from .utils import make_ntuple

__all__ = ["SparseTensor", "PointTensor"]


class SparseTensor:
    """feats [N,C] + coords int32 [N,4] (x,y,z,batch) + stride 3-tuple; `cmaps` / `kmaps` are
    dicts shared BY REFERENCE between all tensors derived from one input."""

    def __init__(self, feats, coords, stride=1):
        self.feats = feats
        self.coords = coords
        self.stride = make_ntuple(stride, ndim=3)
        self.cmaps = {}
        self.kmaps = {}

    @property
    def F(self):
        return self.feats

    @F.setter
    def F(self, feats):
        self.feats = feats

    @property
    def C(self):
        return self.coords

    @C.setter
    def C(self, coords):
        self.coords = coords

    @property
    def s(self):
        return self.stride

    @s.setter
    def s(self, stride):
        self.stride = stride

    def _map(self, fn):
        self.coords = fn(self.coords)
        self.feats = fn(self.feats)
        return self

    def cpu(self):
        return self._map(lambda t: t.cpu())

    def cuda(self):
        return self._map(lambda t: t.cuda())

    def detach(self):
        return self._map(lambda t: t.detach())

    def to(self, device, non_blocking: bool = True):
        return self._map(lambda t: t.to(device, non_blocking=non_blocking))

    def __add__(self, other):
        out = SparseTensor(coords=self.coords, feats=self.feats + other.feats, stride=self.stride)
        out.cmaps = self.cmaps
        out.kmaps = self.kmaps
        return out


class PointTensor:
    def __init__(self, feats, coords, idx_query=None, weights=None):
        self.F = feats
        self.C = coords
        self.idx_query = {} if idx_query is None else idx_query
        self.weights = {} if weights is None else weights
        self.additional_features = {"idx_query": {}, "counts": {}}

    def _map(self, fn):
        self.F = fn(self.F)
        self.C = fn(self.C)
        return self

    def cuda(self):
        return self._map(lambda t: t.cuda())

    def detach(self):
        return self._map(lambda t: t.detach())

    def to(self, device, non_blocking: bool = True):
        return self._map(lambda t: t.to(device, non_blocking=non_blocking))

    def __add__(self, other):
        out = PointTensor(self.F + other.F, self.C, self.idx_query, self.weights)
        out.additional_features = self.additional_features
        return out
