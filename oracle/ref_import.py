"""Run the reference's OWN source files in this container (not on the GPU box).

TEST INFRASTRUCTURE ONLY.  ``/root/reference`` is pure Python but imports
``torch_geometric`` / ``torch_scatter`` unconditionally (tgp/select/base_select.py:8,
tgp/utils/ops.py:19), and neither is installed here.  This module installs stub
packages under those names whose *only* real members are the restated primitives
of ``oracle/pyg_shim.py``; every other attribute resolves to an inert placeholder
class so that class bodies and type annotations in the reference still evaluate.
The reference's control flow (``BaseReduce.forward``, ``sparse_connect``,
``DenseConnect``, ``postprocess_adj_pool_*``, the four losses, ``SelectOutput``,
``TopkSelect``) then executes verbatim on top of those primitives.

Used by ``tests/golden/make_golden.py`` and ``tests/golden/make_golden_unbatched.py`` (to generate the committed
fixtures; the latter also asserts that the oracle reproduces the reference bit for bit).
Never imported by the product package, ``smoke()`` or ``bench.py``.
"""

from __future__ import annotations

import importlib
import importlib.abc
import importlib.machinery
import math
import os
import sys
import types

import torch

from . import pyg_shim

REFERENCE_ROOT = os.environ.get("TGP_REFERENCE_ROOT", "/root/reference")
_STUB_ROOTS = ("torch_geometric", "torch_scatter")


class _PlaceholderMeta(type):
    def __getattr__(cls, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Placeholder

    def __instancecheck__(cls, inst):  # isinstance(x, SparseTensor) etc. -> False
        return type.__instancecheck__(cls, inst) if cls is not _Placeholder else False


class _Placeholder(metaclass=_PlaceholderMeta):
    """Inert stand-in for any PyG symbol the hot path never executes."""

    def __init__(self, *args, **kwargs):
        pass

    def __call__(self, *args, **kwargs):
        return self

    def __class_getitem__(cls, item):
        return cls


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Placeholder


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in _STUB_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        mod = _StubModule(spec.name)
        mod.__path__ = []
        return mod

    def exec_module(self, module):
        _populate(module)


class _Act(torch.nn.Module):
    def __init__(self, fn):
        super().__init__()
        self.fn = fn

    def forward(self, x):
        return self.fn(x)


def _activation_resolver(query="relu", *args, **kwargs):
    if callable(query) and not isinstance(query, str):
        return query
    table = {
        "tanh": torch.tanh,
        "relu": torch.relu,
        "sigmoid": torch.sigmoid,
        "linear": lambda v: v,
        "identity": lambda v: v,
        "softmax": lambda v: torch.softmax(v, -1),
    }
    key = str(query).lower()
    if key not in table:
        raise ValueError(f"Could not resolve '{query}'")
    return _Act(table[key])


def _uniform(size, value):
    if isinstance(value, torch.Tensor):
        bound = 1.0 / math.sqrt(size)
        value.data.uniform_(-bound, bound)


def _populate(module: types.ModuleType) -> None:
    name = module.__name__
    if name == "torch_geometric.utils":
        for fn in (
            "scatter",
            "coalesce",
            "subgraph",
            "remove_self_loops",
            "unbatch",
            "unbatch_edge_index",
            "to_dense_adj",
            "to_dense_batch",
            "softmax",
            "cumsum",
            "degree",
            "index_sort",
        ):
            setattr(module, fn, getattr(pyg_shim, fn))
        module.is_sparse = lambda t: isinstance(t, torch.Tensor) and t.is_sparse
        module.is_torch_sparse_tensor = lambda t: isinstance(t, torch.Tensor) and t.layout != torch.strided
    elif name == "torch_geometric.utils.num_nodes":
        module.maybe_num_nodes = pyg_shim.maybe_num_nodes
    elif name == "torch_geometric.nn.pool.select.topk":
        module.topk = pyg_shim.topk
    elif name == "torch_geometric.nn.resolver":
        module.activation_resolver = _activation_resolver
    elif name == "torch_geometric.nn.inits":
        module.uniform = _uniform
        module.zeros = lambda v: v.data.fill_(0) if isinstance(v, torch.Tensor) else None
    elif name == "torch_geometric.typing":
        module.Adj = torch.Tensor
        module.OptTensor = torch.Tensor
        module.Tensor = torch.Tensor
        module.PairTensor = tuple
        module.WITH_TORCH_SCATTER = False
    elif name == "torch_scatter":
        module.scatter = pyg_shim.torch_scatter_scatter
        module.scatter_add = lambda src, index, dim=0, dim_size=None: pyg_shim.scatter(src, index, dim, dim_size, "sum")
        module.scatter_mul = lambda src, index, dim=0, dim_size=None: pyg_shim.scatter(src, index, dim, dim_size, "mul")


_installed = False


def install_stubs() -> None:
    global _installed
    if _installed:
        return
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "tgp")):
        raise FileNotFoundError(f"reference tree not found at {REFERENCE_ROOT}")
    sys.meta_path.insert(0, _StubFinder())
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed = True


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "tgp"))


def load_reference():
    """Return a namespace with the reference's real hot-path objects."""
    install_stubs()
    ns = types.SimpleNamespace()
    ns.ops = importlib.import_module("tgp.utils.ops")
    ns.losses = importlib.import_module("tgp.utils.losses")
    ns.base_select = importlib.import_module("tgp.select.base_select")
    ns.SelectOutput = ns.base_select.SelectOutput
    ns.topk_select = importlib.import_module("tgp.select.topk_select")
    ns.base_reduce = importlib.import_module("tgp.reduce.base_reduce")
    ns.base_conn = importlib.import_module("tgp.connect.base_conn")
    ns.dense_conn = importlib.import_module("tgp.connect.dense_conn")
    return ns
