"""Install the third-party restatements under their real module names so the UNMODIFIED
first-party reference at /root/reference can be imported and run on CPU in this container.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Used only by oracle/make_golden.py and by
tests that are skipped when /root/reference is absent (the GPU box).  Heavy, irrelevant
dependencies of the reference's import graph (rdkit, Bio, prody, esm, spyrmsd's optional
backends, wandb ...) become inert MagicMock modules: none of them is touched on the hot path.
"""
from __future__ import annotations

import importlib
import os
import sys
import types
from unittest.mock import MagicMock

REFERENCE_ROOT = os.environ.get("CB_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "models"))


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    parent, _, child = name.rpartition(".")
    if parent and parent in sys.modules:
        setattr(sys.modules[parent], child, m)
    return m


_MOCK_ROOTS = set()


class _MockFinder:
    """Any import below a mocked root (rdkit.*, Bio.*, ...) resolves to an inert MagicMock module."""

    @staticmethod
    def find_spec(name, path=None, target=None):
        import importlib.machinery
        if name.split(".")[0] in _MOCK_ROOTS:
            return importlib.machinery.ModuleSpec(name, _MockFinder, is_package=True)
        return None

    @staticmethod
    def create_module(spec):
        mm = MagicMock(name=spec.name)
        mm.__path__ = []
        mm.__name__ = spec.name
        mm.__spec__ = spec
        mm.__loader__ = _MockFinder
        return mm

    @staticmethod
    def exec_module(module):
        pass


def _mock(*roots):
    _MOCK_ROOTS.update(roots)
    if _MockFinder not in sys.meta_path:
        sys.meta_path.insert(0, _MockFinder)


_installed = False


def install():
    """Idempotent.  After this, `import models.score_model`, `import utils.sampling` ... work."""
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    import torch
    from torch import nn

    from . import cluster, o3, scatter
    from confidence_bootstrapping_b200 import data as pyg

    # e3nn ------------------------------------------------------------------------------
    class _Unsupported(nn.Module):
        def __init__(self, *a, **k):
            super().__init__()
            raise NotImplementedError("oracle shim: op unused by the shipped configurations")

    e3 = _module("e3nn")
    _module("e3nn.o3", Irreps=o3.Irreps, Irrep=o3.Irrep, spherical_harmonics=o3.spherical_harmonics,
            FullyConnectedTensorProduct=o3.FullyConnectedTensorProduct, FullTensorProduct=o3.FullTensorProduct,
            wigner_3j=o3.wigner_3j, TensorProduct=_Unsupported, Linear=_Unsupported)
    _module("e3nn.nn", BatchNorm=o3.BatchNorm)
    e3.__path__ = []

    # torch_cluster / torch_scatter -------------------------------------------------------
    def _knn_graph(*a, **k):
        raise NotImplementedError("knn_graph is preprocessing-only (out of scope)")

    _module("torch_cluster", radius=cluster.radius, radius_graph=cluster.radius_graph, knn_graph=_knn_graph)
    _module("torch_scatter", scatter=scatter.scatter, scatter_mean=scatter.scatter_mean,
            scatter_add=lambda src, index, dim=0, out=None, dim_size=None: scatter.scatter(
                src, index, dim, out, dim_size, "sum"))

    # torch_geometric ---------------------------------------------------------------------
    class _DataParallel(nn.Module):
        def __init__(self, module, *a, **k):
            super().__init__()
            self.module = module

        def forward(self, x):
            return self.module(x)

    class _Dataset:
        def __init__(self, root=None, transform=None, *a, **k):
            self.transform = transform

    tg = _module("torch_geometric")
    tg.__path__ = []
    _module("torch_geometric.data", Batch=pyg.Batch, HeteroData=pyg.HeteroData, Data=pyg.HeteroData,
            Dataset=_Dataset)
    _module("torch_geometric.data.dataset", Dataset=_Dataset)
    _module("torch_geometric.loader", DataLoader=pyg.DataLoader, DataListLoader=pyg.DataLoader)
    _module("torch_geometric.loader.dataloader", DataLoader=pyg.DataLoader, Collater=object)
    _module("torch_geometric.nn")
    sys.modules["torch_geometric.nn"].__path__ = []
    _module("torch_geometric.nn.data_parallel", DataParallel=_DataParallel)
    _module("torch_geometric.utils", subgraph=pyg.subgraph, degree=None, to_networkx=None, dense_to_sparse=None,
            to_dense_adj=None)
    _module("torch_geometric.transforms", BaseTransform=object)

    # everything irrelevant to the hot path -----------------------------------------------
    _mock("rdkit", "esm", "Bio", "prody", "wandb", "biopandas", "plotly", "openbabel", "graph_tool",
          "qcelemental", "lmdb", "tqdm_placeholder")

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # the reference's top-level dirs are namespace packages (no __init__.py) and lose against
    # same-named site-packages (HF `datasets`): bind them explicitly.
    for pkg in ("datasets", "utils", "models", "confidence", "bootstrapping", "spyrmsd"):
        for k in [k for k in sys.modules if k == pkg or k.startswith(pkg + ".")]:
            del sys.modules[k]
        m = types.ModuleType(pkg)
        m.__path__ = [os.path.join(REFERENCE_ROOT, pkg)]
        sys.modules[pkg] = m
    _installed = True


def import_reference(name):
    install()
    return importlib.import_module(name)
