"""Make the reference's own scripts pick up the B200 operators, unchanged.

The reference imports its hot-path operators by bare module name
(``from interpolation_layer import ...`` pretrain_interp.py:11, clustering_interp.py:9;
``from rbf import RBF, basis_func_dict`` :12/:10; ``from dec import ...``
clustering_interp.py:11).  ``install()`` registers this package's mirrors under those names
in ``sys.modules`` *before* the reference's modules are imported, so ``pretrain_interp.py``,
``clustering_interp.py`` and the trainers run without a single edit.
"""
from __future__ import annotations

import sys

_NAMES = ("interpolation_layer", "rbf", "dec")
_saved = {}


def install(kmeans=True):
    """Shadow ``interpolation_layer``, ``rbf`` and ``dec`` with the B200 mirrors.

    With ``kmeans=True`` also rebinds ``sklearn.cluster.KMeans`` lookups made by modules
    imported afterwards via ``patch_kmeans(module)`` (call it on clustering_trainer /
    p2_clustering_optK after importing them).
    """
    from . import dec, interpolation_layer, rbf
    for name, mod in zip(_NAMES, (interpolation_layer, rbf, dec)):
        if name in sys.modules and sys.modules[name] is not mod:
            _saved[name] = sys.modules[name]
        sys.modules[name] = mod


def uninstall():
    for name in _NAMES:
        if name in _saved:
            sys.modules[name] = _saved.pop(name)
        else:
            sys.modules.pop(name, None)


def patch_kmeans(module):
    """Rebind the ``KMeans`` global of an already imported reference module
    (clustering_trainer.py:19, p2_clustering_optK.py:14, p4_clustering_final.py) to KMeansB200."""
    from .kmeans import KMeansB200
    if hasattr(module, "KMeans"):
        module.KMeans = KMeansB200
    return module
