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


def patch_trainer(module, class_name="TrainerCluster"):
    """Route the epoch-level feature / label path of the clustering trainer through ``epoch_path``
    without editing the reference: call it on the imported ``clustering_trainer`` module.

      * ``KMeans`` (clustering_trainer.py:19,75-82) -> ``KMeansB200`` (the centre initialisation and the
        validation ``predict``);
      * ``eval_one_epoch`` (:285-422) is wrapped, not replaced: a forward hook on ``self.model`` copies every
        batch's latent ``hidden`` and soft assignment ``cluster_pred`` into an ``EpochFeatures`` buffer in HBM
        while the reference's own loop runs;
      * ``generate_pred_cluster`` (:473-484) then takes argmax_j q and the label delta on the device from that
        buffer (one scalar + the int labels come back) and no longer calls ``merge_ob_pred`` - the
        ``list.extend`` over every row of every tensor of the epoch (:486-493) is what dominates the reference's
        epoch at 10^6 encounters;
      * ``merge_ob_pred`` itself (still used by ``generate_pretrain_feat``, :466-471) concatenates the per-batch
        arrays instead of extending Python lists row by row: same dict of ndarrays.

    Returns the patched class.  ``trainer.dic_epoch_features`` holds the device buffer of the last eval epoch."""
    import numpy as np
    import torch
    from . import epoch_path
    patch_kmeans(module)
    cls = getattr(module, class_name)
    if getattr(cls, "_dic_patched", False):
        return cls
    orig_eval = cls.eval_one_epoch

    def eval_one_epoch(self, scope, dl, denoise=False):
        store = {"feat": None}

        def hook(_mod, _inp, out):
            if not (isinstance(out, tuple) and len(out) >= 3 and isinstance(out[2], dict)):
                return
            hidden, q = out[0], out[2].get("cluster_pred")
            if store["feat"] is None:
                try:
                    capacity = len(dl.dataset)
                except (TypeError, AttributeError):
                    capacity = None
                store["feat"] = epoch_path.EpochFeatures(capacity, hidden.shape[1], hidden.device,
                                                         None if q is None else q.shape[1])
            store["feat"].append(hidden, q)

        handle = self.model.register_forward_hook(hook)
        try:
            out = orig_eval(self, scope, dl, denoise)
        finally:
            handle.remove()
        self.dic_epoch_features = store["feat"]
        return out

    def generate_pred_cluster(self, scope, dl, prev_pred, denoise=False):
        metrics_dict, _ob_pred_lst = self.eval_one_epoch(scope, dl, denoise=denoise)
        feats = self.dic_epoch_features
        cluster_pred = feats.cluster_pred()                                  # :476 on the device
        delta_label = epoch_path.label_delta(cluster_pred, prev_pred)        # :477-483, one scalar to the host
        return delta_label, cluster_pred.cpu().numpy().astype(np.int64), metrics_dict

    def merge_ob_pred(self, ob_pred_lst):
        keys = []
        for d in ob_pred_lst:
            keys += [k for k in d if k not in keys]
        merged = {}
        for k in keys:
            parts = [np.asarray(d[k]) for d in ob_pred_lst if k in d]
            merged[k] = np.concatenate(parts, axis=0) if parts[0].ndim else np.array(parts)
        return merged

    cls.eval_one_epoch = eval_one_epoch
    cls.generate_pred_cluster = generate_pred_cluster
    cls.merge_ob_pred = merge_ob_pred
    cls._dic_patched = True
    cls._dic_orig = {"eval_one_epoch": orig_eval}
    return cls


def patch_lstm(module):
    """Rebind ``EncoderRNN`` / ``DecoderRNN`` of an imported ``pretrain_interp`` / ``clustering_interp`` module
    (pretrain_interp.py:14-41) to the B200 mirrors (lstm.py): ``Net.__init__`` looks the names up in its module at
    call time, so Nets built afterwards run their BiLSTMs on the persistent tcgen05 kernel; the ``encoder.lstm.*`` /
    ``decoder.lstm.*`` state-dict keys are unchanged."""
    from . import lstm
    for name in ("EncoderRNN", "DecoderRNN"):
        if hasattr(module, name):
            setattr(module, name, getattr(lstm, name))
    return module
