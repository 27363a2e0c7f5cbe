"""Drop-in mirror of the reference's ``dec.py`` (DEC soft assignment + target distribution)."""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn
from torch.nn import Parameter

from . import functional as F_


class ClusterAssignment(nn.Module):
    """Student-t soft assignment of latents to cluster centres.  dec.py:13-63.

    Same constructor, ``cluster_centers`` parameter (K, D), ``init_center`` / ``get_center``
    and ``forward(batch (B, D)) -> q (B, K)`` as the reference.
    """

    def __init__(self, cluster_number: int, embedding_dimension: int, alpha: float = 1.0,
                 cluster_centers: Optional[torch.Tensor] = None) -> None:
        super().__init__()
        self.embedding_dimension = embedding_dimension
        self.cluster_number = cluster_number
        self.alpha = alpha
        if cluster_centers is None:
            initial_cluster_centers = torch.zeros(self.cluster_number, self.embedding_dimension,
                                                  dtype=torch.float)
            nn.init.xavier_uniform_(initial_cluster_centers)                    # dec.py:32-38
        else:
            initial_cluster_centers = cluster_centers
        self.cluster_centers = Parameter(initial_cluster_centers)

    def init_center(self, initial_cluster_centers):
        self.cluster_centers.data = initial_cluster_centers                      # dec.py:44

    def get_center(self):
        return self.cluster_centers

    def forward(self, batch: torch.Tensor) -> torch.Tensor:
        return F_.dec_soft_assign(batch, self.cluster_centers, self.alpha)


def target_distribution(batch: torch.Tensor) -> torch.Tensor:
    """p_ij = (q_ij^2 / f_j) / sum_j' (q_ij'^2 / f_j'), f_j = sum_i q_ij.  dec.py:66-76.

    Evaluated without an autograd graph: every caller of the reference detaches the result
    (clustering_interp.py:186).  For a batch sharded across ranks use
    ``parallel.sharded_target_distribution`` which all-reduces f_j first.
    """
    return F_.dec_target_distribution(batch)
