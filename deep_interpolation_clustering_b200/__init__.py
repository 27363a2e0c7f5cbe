"""B200-native hot path of Deep_Interpolation_Clustering (interpolation network, DEC step,
k-means / gap-statistic sweep) behind the reference's own module interfaces."""
from .interpolation_layer import CrossChannelInterp, SingleChannelInterp  # noqa: F401
from .rbf import RBF, CompressFC, TimeDistributed, basis_func_dict, gaussian  # noqa: F401
from .dec import ClusterAssignment, target_distribution  # noqa: F401
from .functional import upload_encounters  # noqa: F401
from .packed import PackedEncounters, PackedStaging  # noqa: F401

__version__ = "0.1.0"
