from .gnn import GNN, DeepSetEncoder, ResGnn          # noqa: F401
from .loss import MixedLoss, MixedNormalCRPS, NormalCRPS  # noqa: F401
from .model_utils import PostProcess                   # noqa: F401
