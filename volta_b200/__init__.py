"""volta_b200: a B200-native (sm_100a) implementation of the VOLTA controlled-study encoder forward and
backward, behind the reference's module API.  See DESIGN.md."""
from .config import BertConfig, named_config, NAMED_CONFIGS  # noqa: F401
from .encoders import (BertModel, BertForVLPreTraining, BertForVLTasks, BertEncoder, BertGatedAttention,  # noqa: F401
                       BertGatedFeedForward, BertPreTrainingHeads, SimpleClassifier, BertLayerNorm)

__version__ = "0.1.0"
