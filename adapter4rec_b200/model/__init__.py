from .bert import BertModel, RobertaModel, TextConfigLite
from .encoders import Bert_Encoder, Text_Encoder, User_Encoder
from .layers import LayerNorm, Linear, LoRALinear
from .model import BertAdaptedSelfOutput, Model, ModelCPC, SASRecAdaptedSelfOutput, SoftEmbedding
from .modules import AdapterBlock, MultiHeadedAttention, PositionwiseFeedForward, TransformerBlock, TransformerEncoder
