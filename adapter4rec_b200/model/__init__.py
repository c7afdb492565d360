"""The names `from model import ...` gives in the reference (Downstream/Text/model/__init__.py: model, modules, layers,
encoders are star-exported there) plus the BERT / RoBERTa bodies the reference takes from transformers."""
from .bert import BertModel, PackedTokens, RobertaModel, TextConfigLite
from .encoders import Bert_Encoder, Text_Encoder, User_Encoder
from .layers import LayerNorm, Linear, LoRALinear, PHMLinear
from .model import (BertAdaptedParallelSelfOutput, BertAdaptedSelfOutput, BertCompacterAdaptedSelfOutput,
                    BertKAdaptedBertModel, BertPfeifferAdaptedSelfOutput, CompacterModel, Model, ModelCPC,
                    SASRecAdaptedSelfOutput, SASRecCompacterAdaptedSelfOutput, SASRecKAdaptedTransformerBlocks,
                    SASRecParallelAdaptedSelfOutput, SASRecPfeifferAdaptedSelfOutput, SASRecPfeifferVer2AdaptedSelfOutput,
                    SoftEmbedding)
from .modules import (AdapterBlock, AdapterPfeifferBlock, HyperComplexAdapterBlock, KAdapterBlock, MultiHeadedAttention,
                      PositionwiseFeedForward, SelfAttention, TransformerBlock, TransformerEncoder)
