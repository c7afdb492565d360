from .model import Model, ModelCPC, SASRecAdaptedSelfOutput, SoftPrompt, VITAdaptedOutput, VITAdaptedSelfOutput, Vit_Encoder
from .vit import ViTConfigLite, ViTForImageClassification, ViTModel
