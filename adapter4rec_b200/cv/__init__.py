"""The image tree (Downstream/CV/model): names of model.py / encoders.py plus the ViT body taken from transformers there."""
from .model import (CompacterModel, Model, ModelCPC, SASRecAdaptedSelfOutput, SASRecCompacterAdaptedSelfOutput,
                    SASRecParallelAdaptedSelfOutput, SASRecPfeifferV2AdaptedSelfOutput, SoftPrompt, VITAdaptedOutput,
                    VITAdaptedParallelOutput, VITAdaptedParallelSelfOutput, VITAdaptedSelfOutput, VITCompacterAdaptedOutput,
                    VITCompacterAdaptedSelfOutput, Vit_Encoder)
from .vit import ViTConfigLite, ViTForImageClassification, ViTModel
