"""The image tree (Downstream/CV/model): names of model.py / encoders.py plus the ViT body taken from transformers there."""
from .model import (CompacterModel, Model, ModelCPC, SASRecAdaptedSelfOutput, SASRecCompacterAdaptedSelfOutput,
                    SASRecParallelAdaptedSelfOutput, SASRecPfeifferV2AdaptedSelfOutput, SoftPrompt, VITAdaptedOutput,
                    VITAdaptedParallelOutput, VITAdaptedParallelSelfOutput, VITAdaptedSelfOutput, VITCompacterAdaptedOutput,
                    VITCompacterAdaptedSelfOutput, VITKAdaptedCVModel, VITPfeifferAdaptedSelfOutput, Vit_Encoder)
from ..model import PHMLinear, SASRecKAdaptedTransformerBlocks  # noqa: F401  (names of run_adapter.py's import line)
from .vit import ViTConfigLite, ViTForImageClassification, ViTModel
