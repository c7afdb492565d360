"""Entry points of Downstream/CV/run.py — full fine-tuning of the image tree on the target domain (the baseline the adapter
runs of run_adapter.py are compared with).  Same flags as run_adapter.py (adapter4rec_b200.cv.parameters mirrors the one
Downstream/CV/parameters.py both scripts share); the model is always `Model`, --pretrained_recsys_model is loaded right after
construction (Downstream/CV/run.py:59-67,156-164), the first --freeze_paras_before ViT parameters stay frozen and the rest
trains in two learning-rate groups; every epoch ranks the validation AND the test users (:277-284) and writes epoch-{n}.pt."""
from ..pretraining import cv_run as _impl
from ..pretraining.cv_run import setup_seed, synthetic_data  # noqa: F401


def build_model(args, item_num, local_rank, vit_config=None, vit_state_dict=None):
    return _impl.build_model(args, item_num, local_rank, vit_config, vit_state_dict, downstream=True)


def train(args, use_modal, local_rank, data, Log_file=None, vit_config=None, users_per_pass=32, model_dir=None):
    return _impl.train(args, use_modal, local_rank, data, Log_file, vit_config, users_per_pass, model_dir, downstream=True)


def test(args, use_modal, local_rank, data, Log_file=None, vit_config=None, model_dir=None):
    return _impl.test(args, use_modal, local_rank, data, Log_file, vit_config, model_dir, downstream=True)
