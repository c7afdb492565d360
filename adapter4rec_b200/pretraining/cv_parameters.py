"""Command-line surface of the image pre-training script, mirroring Pretraining/CV/parameters.py:4-45 flag for flag (same
names, types and defaults: tests/golden/pretrain_cv_flags.json is the reference parser's own dump)."""
import argparse

_FLAGS = (
    # data_dir
    ("mode", str, "train"), ("item_tower", str, "modal"), ("root_data_dir", str, "../"), ("dataset", str, "pinterest"),
    ("behaviors", str, "users_log.tsv"), ("images", str, "images_log.tsv"), ("lmdb_data", str, "image.lmdb"),
    # train parameters
    ("batch_size", int, 64), ("epoch", int, 1), ("lr", float, 1e-3), ("fine_tune_lr", float, 1e-5), ("l2_weight", float, 0),
    ("drop_rate", float, 0.1),
    # model parameters
    ("CV_model_load", str, "resnet-50"), ("freeze_paras_before", int, 45), ("CV_resize", int, 224), ("embedding_dim", int, 64),
    ("num_attention_heads", int, 2), ("transformer_block", int, 2), ("max_seq_len", int, 10), ("min_seq_len", int, 5),
    ("arch", str, "sasrec"),
    # switch and logging setting
    ("num_workers", int, 12), ("load_ckpt_name", str, "None"), ("label_screen", str, "None"), ("logging_num", int, 8),
    ("testing_num", int, 1), ("local_rank", int, -1),
)


def build_parser():
    p = argparse.ArgumentParser()
    for name, typ, default in _FLAGS:
        p.add_argument("--" + name, type=typ, default=default)
    return p


def parse_args(argv=None):
    return build_parser().parse_args(argv)
