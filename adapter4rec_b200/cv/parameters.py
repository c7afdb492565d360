"""Command-line surface of the image tree's entry script, mirroring Downstream/CV/parameters.py:4-70 flag for flag (same
names, types and defaults: tests/golden/cv_flags.json is the reference parser's own dump), so launch lines written for the
reference's run_adapter.py work unchanged."""
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
    ("arch", str, "sasrec"), ("use_scale", str, "half"), ("n_tokens", int, 10),
    # switch and logging setting
    ("num_workers", int, 12), ("load_ckpt_name", str, "None"), ("label_screen", str, "None"), ("logging_num", int, 8),
    ("testing_num", int, 1), ("local_rank", int, -1), ("pretrained_recsys_model", str, "None"),
    # adapters
    ("adapter_down_size", int, 16), ("adding_adapter_to", str, "bert"), ("fine_tune_to", str, "None"),
    ("adapter_cv_lr", float, 5e-4), ("adapter_sasrec_lr", float, 1e-4), ("cv_adapter_down_size", int, 64),
    ("adapter_dropout_rate", float, 0.1), ("adapter_activation", str, "RELU"), ("finetune_layernorm", str, "None"),
    ("is_serial", str, "True"), ("adapter_type", str, "houslby"), ("k_adapter_bert_list", str, "0,11"),
    ("k_adapter_bert_hidden_dim", int, 384), ("num_adapter_heads_sasrec", int, 2), ("num_adapter_heads_bert", int, 12),
    ("num_dnn", int, 0),
    # compacter
    ("hypercomplex_division", int, 4), ("phm_init_range", float, 0.0001),
)


def build_parser():
    p = argparse.ArgumentParser()
    for name, typ, default in _FLAGS:
        p.add_argument("--" + name, type=typ, default=default)
    return p


def parse_args(argv=None):
    return build_parser().parse_args(argv)
