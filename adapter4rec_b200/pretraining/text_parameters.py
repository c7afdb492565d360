"""Command-line surface of the text pre-training script, mirroring Pretraining/Text/parameters.py:4-51 flag for flag (same
names, types, choices and defaults: tests/golden/pretrain_text_flags.json is the reference parser's own dump)."""
import argparse


def build_parser():
    p = argparse.ArgumentParser()
    # data_dir
    p.add_argument("--mode", type=str, default="train", choices=['train', 'test', 'load'])
    p.add_argument("--item_tower", type=str, default="modal", choices=['modal', 'id'])
    p.add_argument("--root_data_dir", type=str, default="../")
    p.add_argument("--dataset", type=str, default='MIND-small')
    p.add_argument("--behaviors", type=str, default='behaviors_l5_tr_v.tsv')
    p.add_argument("--news", type=str, default='news_l5_tr_v.tsv')
    # train parameters
    for name, typ, default in (("batch_size", int, 64), ("epoch", int, 1), ("lr", float, 1e-5), ("fine_tune_lr", float, 1e-5),
                               ("l2_weight", float, 0), ("drop_rate", float, 0.1)):
        p.add_argument("--" + name, type=typ, default=default)
    # model parameters
    p.add_argument("--bert_model_load", type=str, default='bert-base-uncased')
    for name, default in (("freeze_paras_before", 165), ("word_embedding_dim", 768), ("embedding_dim", 256),
                          ("num_attention_heads", 2), ("transformer_block", 2), ("max_seq_len", 20), ("min_seq_len", 5)):
        p.add_argument("--" + name, type=int, default=default)
    p.add_argument("--arch", type=str, default="sasrec")
    # switch and logging setting
    p.add_argument("--num_workers", type=int, default=12)
    p.add_argument("--load_ckpt_name", type=str, default='None')
    p.add_argument("--label_screen", type=str, default='None')
    p.add_argument("--logging_num", type=int, default=8)
    p.add_argument("--testing_num", type=int, default=1)
    p.add_argument("--local_rank", default=-1, type=int)
    # news information
    p.add_argument("--num_words_title", type=int, default=30)
    p.add_argument("--num_words_abstract", type=int, default=50)
    p.add_argument("--num_words_body", type=int, default=50)
    p.add_argument("--news_attributes", type=str, default='title')
    return p


def parse_args(argv=None):
    args = build_parser().parse_args(argv)
    args.news_attributes = args.news_attributes.split(',')
    return args
