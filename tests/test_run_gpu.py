"""-m gpu: the entry-script mirror (adapter4rec_b200.run.train / run_eval) trains a small Houlsby TransRec on synthetic
arrays through the same flags as Downstream/Text/run.py and the loss goes down."""
import logging

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_train_entry_reduces_loss_and_evaluates(tmp_path):
    from adapter4rec_b200 import run
    from adapter4rec_b200.model import TextConfigLite
    from adapter4rec_b200.parameters import parse_args
    args = parse_args(["--embedding_dim", "64", "--batch_size", "32", "--epoch", "3", "--adapter_type", "houslby",
                       "--adding_adapter_to", "all", "--bert_model_load", "bert_tiny", "--word_embedding_dim", "128",
                       "--bert_adapter_down_size", "16", "--adapter_bert_lr", "5e-3", "--adapter_sasrec_lr", "5e-3",
                       "--max_seq_len", "10", "--num_words_title", "12", "--drop_rate", "0.0",
                       "--pretrained_model_name", "None"])   # the reference default 'epoch-15' loads a checkpoint (run.py:374-381)
    run.setup_seed(123456)
    data = run.synthetic_data(item_num=300, users=96, num_words=12, max_seq_len=10, vocab=500)
    cfg = TextConfigLite(vocab_size=500, hidden_size=128, num_hidden_layers=2, num_attention_heads=2,
                         intermediate_size=512, max_position_embeddings=32, hidden_dropout_prob=0.0,
                         attention_probs_dropout_prob=0.0)
    log = logging.getLogger("run_test")
    records = []
    log.addHandler(type("H", (logging.Handler,), {"emit": lambda self, r: records.append(r.getMessage())})())
    log.setLevel(logging.INFO)
    model, trainer, hit10 = run.train(args, True, 0, data, Log_file=log, bert_config=cfg, users_per_pass=16,
                                      model_dir=str(tmp_path))
    losses = [float(m.split(":")[-1]) for m in records if "mean batch loss" in m]
    assert len(losses) == 3 and losses[-1] < losses[0], losses
    assert 0.0 <= hit10 <= 1.0 and any("valid_results" in m for m in records)
    assert trainer.step_count == 9                                        # 96 users / 32 per batch x 3 epochs
    # checkpoints in the reference's format (data_utils/utils.py:109-115) and the resume path (--load_ckpt_name, run.py:481-493)
    import os
    saved = sorted(f for f in os.listdir(tmp_path) if f.startswith("epoch-"))
    assert saved, "an improving epoch must have written epoch-N.pt"
    ckpt = torch.load(os.path.join(tmp_path, saved[-1]), weights_only=False)
    assert set(ckpt) >= {"model_state_dict", "optimizer", "rng_state", "cuda_rng_state"}
    assert set(ckpt["model_state_dict"]) == set(model.state_dict())
    args.load_ckpt_name, args.epoch = saved[-1], 1
    model2, trainer2, _ = run.train(args, True, 0, data, Log_file=log, bert_config=cfg, users_per_pass=16, model_dir=str(tmp_path))
    assert trainer2.step_count == ckpt["optimizer"]["step"] + 3
    assert any("epoch %d mean batch loss" % (int(saved[-1].split("-")[1].split(".")[0]) + 1) in m for m in records)


def test_cv_train_entry_reduces_loss_and_evaluates(tmp_path):
    """adapter4rec_b200.cv.run_adapter.train (Downstream/CV/run_adapter.py:283-620) on synthetic images through the reference's
    flags: ViT (2 layers, 48 x 48 images = 10 tokens) + serial Houlsby adapters + SASRec; loss goes down, both evaluations
    run, a checkpoint in the reference's format is written."""
    import os
    from adapter4rec_b200.cv import ViTConfigLite
    from adapter4rec_b200.cv import run_adapter as run
    from adapter4rec_b200.cv.parameters import parse_args
    args = parse_args(["--CV_model_load", "vit-base-patch16-224", "--CV_resize", "48", "--batch_size", "16", "--epoch", "3",
                       "--adapter_type", "houslby", "--adding_adapter_to", "all", "--cv_adapter_down_size", "16",
                       "--adapter_cv_lr", "5e-3", "--adapter_sasrec_lr", "5e-3", "--max_seq_len", "6", "--drop_rate", "0.0"])
    run.setup_seed(12345)
    data = run.synthetic_data(item_num=60, users=48, resize=48, max_seq_len=6)
    cfg = ViTConfigLite(hidden_size=768, num_hidden_layers=2, num_attention_heads=12, intermediate_size=256, image_size=48,
                        patch_size=16)
    log = logging.getLogger("run_cv_test")
    records = []
    log.addHandler(type("H", (logging.Handler,), {"emit": lambda self, r: records.append(r.getMessage())})())
    log.setLevel(logging.INFO)
    model, trainer, hit10 = run.train(args, True, 0, data, Log_file=log, vit_config=cfg, users_per_pass=8, model_dir=str(tmp_path))
    losses = [float(m.split(":")[-1]) for m in records if "mean batch loss" in m]
    assert len(losses) == 3 and losses[-1] < losses[0], losses
    assert 0.0 <= hit10 <= 1.0
    assert sum("valid_results" in m for m in records) == 3 and sum("test_results" in m for m in records) == 3
    assert trainer.step_count == 9
    names = {n for n, _, _ in trainer.names}
    assert names and all("adapter" in n for n in names)              # fine_tune_to None: only the adapters train
    saved = sorted(f for f in os.listdir(tmp_path) if f.startswith("epoch-"))
    assert saved
    ckpt = torch.load(os.path.join(tmp_path, saved[-1]), weights_only=False)
    assert set(ckpt["model_state_dict"]) == set(model.state_dict())
