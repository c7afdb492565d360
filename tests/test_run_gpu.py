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
    assert "epoch-1.pt" in saved and "epoch-3.pt" in saved      # first evaluation (run.py:618) and the final state (:637-638)
    assert any("test_results" in m for m in records)             # the test users are ranked when the validation HR@10 improves
    assert any(m.startswith("cnt: ") for m in records)           # --logging_num progress lines (run.py:606-608)
    ckpt = torch.load(os.path.join(tmp_path, saved[-1]), weights_only=False)
    assert set(ckpt) >= {"model_state_dict", "optimizer", "rng_state", "cuda_rng_state"}
    assert set(ckpt["model_state_dict"]) == set(model.state_dict())
    args.load_ckpt_name, args.epoch = saved[-1], 1
    model2, trainer2, _ = run.train(args, True, 0, data, Log_file=log, bert_config=cfg, users_per_pass=16, model_dir=str(tmp_path))
    assert {"state", "param_groups"} <= set(ckpt["optimizer"])         # torch.optim.Adam's own format (utils.py:109-115)
    assert trainer2.step_count == int(ckpt["optimizer"]["state"][0]["step"]) + 3
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
    assert "epoch-1.pt" in saved and "epoch-3.pt" in saved               # first evaluation and the final state (:629-630)
    ckpt = torch.load(os.path.join(tmp_path, saved[-1]), weights_only=False)
    assert set(ckpt["model_state_dict"]) == set(model.state_dict())


def _capture(name):
    log = logging.getLogger(name)
    records = []
    log.addHandler(type("H", (logging.Handler,), {"emit": lambda self, r: records.append(r.getMessage())})())
    log.setLevel(logging.INFO)
    return log, records


def test_text_pretraining_entry_trains_the_unfrozen_tail_and_resumes(tmp_path):
    """adapter4rec_b200.pretraining.text_run.train (Pretraining/Text/run.py:127-352): no adapters, embeddings + encoder layer 0
    frozen by --freeze_paras_before, layer 1 + projection + user encoder trained in two learning-rate groups; loss goes down,
    a checkpoint is written EVERY epoch and --load_ckpt_name continues from it."""
    import os
    from adapter4rec_b200.model import TextConfigLite
    from adapter4rec_b200.pretraining import text_run as run
    from adapter4rec_b200.pretraining.text_parameters import parse_args
    args = parse_args(["--embedding_dim", "64", "--batch_size", "32", "--epoch", "3", "--bert_model_load", "bert_tiny",
                       "--freeze_paras_before", "21", "--lr", "2e-3", "--fine_tune_lr", "5e-4", "--max_seq_len", "10",
                       "--num_words_title", "12", "--drop_rate", "0.0"])
    run.setup_seed(123456)
    data = run.synthetic_data(item_num=300, users=96, num_words=12, max_seq_len=10, vocab=500)
    cfg = TextConfigLite(vocab_size=500, hidden_size=128, num_hidden_layers=2, num_attention_heads=2,
                         intermediate_size=512, max_position_embeddings=32, hidden_dropout_prob=0.0,
                         attention_probs_dropout_prob=0.0)
    log, records = _capture("pretrain_text_test")
    model, trainer, hit10 = run.train(args, True, 0, data, Log_file=log, bert_config=cfg, users_per_pass=16,
                                      model_dir=str(tmp_path))
    losses = [float(m.split(":")[-1]) for m in records if "mean batch loss" in m]
    assert len(losses) == 3 and losses[-1] < losses[0], losses
    assert 0.0 <= hit10 <= 1.0 and sum("valid_results" in m for m in records) == 3
    names = [n for n, _, _ in trainer.names]
    assert any("bert_model.encoder.layer.1." in n for n in names) and not any("bert_model.encoder.layer.0." in n for n in names)
    assert not any("bert_model.embeddings" in n or "pooler" in n for n in names)
    assert sorted(f for f in os.listdir(tmp_path)) == ["epoch-1.pt", "epoch-2.pt", "epoch-3.pt"]
    frozen = model.bert_encoder.text_encoders.title.bert_model.encoder.layer[0].output.dense.weight
    ckpt = torch.load(os.path.join(tmp_path, "epoch-3.pt"), weights_only=False)
    assert set(ckpt["model_state_dict"]) == set(model.state_dict())
    assert torch.equal(ckpt["model_state_dict"]["bert_encoder.text_encoders.title.bert_model.encoder.layer.0.output.dense.weight"]
                       .cpu(), frozen.detach().cpu()) and frozen.grad is None
    args.load_ckpt_name, args.epoch = "epoch-3.pt", 1
    _, trainer2, _ = run.train(args, True, 0, data, Log_file=log, bert_config=cfg, users_per_pass=16, model_dir=str(tmp_path))
    assert trainer2.step_count == 12 and os.path.exists(os.path.join(tmp_path, "epoch-4.pt"))
    assert 0.0 <= run.test(args, True, 0, data, Log_file=log, bert_config=cfg, model_dir=str(tmp_path)) <= 1.0
    assert any("test_results" in m for m in records)


@pytest.mark.parametrize("downstream", [False, True])
def test_cv_full_finetuning_entries_train_and_evaluate(tmp_path, downstream):
    """Pretraining/CV/run.py (downstream=False) and Downstream/CV/run.py (True) mirrors: ViT with its embeddings + layer 0 frozen
    by --freeze_paras_before, the rest fine-tuned without adapters; the downstream script also ranks the test users every
    epoch and starts from the pre-training checkpoint's format."""
    import os
    from adapter4rec_b200.cv import ViTConfigLite
    if downstream:
        from adapter4rec_b200.cv import run
        from adapter4rec_b200.cv.parameters import parse_args
    else:
        from adapter4rec_b200.pretraining import cv_run as run
        from adapter4rec_b200.pretraining.cv_parameters import parse_args
    args = parse_args(["--CV_model_load", "vit-base-patch16-224", "--CV_resize", "48", "--batch_size", "16", "--epoch", "2",
                       "--freeze_paras_before", "20", "--lr", "2e-3", "--fine_tune_lr", "2e-4", "--max_seq_len", "6",
                       "--drop_rate", "0.0"])
    run.setup_seed(12345)
    data = run.synthetic_data(item_num=60, users=48, resize=48, max_seq_len=6)
    cfg = ViTConfigLite(hidden_size=768, num_hidden_layers=2, num_attention_heads=12, intermediate_size=256, image_size=48,
                        patch_size=16)
    log, records = _capture("cv_full_ft_test_%d" % downstream)
    model, trainer, hit10 = run.train(args, True, 0, data, Log_file=log, vit_config=cfg, users_per_pass=8, model_dir=str(tmp_path))
    losses = [float(m.split(":")[-1]) for m in records if "mean batch loss" in m]
    assert len(losses) == 2 and losses[-1] < losses[0], losses
    assert 0.0 <= hit10 <= 1.0 and sum("valid_results" in m for m in records) == 2
    assert sum("test_results" in m for m in records) == (2 if downstream else 0)
    names = [n for n, _, _ in trainer.names]
    assert any("vit.encoder.layer.1." in n for n in names) and any("classifier" in n for n in names)
    assert not any("vit.encoder.layer.0." in n or "vit.embeddings" in n for n in names)
    assert sorted(os.listdir(tmp_path)) == ["epoch-1.pt", "epoch-2.pt"]
    ckpt = torch.load(os.path.join(tmp_path, "epoch-2.pt"), weights_only=False)
    assert set(ckpt["model_state_dict"]) == set(model.state_dict())


def test_two_stage_workflow_from_tsv_files(tmp_path, monkeypatch):
    """The reference's whole workflow through the launchable mains, from files: (1) `pretraining.text_run.main` — Pretraining/
    Text/run.py's flags — reads the news / behaviours TSVs with the real BertTokenizer, fine-tunes the unfrozen tail and writes
    epoch-N.pt under the reference's checkpoint directory name; (2) `run.main` — Downstream/Text/run.py's flags — loads that
    file through --pretrained_model_dir / --pretrained_model_name BEFORE inserting Houlsby adapters (run.py:374-381), trains
    only the adapters and evaluates; (3) --mode test ranks the test users from the saved adapter checkpoint."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import preprocess_fixture as F
    from adapter4rec_b200 import run
    from adapter4rec_b200.pretraining import text_run
    monkeypatch.chdir(tmp_path)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)
    logging.getLogger("Log_file").handlers.clear()
    name = F.write_tiny_body(str(tmp_path / "pretrained_models"))
    data_flags = ["--root_data_dir", os.path.dirname(F.DIR), "--dataset", os.path.basename(F.DIR), "--news", "news.tsv",
                  "--behaviors", "behaviors.tsv", "--max_seq_len", str(F.MAX_SEQ_LEN), "--min_seq_len", str(F.MIN_SEQ_LEN),
                  "--num_words_title", str(F.NUM_WORDS), "--bert_model_load", name, "--embedding_dim", "64",
                  "--batch_size", "16", "--drop_rate", "0.0", "--local_rank", "0"]
    root = str(tmp_path / "pretrained_models")
    model, trainer, _ = text_run.main(data_flags + ["--epoch", "2", "--freeze_paras_before", "21", "--lr", "1e-3",
                                                     "--fine_tune_lr", "1e-4"], pretrained_root=root, users_per_pass=8)
    assert trainer.step_count == 2 * 4                                # 53 users / 16 per batch, 2 epochs
    stage1 = [os.path.join(d, f) for d, _, fs in os.walk(tmp_path) for f in fs if f == "epoch-2.pt"]
    assert len(stage1) == 1 and "checkpoint_modal_%s_freeze_21" % name in stage1[0]
    body_after_stage1 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model2, trainer2, hit10 = run.main(
        data_flags + ["--epoch", "2", "--adapter_type", "houslby", "--adding_adapter_to", "all", "--bert_adapter_down_size", "16",
                      "--adapter_bert_lr", "5e-3", "--adapter_sasrec_lr", "5e-3", "--pretrained_model_dir",
                      os.path.dirname(stage1[0]), "--pretrained_model_name", "epoch-2"], pretrained_root=root, users_per_pass=8)
    assert 0.0 <= hit10 <= 1.0 and all("adapter" in n for n, _, _ in trainer2.names)
    sd2 = model2.state_dict()
    for k, v in body_after_stage1.items():                            # the frozen backbone IS the pre-trained one
        k2 = k if k in sd2 else None
        if k2 is not None:
            assert torch.equal(sd2[k2], v), k
    shared = [k for k in body_after_stage1 if k in sd2]          # the wrapped sub-layers carry other key names
    assert len(shared) >= 20 and "bert_encoder.text_encoders.title.bert_model.embeddings.word_embeddings.weight" in shared
    assert any(k.startswith("user_encoder.") for k in shared)
    stage2 = sorted(os.path.join(d, f) for d, _, fs in os.walk(tmp_path) for f in fs
                    if f.startswith("epoch-") and "add_adapter_to_all" in d)
    assert stage2
    hit_test = run.main(data_flags + ["--mode", "test", "--adapter_type", "houslby", "--adding_adapter_to", "all",
                                      "--bert_adapter_down_size", "16", "--adapter_bert_lr", "5e-3", "--adapter_sasrec_lr", "5e-3",
                                      "--pretrained_model_dir", os.path.dirname(stage1[0]), "--pretrained_model_name", "epoch-2",
                                      "--load_ckpt_name", os.path.basename(stage2[-1])], pretrained_root=root)
    assert 0.0 <= hit_test <= 1.0
    logging.getLogger("Log_file").handlers.clear()
