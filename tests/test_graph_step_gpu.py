"""GPU: the train step replayed from a CUDA graph (FlatAdamTrainer.train_step_graphed; SURVEY.md 8e "capture the step in
a CUDA graph per rank", 8f-2 "the step under one CUDA graph") against the eager step of Downstream/Text/run.py:586-600
restated in trainer.train_step: same model, same batches, dropout ON.  The recording reads its dropout seed through a
pointer (A4R_SEED_INDIRECT) and Adam's bias corrections from device memory (a4r_adam_step_dev); an eager step given the
replay's seed and the recording's first counter draws the same masks, so parameters, Adam moments and losses must be
BIT-identical after every step — and the evaluator that runs after the replays must see the updated weights."""
import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
sys.path.insert(0, HERE)

pytestmark = pytest.mark.gpu


def _setup(kind):
    import cases
    from test_model_gpu import build_gpu_model
    from adapter4rec_b200.trainer import FlatAdamTrainer
    c = cases.tiny_case(kind)
    sd = cases.build_state_dict(c)
    out = []
    for _ in range(2):
        model, _ = build_gpu_model(c, sd)
        model.train()
        out.append((model, FlatAdamTrainer(model, 1e-3, 1e-4, 1e-3, 1e-3, users_per_pass=max(1, c.B // 2), overlap=False)))
    items = cases.build_item_content(c)
    sample_items, log_mask, _ = cases.build_batch(c, items)
    return c, out, sample_items, log_mask


@pytest.mark.parametrize("kind", ["lora", "houlsby"])
def test_graphed_step_is_bit_identical_to_eager(kind):
    from adapter4rec_b200 import functional as Fn
    c, ((model_g, tr_g), (model_e, tr_e)), sample_items, log_mask = _setup(kind)
    rows_per_user = (c.S + 1) * 2

    def batch(t):           # a different batch every step: the users rotated by t
        si = torch.roll(sample_items, shifts=t, dims=0).reshape(c.B * rows_per_user, 2 * c.L).cuda()
        return si, torch.roll(log_mask, shifts=t, dims=0).cuda()

    # step 1 runs eagerly in both trainers (the graphed trainer's first call loads the kernels and builds the caches)
    Fn.DropoutState.manual_seed(77)
    l_g = tr_g.train_step_graphed(*batch(0)).clone()
    Fn.DropoutState.manual_seed(77)
    l_e = tr_e.train_step(*batch(0)).clone()
    assert torch.equal(l_g, l_e) and torch.equal(tr_g.flat_param, tr_e.flat_param)
    for t in range(1, 5):
        Fn.DropoutState.manual_seed(77)                    # base seed of the recording (made at t == 1)
        l_g = tr_g.train_step_graphed(*batch(t)).clone()
        assert tr_g.step_count == t + 1
        Fn.DropoutState.seed, Fn.DropoutState.counter = tr_g.graph_seed(tr_g.step_count), tr_g.graph_counter0
        l_e = tr_e.train_step(*batch(t)).clone()
        assert torch.isfinite(l_g).all()
        assert torch.equal(l_g, l_e), "step %d: loss %r vs %r" % (t, float(l_g), float(l_e))
        assert torch.equal(tr_g.flat_param, tr_e.flat_param), "step %d: parameters differ" % t
        assert torch.equal(tr_g.exp_avg, tr_e.exp_avg) and torch.equal(tr_g.exp_avg_sq, tr_e.exp_avg_sq)
    assert len(tr_g._graphs) == 1 and tr_g._g_replays == 4
    # replays draw fresh masks: the same batch replayed twice gives two different losses (dropout p = 0.1 is on)
    a = tr_g.train_step_graphed(*batch(0)).clone()
    b = tr_g.train_step_graphed(*batch(0)).clone()
    assert not torch.equal(a, b)
    for _ in range(2):
        Fn.DropoutState.seed, Fn.DropoutState.counter = tr_g.graph_seed(tr_e.step_count + 1), tr_g.graph_counter0
        tr_e.train_step(*batch(0))
    assert torch.equal(tr_g.flat_param, tr_e.flat_param)
    # eager consumers after the replays (the evaluator's encoder) must run on the UPDATED weights: eval-mode loss equal
    model_g.eval(), model_e.eval()
    si, lm = batch(3)
    with torch.no_grad():
        assert torch.equal(model_g(si, lm, 0), model_e(si, lm, 0))
    tr_g.release_graph()


def test_graphed_step_records_again_for_new_shapes():
    from adapter4rec_b200 import functional as Fn
    c, ((model_g, tr_g), _), sample_items, log_mask = _setup("lora")
    rows_per_user = (c.S + 1) * 2
    si = sample_items.reshape(c.B * rows_per_user, 2 * c.L).cuda()
    lm = log_mask.cuda()
    Fn.DropoutState.manual_seed(5)
    tr_g.train_step_graphed(si, lm)
    tr_g.train_step_graphed(si, lm)
    first = tr_g._graphs[0]
    half = c.B // 2
    loss = tr_g.train_step_graphed(si[:half * rows_per_user], lm[:half])
    second = tr_g._graphs[0]
    assert second is not first and torch.isfinite(loss).all() and tr_g.step_count == 3
    # back to the first shape: its recording is still there (an epoch alternates full batches and the last, shorter one)
    tr_g.train_step_graphed(si, lm)
    assert tr_g._graphs[0] is first and len(tr_g._recordings) == 2
    # a third shape drops the least recently used recording (the half batch), never the one just replayed
    one = rows_per_user
    loss = tr_g.train_step_graphed(si[:one], lm[:1])
    assert len(tr_g._recordings) == 2 and tr_g._graphs[0] not in (first, second) and torch.isfinite(loss).all()
    tr_g.train_step_graphed(si, lm)
    assert tr_g._graphs[0] is first and tr_g.step_count == 6
    tr_g.release_graph()
    assert tr_g._recordings == {} and tr_g._graphs is None


def test_entry_script_loop_with_graphed_steps(tmp_path):
    """run.train(graphed=True): the training loop of Downstream/Text/run.py with every step replayed from a CUDA graph — 96
    users at batch 40 give two full batches and a 16-user tail per epoch, i.e. two recordings that alternate; same losses as
    the eager loop (no dropout in this configuration: the two runs are the same arithmetic)."""
    import logging
    from adapter4rec_b200 import run
    from adapter4rec_b200.model import TextConfigLite
    from adapter4rec_b200.parameters import parse_args
    flags = ["--embedding_dim", "64", "--batch_size", "40", "--epoch", "2", "--adapter_type", "houslby",
             "--adding_adapter_to", "all", "--bert_model_load", "bert_tiny", "--bert_adapter_down_size", "16",
             "--adapter_bert_lr", "5e-3", "--adapter_sasrec_lr", "5e-3", "--max_seq_len", "10", "--num_words_title", "12",
             "--drop_rate", "0.0", "--adapter_dropout_rate", "0.0", "--pretrained_model_name", "None"]
    cfg = TextConfigLite(vocab_size=500, hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=512,
                         max_position_embeddings=32, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    losses = {}
    for graphed in (False, True):
        run.setup_seed(123456)
        data = run.synthetic_data(item_num=300, users=96, num_words=12, max_seq_len=10, vocab=500)
        log = logging.getLogger("graphed_loop_%d" % graphed)
        records = []
        log.addHandler(type("H", (logging.Handler,), {"emit": lambda self, r, rec=records: rec.append(r.getMessage())})())
        log.setLevel(logging.INFO)
        _, trainer, _ = run.train(parse_args(flags), True, 0, data, Log_file=log, bert_config=cfg, users_per_pass=8,
                                  graphed=graphed)
        assert trainer.step_count == 6
        losses[graphed] = [float(m.split(":")[-1]) for m in records if "mean batch loss" in m]
    assert len(losses[True]) == 2 and losses[True] == losses[False], losses
