"""Entry points of the text tree, mirroring Downstream/Text/run.py: `train(args, use_modal, local_rank, data)` builds
the model exactly as run.py:286-529 does (BERT/RoBERTa body -> Model / ModelCPC -> freeze -> adapter surgery ->
LayerNorm unfreeze -> 4 learning-rate groups), trains for args.epoch epochs with the step of run.py:586-600 and
evaluates with get_item_embeddings + eval_model (run.py:649-670).

The reference's TSV readers / tokeniser (data_utils/preprocess.py) are host-side and out of scope: the
caller hands in the arrays they produce — `item_content` [I+1, 2L] int (ids | mask rows), `users_train` {uid: [item
ids]}, and the eval dictionaries — via `data`.  `synthetic_data` builds such arrays for smoke runs."""
import logging
import random
import types

import numpy as np
import torch
import torch.distributed as dist

from . import surgery
from .data_utils.dataset import BuildTrainDataset
from .data_utils.metrics import eval_model, get_item_embeddings
from .model import BertModel, Model, ModelCPC, RobertaModel, TextConfigLite
from .trainer import FlatAdamTrainer


def setup_seed(seed):      # run.py:673-678
    torch.manual_seed(seed)
    torch.cuda.manual_seed_all(seed)
    np.random.seed(seed)
    random.seed(seed)


def build_train_batch(users, u2seq, item_content, item_num, max_seq_len):
    """BuildTrainDataset.__getitem__ (data_utils/dataset.py:24-49) for a list of users, stacked:
    sample_items [B, S+1, 2, 2L] int64, log_mask [B, S] float32 (left padded, one uniform negative per position)."""
    S1 = max_seq_len + 1
    ids = np.zeros((len(users), S1, 2), dtype=np.int64)
    log_mask = np.zeros((len(users), max_seq_len), dtype=np.float32)
    for b, u in enumerate(users):
        seq = list(u2seq[u])
        n = len(seq)
        ids[b, S1 - n:, 0] = seq
        log_mask[b, max_seq_len - (n - 1):] = 1.0
        taken = set(seq)
        for i in range(n - 1):
            neg = random.randint(1, item_num)
            while neg in taken:
                neg = random.randint(1, item_num)
            ids[b, S1 - n + i, 1] = neg
    content = torch.as_tensor(np.asarray(item_content)).long()
    return content[torch.from_numpy(ids)], torch.from_numpy(log_mask)


def synthetic_data(item_num=2000, users=256, num_words=30, max_seq_len=20, vocab=30522, seed=0):
    rng = np.random.RandomState(seed)
    item_content = np.zeros((item_num + 1, 2 * num_words), dtype=np.int64)
    for i in range(1, item_num + 1):
        n = rng.randint(8, num_words + 1)
        item_content[i, :n] = rng.randint(min(1000, vocab // 2), vocab, n)
        item_content[i, 0], item_content[i, n - 1] = min(101, vocab - 2), min(102, vocab - 1)
        item_content[i, num_words:num_words + n] = 1
    seqs = {u: (rng.permutation(item_num)[:rng.randint(5, max_seq_len + 4)] + 1).tolist() for u in range(users)}
    d = types.SimpleNamespace(item_content=item_content, item_num=item_num)
    d.users_train = {u: s[:-2][-(max_seq_len + 1):] for u, s in seqs.items()}
    d.users_valid = {u: s[:-1][-(max_seq_len + 1):] for u, s in seqs.items()}
    d.users_test = {u: s[-(max_seq_len + 1):] for u, s in seqs.items()}
    d.users_history_for_valid = {u: torch.LongTensor(s[:-2]) for u, s in seqs.items()}
    d.users_history_for_test = {u: torch.LongTensor(s[:-1]) for u, s in seqs.items()}
    return d


def build_model(args, item_num, local_rank, bert_config=None, bert_state_dict=None):
    """run.py:286-503 without the DDP wrap (the trainer owns the gradient all-reduce)."""
    cfg = bert_config if bert_config is not None else TextConfigLite()
    roberta = 'roberta' in args.bert_model_load
    bert_model = (RobertaModel if roberta else BertModel)(cfg)
    if bert_state_dict is not None:
        bert_model.load_state_dict(bert_state_dict, strict=False)
    model = (ModelCPC if "cpc" in args.arch else Model)(args, item_num, True, bert_model).to(local_rank)
    if 'None' in args.fine_tune_to:
        surgery.freeze_all(model)
    elif 'all' not in args.fine_tune_to:
        raise AssertionError("fine_tune_to should be defined properly")
    model = surgery.insert_adapters(model, args)
    surgery.unfreeze_layernorm(model, args)
    return model


def train(args, use_modal, local_rank, data, Log_file=None, bert_config=None, users_per_pass=128):
    Log_file = Log_file or logging.getLogger("adapter4rec_b200")
    model = build_model(args, data.item_num, local_rank, bert_config)
    trainer = FlatAdamTrainer(model, args.lr, args.fine_tune_lr, args.adapter_bert_lr, args.adapter_sasrec_lr,
                              users_per_pass=users_per_pass)
    Log_file.info("##### trainable_num {} #####".format(trainer.num_trainable))
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    users = sorted(data.users_train.keys())
    train_ds = BuildTrainDataset(data.users_train, data.item_content, data.item_num, args.max_seq_len, True,
                                 device=next(model.parameters()).device, seed=123456 + rank)   # run.py:686 seed; per-rank stream
    max_hit10 = 0.0
    for ep in range(args.epoch):
        model.train()
        random.Random(ep).shuffle(users)
        mine = users[rank::world]                                   # DistributedSampler's round-robin split
        loss_sum, batches = 0.0, 0
        for b0 in range(0, len(mine), args.batch_size):
            # negatives + token-row gather on the device (only the user indices cross PCIe); the host restatement
            # build_train_batch above stays as the reference-order implementation for CPU-side tools
            items, log_mask = train_ds.batch(mine[b0:b0 + args.batch_size])
            loss = trainer.train_step(items.view(-1, items.size(-1)), log_mask)
            loss_sum, batches = loss_sum + float(loss), batches + 1
            if loss != loss:                                        # NaN guard of run.py:602-604
                raise FloatingPointError("loss is NaN")
        Log_file.info('epoch {} mean batch loss: {:.5f}'.format(ep + 1, loss_sum / max(1, batches)))
        hit10 = run_eval(model, data, args, Log_file, "valid", local_rank)
        max_hit10 = max(max_hit10, hit10)
    return model, trainer, max_hit10


def run_eval(model, data, args, Log_file, v_or_t, local_rank, batch_size=512):
    """run.py:649-670: item table, then full-ranking HR@10 / NDCG@10."""
    table = get_item_embeddings(model, data.item_content, batch_size, args, True, local_rank)
    hist, seqs = (data.users_history_for_valid, data.users_valid) if v_or_t == "valid" else \
        (data.users_history_for_test, data.users_test)
    return eval_model(model, hist, seqs, table, batch_size, args, data.item_num, Log_file, v_or_t, local_rank)
