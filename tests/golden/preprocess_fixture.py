"""Shared by tests/golden/make_golden_preprocess.py (runs the UNMODIFIED reference readers) and tests/test_preprocess_cpu.py (runs
adapter4rec_b200.data_utils.preprocess): the fixture files and a deterministic stand-in tokenizer with the call signature the
reference uses on transformers' tokenizers (Downstream/Text/data_utils/preprocess.py:86)."""
import os
import random
import zlib

HERE = os.path.dirname(os.path.abspath(__file__))
DIR = os.path.join(HERE, "preprocess")
NEWS = os.path.join(DIR, "news.tsv")
BEHAVIORS = os.path.join(DIR, "behaviors.tsv")
MAX_SEQ_LEN, MIN_SEQ_LEN, NUM_WORDS = 6, 4, 9


def toy_tokenizer(text, max_length, padding, truncation):
    """[CLS]=101, one crc32-derived id per whitespace word (truncated to max_length - 2), [SEP]=102, zero padding."""
    assert padding == 'max_length' and truncation is True
    ids = [101] + [1000 + zlib.crc32(w.encode()) % 29000 for w in text.split()][:max_length - 2] + [102]
    mask = [1] * len(ids)
    pad = max_length - len(ids)
    return {'input_ids': ids + [0] * pad, 'attention_mask': mask + [0] * pad}


def write_fixture():
    """120 news lines (mixed case, one empty title, one title longer than the token budget) and 80 behaviour lines: users
    below the minimum length, users longer than max_seq_len + 3 (truncated from the left), items nobody keeps (dropped and
    the ids renumbered), a user name that appears twice (first position, last sequence, both lines counted)."""
    rng = random.Random(20240229)
    os.makedirs(DIR, exist_ok=True)
    words = ["Oslo", "fjord", "News", "vinter", "sport", "Trump", "bus", "tram", "photo", "stamp", "holiday", "fire", "year"]
    names = ["n%04x" % (i * 37 + 11) for i in range(120)]
    with open(NEWS, "w") as f:
        for i, n in enumerate(names):
            k = 0 if i == 17 else 14 if i == 5 else rng.randint(1, 8)
            f.write("%s\t%s\n" % (n, " ".join(rng.choice(words) for _ in range(k))))
    pool = names[:90]                      # 30 news never appear in a behaviour line
    with open(BEHAVIORS, "w") as f:
        for u in range(80):
            n = rng.choice([1, 2, 3, 4, 5, 6, 8, 9, 10, 12, 15])
            name = "u%03d" % (u if u != 40 else 7)          # line 40 repeats user u007
            f.write("%s\t%s\n" % (name, " ".join(rng.sample(pool, n))))


def write_tiny_body(root, name="bert_tiny_fixture", hidden=128, layers=2):
    """<root>/bert/<name>/{config.json, vocab.txt}: the directory layout Downstream/Text/run.py:295-300 reads a body from (no
    weights file: the body is then randomly initialised, as in every offline run here)."""
    import json
    path = os.path.join(root, "bert", name)
    os.makedirs(path, exist_ok=True)
    words = ["oslo", "fjord", "news", "vinter", "sport", "trump", "bus", "tram", "photo", "stamp", "holiday", "fire", "year"]
    vocab = ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"] + words
    with open(os.path.join(path, "vocab.txt"), "w") as f:
        f.write("\n".join(vocab) + "\n")
    with open(os.path.join(path, "config.json"), "w") as f:
        json.dump({"model_type": "bert", "vocab_size": len(vocab), "hidden_size": hidden, "num_hidden_layers": layers,
                   "num_attention_heads": 2, "intermediate_size": 4 * hidden, "max_position_embeddings": 32,
                   "type_vocab_size": 2, "layer_norm_eps": 1e-12, "pad_token_id": 0, "hidden_dropout_prob": 0.0,
                   "attention_probs_dropout_prob": 0.0, "architectures": ["BertModel"]}, f)
    return name
