"""CPU: the host side of the recorded train step (trainer.train_step_graphed): the indirect dropout seed encoding
(A4R_SEED_INDIRECT), the per-replay seeds and Adam's bias-correction factors as the library computes them
(csrc/loss_adam.cu: a4r_adam_step).  The replay itself needs a GPU (tests/test_graph_step_gpu.py)."""
import math
import re
import os

import numpy as np

from adapter4rec_b200 import functional as Fn
from adapter4rec_b200 import ops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_indirect_seed_encoding_matches_header():
    src = open(os.path.join(ROOT, "include", "adapter4rec.h")).read()
    assert re.search(r"#define\s+A4R_SEED_INDIRECT\s+\(1ull\s*<<\s*63\)", src)
    assert Fn.DropoutState.SEED_INDIRECT == 1 << 63
    saved = (Fn.DropoutState.seed, Fn.DropoutState.counter)
    try:
        Fn.DropoutState.manual_seed(123456)
        assert Fn.DropoutState.draw(10) == (123456, 0)
        Fn.DropoutState.seed_address = 0x7F0012345678
        seed, off = Fn.DropoutState.draw(6)
        assert seed == (1 << 63) | 0x7F0012345678 and off == 10 and seed < 1 << 64
        Fn.DropoutState.seed_address = None
        assert Fn.DropoutState.draw(1) == (123456, 16)          # the counter keeps running through a recording
    finally:
        Fn.DropoutState.seed_address = None
        Fn.DropoutState.seed, Fn.DropoutState.counter = saved


def test_replay_seeds_are_distinct_direct_seeds():
    seeds = [Fn.DropoutState.replay_seed(t, base=77) for t in range(1, 2001)]
    assert len(set(seeds)) == len(seeds)
    assert all(0 <= s < 1 << 63 for s in seeds)                 # bit 63 clear: never mistaken for an address
    assert Fn.DropoutState.replay_seed(5, base=77) != Fn.DropoutState.replay_seed(5, base=78)
    bits = np.array([[(s >> b) & 1 for b in range(63)] for s in seeds], dtype=np.float64)
    assert np.all(np.abs(bits.mean(0) - 0.5) < 0.06)           # every bit of the seed moves from replay to replay


def test_bias_corrections_follow_the_library_arithmetic():
    """a4r_adam_step receives the betas as C floats and evaluates 1 - pow((double)beta, step) in double."""
    b1, b2 = float(np.float32(0.9)), float(np.float32(0.999))
    for step in (1, 2, 10, 1000, 100000):
        bc1, bc2s = ops.adam_bias_corrections(0.9, 0.999, step)
        assert bc1 == 1.0 - math.pow(b1, step)
        assert bc2s == math.sqrt(1.0 - math.pow(b2, step))
    assert ops.adam_bias_corrections(0.9, 0.999, 1)[0] == 1.0 - b1


def test_recording_is_refused_for_value_dependent_steps():
    """unpadded-token and dedup layouts read counts back from the batch (host syncs, data-dependent shapes); bucketed overlap
    issues NCCL from gradient hooks: none of them can be baked into a recording — refused with the reason, never mis-recorded"""
    import torch
    from adapter4rec_b200.trainer import FlatAdamTrainer
    m = torch.nn.ModuleDict({"bert_encoder": torch.nn.ModuleDict({"adapter": torch.nn.Linear(2, 2)}),
                             "user_encoder": torch.nn.Linear(2, 2)})
    tr = FlatAdamTrainer(m, 1e-3, 1e-3, 1e-3, 1e-3)
    assert tr.graph_unsafe_reason() is None
    m["bert_encoder"].unpad = True
    assert "unpad" in tr.graph_unsafe_reason()
    m["bert_encoder"].unpad = False
    m.dedup_items = True
    assert "dedup_items" in tr.graph_unsafe_reason()
    m.dedup_items = False
    tr.buckets = [[0, 1, 1]]
    assert "overlap=False" in tr.graph_unsafe_reason()
