"""Multi-GPU check under real NCCL (run with torchrun, N >= 2):
  torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/dist_check_gpu.py
(1) full-ranking eval with the item table sharded by item id (all-gather of partial top-10 lists + merge) gives the SAME
    top-10 ids / HR / NDCG as the unsharded evaluation of the gathered table;
(2) data-parallel training: after two steps on different per-rank batches every rank holds bit-identical parameters
    (one flat-gradient all-reduce per step) and the loss is finite;
(3) full fine-tuning with the backward-overlapped bucketed reduction (64 KB buckets issued from gradient hooks) gives the
    gradients of the single blocking all-reduce;
(4) the train step replayed from CUDA graphs (trainer.train_step_graphed: two recordings around one eager NCCL all-reduce) is
    bit-identical to the eager data-parallel step on every rank, dropout on.
Prints one line per check; exit code 0 iff all pass."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import cases  # noqa: E402
from test_model_gpu import build_gpu_model  # noqa: E402
from adapter4rec_b200.data_utils.metrics import ItemTable, build_eval_arrays, eval_arrays, get_item_embeddings  # noqa: E402
from adapter4rec_b200.trainer import FlatAdamTrainer  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
ok = True

c = cases.tiny_case("lora")
sd = cases.build_state_dict(c)
model, args = build_gpu_model(c, sd)
items = cases.build_item_content(c)
seqs, hist = cases.build_eval_users(c)

# ---- (1) sharded == unsharded evaluation
table = get_item_embeddings(model, items.numpy(), 16, args, True, dev)
assert table.world == world and table.n_local < items.shape[0]
tok, mask, tgt, hs = [torch.from_numpy(a) for a in build_eval_arrays({i: s for i, s in enumerate(seqs)},
                                                                      {i: torch.LongTensor(h) for i, h in enumerate(hist)}, c.S)]
hit, ndcg, ids = eval_arrays(model, tok, mask, tgt, hs, table, 4)
per = (items.shape[0] + world - 1) // world
pad = torch.zeros((per, table.dim), dtype=torch.bfloat16, device=dev)
pad[:table.n_local] = table.shard
full = torch.empty((world * per, table.dim), dtype=torch.bfloat16, device=dev)
dist.all_gather_into_tensor(full, pad)
rows = torch.cat([full[r * per: r * per + min(per, max(0, items.shape[0] - r * per))] for r in range(world)])
one = ItemTable(rows, 0, items.shape[0], rank=0, world=1)
hit1, ndcg1, ids1 = eval_arrays(model, tok, mask, tgt, hs, one, 4)
same = torch.equal(ids, ids1) and torch.equal(hit, hit1) and torch.equal(ndcg, ndcg1)
ok &= same
if rank == 0:
    print("sharded eval == unsharded eval (ids, HR, NDCG bit-exact):", same, "| HR@10 %.4f" % float(hit.mean()))

# ---- (2) data-parallel step: parameters stay identical across ranks
model.train()
trainer = FlatAdamTrainer(model, 1e-3, 1e-4, 1e-3, 1e-3, users_per_pass=4)
sample_items, log_mask, _ = cases.build_batch(c, items)
B = sample_items.shape[0]
half = slice(rank * (B // world), (rank + 1) * (B // world))
x = sample_items[half].reshape(-1, 2 * c.L).to(dev)
lm = log_mask[half].to(dev)
for _ in range(2):
    loss = trainer.train_step(x, lm)
flat = trainer.flat_param.detach().clone()
ref = flat.clone()
dist.broadcast(ref, 0)
same = torch.equal(flat, ref) and bool(torch.isfinite(torch.as_tensor(float(loss))))
ok &= same
if rank == 0:
    print("parameters bit-identical on all ranks after 2 DP steps:", same, "| loss %.5f" % float(loss))

# ---- (3) bucketed, backward-overlapped reduction == one blocking all-reduce (full fine-tuning: every tensor trainable)
cf = cases.tiny_case("full_ft")
sdf = cases.build_state_dict(cf)
itf = cases.build_item_content(cf)
sif, lmf, _ = cases.build_batch(cf, itf)
Bf = sif.shape[0]
halff = slice(rank * (Bf // world), (rank + 1) * (Bf // world))
xf, lf = sif[halff].reshape(-1, 2 * cf.L).to(dev), lmf[halff].to(dev)
grads = []
names3 = None
for overlap in (True, False, False):
    mf, _ = build_gpu_model(cf, sdf)
    mf.eval()                                  # no dropout: both runs see the same arithmetic
    tf = FlatAdamTrainer(mf, 1e-3, 1e-4, 1e-3, 1e-3, users_per_pass=4, bucket_bytes=64 << 10, overlap=overlap)
    tf.zero_grad()
    tf.forward_backward(xf, lf)
    issued_early = len(tf._works)
    tf.reduce_gradients()
    torch.cuda.synchronize()
    grads.append(tf.flat_grad.clone())
    names3 = tf.names
    if overlap:
        nb, early = len(tf.buckets), issued_early
err = float((grads[0] - grads[1]).abs().max() / grads[1].abs().max())
# Two blocking runs differ from each other by the same amount: the embedding-table gradients are fp32 atomics
# (a4r_scatter_add_rows: red.global.add, the order of the adds of a repeated token is not fixed), everything else is bit-equal.
run_to_run = float((grads[1] - grads[2]).abs().max() / grads[1].abs().max())
atomic = ("word_embeddings", "position_embeddings", "token_type_embeddings")
exact = all(torch.equal(grads[0][off:off + k], grads[1][off:off + k]) for n, off, k in names3 if not any(a in n for a in atomic))
differing = [n for n, off, k in names3 if not torch.equal(grads[0][off:off + k], grads[1][off:off + k])]
same3 = exact and err < 1e-6 and nb > 1 and early > 0
ok &= same3
if rank == 0:
    print("bucketed overlapped reduction == blocking all-reduce (full fine-tuning, %d buckets, %d issued during the backward):"
          % (nb, early), same3, "| max rel diff %.2e (two blocking runs: %.2e; tensors that differ: %s)" % (err, run_to_run, differing))
# ---- (4) the step replayed from CUDA graphs (two recordings around ONE eager NCCL all-reduce) == the eager DP step
from adapter4rec_b200 import functional as Fn  # noqa: E402
pair = []
for _ in range(2):
    mg, _ = build_gpu_model(c, sd)
    mg.train()                                 # dropout on: the replay reads its seed through the pointer
    pair.append(FlatAdamTrainer(mg, 1e-3, 1e-4, 1e-3, 1e-3, users_per_pass=max(1, x.shape[0] // (2 * (c.S + 1) * 2)),
                                overlap=False))
tg, te = pair
Fn.DropoutState.manual_seed(31 + rank)
tg.train_step_graphed(x, lm)                   # first call: eager
Fn.DropoutState.manual_seed(31 + rank)
te.train_step(x, lm)
same4 = torch.equal(tg.flat_param, te.flat_param)
for t in range(3):
    Fn.DropoutState.manual_seed(31 + rank)
    lg = tg.train_step_graphed(x, lm).clone()
    Fn.DropoutState.seed, Fn.DropoutState.counter = tg.graph_seed(tg.step_count), tg.graph_counter0
    le = te.train_step(x, lm).clone()
    same4 &= torch.equal(lg, le) and torch.equal(tg.flat_param, te.flat_param) and torch.equal(tg.exp_avg, te.exp_avg)
ref4 = tg.flat_param.detach().clone()
dist.broadcast(ref4, 0)
same4 &= torch.equal(ref4, tg.flat_param) and len(tg._graphs) == 2
tg.release_graph()
ok &= same4
if rank == 0:
    print("graphed DP step (2 recordings + 1 eager all-reduce per step) bit-identical to the eager DP step, 3 replays:", same4)

flags = torch.tensor([int(ok)], device=dev)
dist.all_reduce(flags, op=dist.ReduceOp.MIN)
if rank == 0:
    print("ALL OK" if int(flags) else "FAILED")
dist.destroy_process_group()
sys.exit(0 if int(flags) else 1)
