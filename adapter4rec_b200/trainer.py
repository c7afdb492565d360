"""Data-parallel adapter-tuning step: forward + backward through the sm_100a kernels, sum all-reduce of the flat
trainable-gradient buffer over NCCL — one collective for adapter tuning (<= 10 MB), `bucket_bytes` slices issued from
gradient hooks while the backward is still running for full fine-tuning (440 MB) — and fused flat Adam.  Mirrors the
optimisation set-up of Downstream/Text/run.py:503-529,595-600 (4 learning-rate groups chosen by parameter NAME,
torch.optim.Adam defaults, DistributedDataParallel's gradient averaging and bucketed overlap)."""
import os

import torch
import torch.distributed as dist

from . import functional as Fn
from . import ops


def group_parameters(model):
    """run.py:505-523: ('bert_encoder' in name) x ('adapter' in name or 'lora' in name)."""
    groups = {"bert": [], "recsys": [], "adapter_bert": [], "adapter_recsys": []}
    for name, param in model.named_parameters():
        if not param.requires_grad:
            continue
        is_adapter = "adapter" in name or "lora" in name
        if "bert_encoder" in name:
            groups["adapter_bert" if is_adapter else "bert"].append((name, param))
        else:
            groups["adapter_recsys" if is_adapter else "recsys"].append((name, param))
    return groups


class FlatAdamTrainer:
    """Owns one contiguous fp32 buffer each for trainable parameters, gradients and Adam moments.

    Each trainable nn.Parameter becomes a view into the flat parameter buffer and its .grad a view into the flat
    gradient buffer, so the per-step communication is exactly one all-reduce and the optimizer is one
    kernel launch per learning-rate group."""

    def __init__(self, model, lr, fine_tune_lr, adapter_bert_lr, adapter_sasrec_lr, betas=(0.9, 0.999), eps=1e-8,
                 weight_decay=0.0, users_per_pass=128, process_group=None, grouping=None, bucket_bytes=25 << 20,
                 overlap=None):
        self.model = model
        self.betas, self.eps, self.weight_decay = betas, eps, weight_decay
        self.users_per_pass = users_per_pass
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        groups = (grouping or group_parameters)(model)      # the image tree groups by other names (cv/run_adapter.py)
        lrs = {"bert": fine_tune_lr, "recsys": lr, "adapter_bert": adapter_bert_lr, "adapter_recsys": adapter_sasrec_lr}
        plist = [(g, n, p) for g in ("bert", "recsys", "adapter_bert", "adapter_recsys") for n, p in groups[g]]
        self._plist, self._lrs, self._group_order = plist, lrs, ("bert", "recsys", "adapter_bert", "adapter_recsys")
        if not plist:
            raise ValueError("no trainable parameters")
        dev = plist[0][2].device
        total = sum(p.numel() for _, _, p in plist)
        self.flat_param = torch.empty(total, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(total, dtype=torch.float32, device=dev)
        self.segments = []   # (group, lr, offset, numel)
        self.names = []
        off = 0
        for g in ("bert", "recsys", "adapter_bert", "adapter_recsys"):
            start = off
            for n, p in groups[g]:
                k = p.numel()
                self.flat_param[off:off + k].copy_(p.detach().reshape(-1))
                p.data = self.flat_param[off:off + k].view(p.shape)
                p.grad = self.flat_grad[off:off + k].view(p.shape)
                self.names.append((n, off, k))
                off += k
            if off > start:
                self.segments.append((g, lrs[g], start, off - start))
        self.step_count = 0
        self.num_trainable = total
        # Backward-overlapped reduction (DistributedDataParallel's bucket hooks, run.py:503): the flat buffer is cut into
        # contiguous buckets of <= bucket_bytes; a bucket's all-reduce is issued (async: NCCL runs it on its own stream
        # behind everything the current stream has enqueued so far) as soon as every gradient in it has been accumulated
        # in the LAST pass of the step.  Buckets are issued in one fixed order on every rank — from the end of the buffer
        # (the last layers, whose gradients come first) to the start — whatever order they complete in.
        # overlap=None (default): overlapped only when the job spans more than one node.  Inside one NVSwitch node the blocking
        # all-reduce of the whole 440 MB full-fine-tuning buffer costs ~1.5 ms, less than the SMs the NCCL kernels take from
        # the backward when they run beside it (8 x B200, tools/bench_full_ft.py: 107.2 ms blocking, 109.3 ms overlapped).
        if overlap is None:
            local = int(os.environ.get("LOCAL_WORLD_SIZE", self.world))
            overlap = self.world > max(local, 1)
        self.buckets = []            # [offset, numel, number of parameters]
        self._works, self._live, self._next = [], False, -1
        if self.world > 1 and overlap:
            cap = max(1, int(bucket_bytes) // 4)
            for (g, n, p), (_, off, k) in zip(plist, self.names):
                if not self.buckets or self.buckets[-1][1] + k > cap:
                    self.buckets.append([off, 0, 0])
                b = self.buckets[-1]
                b[1] += k
                b[2] += 1
                p.register_post_accumulate_grad_hook(self._make_hook(len(self.buckets) - 1))
        self._pending = [b[2] for b in self.buckets]
        self._ready = [False] * len(self.buckets)

    def _make_hook(self, b):
        def hook(_param):
            if not self._live:
                return
            self._pending[b] -= 1
            if self._pending[b] == 0:
                self._ready[b] = True
                self._issue_ready()
        return hook

    def _issue_ready(self, force=False):
        """issue the all-reduces of the completed buckets, strictly in descending bucket order (same on every rank);
        force: issue the rest too (parameters that received no gradient this step never fire their hook)"""
        while self._next >= 0 and (force or self._ready[self._next]):
            off, n, _ = self.buckets[self._next]
            self._works.append(dist.all_reduce(self.flat_grad[off:off + n], op=dist.ReduceOp.SUM, group=self.pg,
                                               async_op=True))
            self._next -= 1

    def _arm(self):
        self._pending = [b[2] for b in self.buckets]
        self._ready = [False] * len(self.buckets)
        self._next = len(self.buckets) - 1
        self._live = bool(self.buckets)

    def zero_grad(self):
        self.flat_grad.zero_()

    def forward_backward(self, sample_items, log_mask):
        """sample_items int64 [B*(S+1)*2, 2L], log_mask f32 [B,S] (both on the device).  The batch is processed in
        passes of users_per_pass users to bound activation memory; pass c is weighted count_c / count so that the
        accumulated gradient is exactly that of the batch loss (a mean over ALL valid positions)."""
        model = self.model
        B, S = log_mask.shape
        rows_per_user = sample_items.shape[0] // B
        upp = min(self.users_per_pass, B)
        cpc = getattr(model, "cpc", False)
        total = None
        if not cpc:
            count_all = (log_mask != 0).sum().float()
        for b0 in range(0, B, upp):
            b1 = min(B, b0 + upp)
            lm = log_mask[b0:b1]
            loss_c = model(sample_items[b0 * rows_per_user:b1 * rows_per_user], lm, sample_items.device)
            w = (float(b1 - b0) / B) if cpc else (lm != 0).sum().float() / count_all
            weighted = loss_c * w
            if b1 == B:
                self._arm()      # gradients are final after this pass: buckets may leave while it is running
            weighted.backward()
            total = weighted.detach() if total is None else total + weighted.detach()
        return total

    def reduce_gradients(self):
        """Sum all-reduce of the flat gradient buffer (NCCL on GPUs): whatever the backward hooks have not issued yet is
        issued here, then the current stream waits for all of it.  Without hooks (overlap=False, or gradients written
        outside forward_backward) this is the ONE collective of the step."""
        if self.world <= 1:
            return
        if self._live:
            self._issue_ready(force=True)
            for w in self._works:
                w.wait()
            self._works, self._live = [], False
        else:
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=self.pg)

    def optimizer_step(self):
        self.reduce_gradients()
        self.step_count += 1
        for _, lr, off, n in self.segments:
            ops.adam_step(self.flat_param[off:off + n], self.flat_grad[off:off + n], self.exp_avg[off:off + n],
                          self.exp_avg_sq[off:off + n], lr, self.betas[0], self.betas[1], self.eps, self.weight_decay,
                          self.step_count, grad_scale=1.0 / self.world)
        Fn.bump_param_epoch()

    def train_step(self, sample_items, log_mask):
        self._eager_steps = getattr(self, "_eager_steps", 0) + 1
        self.zero_grad()
        loss = self.forward_backward(sample_items, log_mask)
        self.optimizer_step()
        return loss

    # ---- the step under a CUDA graph (SURVEY.md 8e "capture the step in a CUDA graph per rank", 8f-2) --------------------
    # One rank records zero_grad + every pass of forward/backward + Adam ONCE and replays it per step: ~1,800 launches
    # become one cudaGraphLaunch.  What a recording would bake but must change between steps lives in 16 bytes of device
    # memory the host refreshes before each replay: the dropout seed (A4R_SEED_INDIRECT: the kernels read it through a
    # pointer; the counter offsets stay those of the recording) and Adam's two bias-correction factors
    # (a4r_adam_step_dev).  With more than one rank the all-reduce stays OUTSIDE the recording (two graphs around one
    # eager NCCL call), so nothing of NCCL's is captured.  Learning rates, shapes and users_per_pass are baked: a step
    # with other shapes gets its own recording (the two most recently used are kept).
    _GRAPH_SLOTS = 64

    def _graph_body_backward(self):
        self.zero_grad()
        return self.forward_backward(self._g_items, self._g_mask)

    def _graph_body_adam(self):
        for _, lr, off, n in self.segments:
            ops.adam_step_dev(self.flat_param[off:off + n], self.flat_grad[off:off + n], self.exp_avg[off:off + n],
                              self.exp_avg_sq[off:off + n], lr, self.betas[0], self.betas[1], self.eps,
                              self.weight_decay, self._g_state.view(torch.float32)[2:4], grad_scale=1.0 / self.world)

    def graph_unsafe_reason(self):
        """why this trainer's step cannot be recorded (None if it can): a recording needs shapes and launch sequences that do
        not depend on the batch's VALUES"""
        if self.buckets:
            return ("graphed steps keep the gradient all-reduce outside the recording: build the trainer with overlap=False")
        for name, m in self.model.named_modules():
            if getattr(m, "unpad", False):
                return "%s.unpad = True executes only the kept tokens: the token count is read back from the batch" % (name or "model")
            if getattr(m, "dedup_items", False):
                return "%s.dedup_items = True encodes the distinct items: their number is read back from the batch" % (name or "model")
        return None

    _MAX_RECORDINGS = 2      # an epoch has two batch shapes: the full batches and the last, shorter one

    def _record(self, sample_items, log_mask):
        reason = self.graph_unsafe_reason()
        if reason is not None:
            raise RuntimeError("train_step_graphed: " + reason)
        dev = sample_items.device
        if getattr(self, "_g_state", None) is None:      # shared by every recording of this trainer
            self._g_state = torch.zeros(2, dtype=torch.int64, device=dev)      # [seed | bc1 f32, bc2_sqrt f32]
            self._g_host = torch.zeros(self._GRAPH_SLOTS, 2, dtype=torch.int64).pin_memory()
            self._g_events = [None] * self._GRAPH_SLOTS
            self._g_replays = 0
        self._g_items, self._g_mask = sample_items.clone(), log_mask.clone()
        Fn.bump_param_epoch()        # whatever built the weight caches last: the recording must contain their rebuild
        Fn.DropoutState.seed_address = self._g_state.data_ptr()
        self._g_counter0, self._g_base_seed = Fn.DropoutState.counter, Fn.DropoutState.seed
        torch.cuda.synchronize(dev)
        try:
            g1 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g1):
                self._g_loss = self._graph_body_backward()
                if self.world <= 1:
                    self._graph_body_adam()
            graphs = [g1]
            if self.world > 1:
                g2 = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g2, pool=g1.pool()):
                    self._graph_body_adam()
                graphs.append(g2)
        finally:
            Fn.DropoutState.seed_address = None
        return {"graphs": graphs, "items": self._g_items, "mask": self._g_mask, "loss": self._g_loss,
                "counter0": self._g_counter0, "base_seed": self._g_base_seed}

    def _select_recording(self, sample_items, log_mask):
        """the recording for these shapes becomes the current one (recorded now if there is none; the least recently used of
        more than _MAX_RECORDINGS is dropped with its memory pool)"""
        shapes = (tuple(sample_items.shape), tuple(log_mask.shape), self.users_per_pass)
        recs = self.__dict__.setdefault("_recordings", {})
        rec = recs.pop(shapes, None)
        if rec is None:
            while len(recs) >= self._MAX_RECORDINGS:
                recs.pop(next(iter(recs)))
            self._graphs = self._g_loss = self._g_items = self._g_mask = None
            rec = self._record(sample_items, log_mask)
        recs[shapes] = rec                                # most recently used last
        self._graphs, self._g_items, self._g_mask, self._g_loss = rec["graphs"], rec["items"], rec["mask"], rec["loss"]
        self._g_counter0, self._g_base_seed, self._g_shapes = rec["counter0"], rec["base_seed"], shapes

    def graph_seed(self, step):
        """the dropout seed replay `step` (= step_count after the step) runs with — an eager step given this seed and the
        recording's first counter (graph_counter0) draws the same masks (tests/test_graph_step_gpu.py)"""
        return Fn.DropoutState.replay_seed(step, self._g_base_seed)

    @property
    def graph_counter0(self):
        return self._g_counter0

    def release_graph(self):
        """drop every recording and its private memory pool"""
        self._recordings = {}
        self._graphs = None
        self._g_loss = self._g_items = self._g_mask = None

    def train_step_graphed(self, sample_items, log_mask):
        """train_step replayed from a recording (CUDA only).  The first call of a trainer runs eagerly (it loads every
        kernel and builds the frozen-weight caches, neither of which may happen inside a recording); the next call
        records; every call after that is: two small copies into the static inputs, 16 bytes of per-step state, one
        graph launch (two around the all-reduce when world > 1).  One recording per batch shape, the two most recently
        used are kept (an epoch alternates between its full batches and the last, shorter one).  Returns the loss tensor
        of the recording (overwritten by its next replay)."""
        if not getattr(self, "_eager_steps", 0):
            return self.train_step(sample_items, log_mask)
        self._select_recording(sample_items, log_mask)
        self._g_items.copy_(sample_items, non_blocking=True)
        self._g_mask.copy_(log_mask, non_blocking=True)
        self.step_count += 1
        slot = self._g_replays % self._GRAPH_SLOTS
        self._g_replays += 1
        if self._g_events[slot] is not None:
            self._g_events[slot].synchronize()       # the copy that last read this pinned slot has run
        bc1, bc2s = ops.adam_bias_corrections(self.betas[0], self.betas[1], self.step_count)
        host = self._g_host[slot]
        host[0] = self.graph_seed(self.step_count)
        host.view(torch.float32)[2] = bc1
        host.view(torch.float32)[3] = bc2s
        self._g_state.copy_(host, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._g_events[slot] = ev
        self._graphs[0].replay()
        if self.world > 1:
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=self.pg)
            self._graphs[1].replay()
        Fn.bump_param_epoch()        # eager consumers (the evaluator) rebuild their weight caches from the new parameters
        return self._g_loss

    def state_dict(self):
        """The optimizer entry of a checkpoint (data_utils/utils.py:109-115 stores `optimizer.state_dict()`), in
        torch.optim.Adam's OWN format — `state` {index: step, exp_avg, exp_avg_sq} per parameter, `param_groups` with the four
        groups of run.py:524-529 in the reference's order (bert, recsys, adapter_bert, adapter_recsys; an empty group stays in
        the list) and consecutive parameter indices — so that a checkpoint written here resumes under the reference's
        `optimizer.load_state_dict` and one written by the reference resumes here.  The extra key `a4r` carries what torch has no
        slot for: the parameter names (checked on load) and the dropout RNG position (the reference checkpoints torch's RNG
        states for the same purpose: a resumed run draws the masks the uninterrupted run would have drawn)."""
        template = torch.optim.Adam([torch.nn.Parameter(torch.zeros(1))], lr=1.0, betas=self.betas, eps=self.eps,
                                    weight_decay=self.weight_decay).state_dict()['param_groups'][0]
        lr_of = {g: lr for g, lr, _, _ in self.segments}
        state, groups, index = {}, [], 0
        by_group = {g: [] for g in self._group_order}
        for (g, _n, p), (_name, off, k) in zip(self._plist, self.names):
            by_group[g].append(index)
            if self.step_count > 0:
                state[index] = {"step": torch.tensor(float(self.step_count)),
                                "exp_avg": self.exp_avg[off:off + k].view(p.shape).clone(),
                                "exp_avg_sq": self.exp_avg_sq[off:off + k].view(p.shape).clone()}
            index += 1
        for g in self._group_order:
            entry = dict(template)
            entry["lr"] = lr_of.get(g, self._lrs[g])
            entry["params"] = by_group[g]
            groups.append(entry)
        return {"state": state, "param_groups": groups,
                "a4r": {"names": list(self.names), "dropout_seed": Fn.DropoutState.seed,
                        "dropout_counter": Fn.DropoutState.counter}}

    def load_state_dict(self, sd):
        """Accepts (a) state_dict() above, (b) the state_dict of the reference's torch.optim.Adam over the same trainable set
        (same four groups; checked by count and shape — torch's format carries no names), (c) the flat format this package
        wrote before (keys step / exp_avg / exp_avg_sq / names)."""
        if "param_groups" not in sd:                                                   # (c)
            if [tuple(x) for x in sd["names"]] != [tuple(x) for x in self.names]:
                raise ValueError("optimizer state was saved for a different set of trainable parameters")
            self.step_count = sd["step"]
            self.exp_avg.copy_(sd["exp_avg"])
            self.exp_avg_sq.copy_(sd["exp_avg_sq"])
            extra = sd
        else:
            extra = sd.get("a4r", {})
            if "names" in extra and [tuple(x) for x in extra["names"]] != [tuple(x) for x in self.names]:
                raise ValueError("optimizer state was saved for a different set of trainable parameters")
            sizes = [len(g["params"]) for g in sd["param_groups"]]
            mine = [sum(1 for g, _, _ in self._plist if g == name) for name in self._group_order]
            if sizes != mine:
                raise ValueError("optimizer state has parameter groups of sizes %r, this trainer's are %r (bert, recsys, "
                                 "adapter_bert, adapter_recsys)" % (sizes, mine))
            order = [i for g in sd["param_groups"] for i in g["params"]]
            steps = set()
            self.exp_avg.zero_()
            self.exp_avg_sq.zero_()
            for idx, (g, _n, p), (_name, off, k) in zip(order, self._plist, self.names):
                st = sd["state"].get(idx)
                if st is None:                     # a parameter that never received a gradient has no state in torch
                    continue
                if tuple(st["exp_avg"].shape) != tuple(p.shape):
                    raise ValueError("optimizer state of parameter %d has shape %r, %s has %r"
                                     % (idx, tuple(st["exp_avg"].shape), _name, tuple(p.shape)))
                self.exp_avg[off:off + k].copy_(st["exp_avg"].reshape(-1))
                self.exp_avg_sq[off:off + k].copy_(st["exp_avg_sq"].reshape(-1))
                steps.add(int(st["step"]))
            if len(steps) > 1:
                raise ValueError("parameters with different step counts (%r): one flat Adam step count cannot represent them"
                                 % sorted(steps))
            self.step_count = steps.pop() if steps else 0
        if "dropout_seed" in extra:
            Fn.DropoutState.seed, Fn.DropoutState.counter = int(extra["dropout_seed"]), int(extra["dropout_counter"])
