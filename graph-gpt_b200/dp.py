"""Data-parallel training engine: NCCL gradient all-reduce over NVLink + fused AdamW on the flat buffers.

Replaces, for the GraphGPT hot path, the reference's two wrappers:
  * DeepSpeed ZeRO-2 engine  (`deepspeed.initialize`, pretrain_mode.py:271-287; ds_config2_pt_bf16.json:24-27:
    grad reduce-scatter + sharded Adam + param all-gather), and
  * torch DDP + AdamW + clip (opt_utils.py:7-37, training_utils.py:66-86).
The model is 122-417 M parameters, i.e. <= 7 GB of fp32 master + Adam state per GPU, so nothing needs sharding on
180 GB parts: every rank keeps the full state, gradients are summed with ONE all-reduce per contiguous segment
of the flat gradient buffer (head | layer L-1 | ... | layer 0 | embedding), launched from inside backward as soon
as the segment is final so NCCL overlaps the remaining backward kernels, and the 1/world factor is folded into
the AdamW kernel.  The object mimics the DeepSpeed engine surface the reference's loop uses:
`engine(**batch)`, `engine.backward(loss)`, `engine.step()`, `engine.module`, `engine.global_steps`.
"""
import math

import torch
import torch.distributed as dist

from . import ops


class GradReducer:
    """Launches async all-reduces over [start,end) spans of one flat tensor and waits for them later.
    Device-agnostic (NCCL on GPUs, gloo in the CPU tests).

    wire_dtype = torch.bfloat16 exchanges the gradients in bf16 — what the reference's DeepSpeed bf16 configuration does
    (ds_config2_pt_bf16.json: bf16 enabled => gradients are reduced in bf16): each span is cast into a bf16 staging buffer
    (one streaming kernel), all-reduced at half the NVLink bytes, and cast back into the fp32 flat gradient after the
    wait.  The fp32 accumulation of the wgrad GEMMs, the clipping norm and AdamW are unchanged."""

    def __init__(self, flat_grad, group=None, wire_dtype=None, staging=None):
        self.flat = flat_grad
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.pending = []
        self.spans = []
        self.wire_dtype = wire_dtype if wire_dtype not in (None, torch.float32) else None
        self.staging = staging
        if self.wire_dtype is not None and self.staging is None and self.world > 1:
            self.staging = torch.empty(flat_grad.shape, device=flat_grad.device, dtype=self.wire_dtype)

    def reduce_span(self, start, end):
        self.spans.append((start, end))
        if self.world == 1:
            return
        if self.wire_dtype is None:
            buf = self.flat[start:end]
        else:
            buf = self.staging[start:end]
            if self.flat.is_cuda:
                ops.cast_f32_bf16(self.flat[start:end], buf)
            else:                               # CPU unit tests (gloo) only
                buf.copy_(self.flat[start:end])
        self.pending.append((dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group, async_op=True), start, end))

    def wait(self):
        for w, start, end in self.pending:
            w.wait()
            if self.wire_dtype is not None:
                if self.flat.is_cuda:
                    ops.cast_bf16_f32(self.staging[start:end], self.flat[start:end])
                else:
                    self.flat[start:end].copy_(self.staging[start:end])
        self.pending = []
        spans, self.spans = self.spans, []
        return spans


def warmup_decay_lr(step, *, max_lr, min_lr=0.0, warmup_steps=0, total_steps=1, warmup_type="log"):
    """DeepSpeed WarmupDecayLR (deepspeed==0.15.4, ds_config2_pt_bf16.json:15-23; the reference's helper
    loss_utils.py:186-201 sets min/max lr and the two step counts and leaves `warmup_type` at DeepSpeed's default "log"),
    as a function of DeepSpeed's `last_batch_iteration` (= `step`):
      step <  0            : min_lr   (get_lr() before the scheduler has stepped)
      step <  warmup_steps : gamma = log(step + 1) / log(warmup_steps)     ("log")   |   step / warmup_steps   ("linear")
      step >= warmup_steps : gamma = max(0, (total_steps - step) / max(1, total_steps - warmup_steps))
      lr = min_lr + (max_lr - min_lr) * gamma
    GraphGPTEngine evaluates it the way the DeepSpeed engine does — the scheduler steps AFTER the optimizer, starting from
    last_batch_iteration = -1 — so the k-th parameter update runs at warmup_decay_lr(k - 2): the first two updates use
    min_lr (see GraphGPTEngine.schedule_iteration)."""
    if step < 0:
        return min_lr
    if warmup_steps > 0 and step < warmup_steps:
        if warmup_type == "log":
            gamma = math.log(step + 1) / math.log(warmup_steps) if warmup_steps > 1 else 1.0
        elif warmup_type == "linear":
            gamma = step / warmup_steps
        else:
            raise ValueError(f"warmup_type={warmup_type!r}")
    else:
        gamma = max(0.0, (total_steps - step) / max(1, total_steps - warmup_steps))
    return min_lr + (max_lr - min_lr) * gamma


def write_model_pt(module, ckp_dir):
    """`<ckp_dir>/model.pt` = the plain fp32 state dict on the CPU.  This is the file the reference's fine-tuning / eval
    pipeline opens FIRST when it loads a pre-trained checkpoint (`load_from_ckp_with_try`, loader_utils.py:176-220:
    torch.load(os.path.join(ckp, "model.pt")), falling back to DeepSpeed's zero_to_fp32 only on failure), so a checkpoint
    written by GraphGPTEngine.save_checkpoint is consumed by the unchanged pipeline."""
    import os
    sd = {k: v.detach().to("cpu", copy=True) for k, v in module.state_dict().items()}
    path = os.path.join(ckp_dir, "model.pt")
    torch.save(sd, path)
    return path


class GraphGPTEngine:
    def __init__(self, model, *, lr=3e-4, betas=(0.9, 0.95), eps=1e-6, weight_decay=0.1, max_grad_norm=1.0,
                 lr_schedule=None, process_group=None, overlap_comm=True, gradient_accumulation_steps=1,
                 grad_reduce_dtype=torch.bfloat16):
        self.module = model
        # DeepSpeed semantics (ds_config["gradient_accumulation_steps"], conf_utils.py:62-65): backward() scales the loss
        # by 1/gas and accumulates; the gradient exchange and the optimizer run on every gas-th step() only
        self.gas = max(1, int(gradient_accumulation_steps))
        self.micro_steps = 0
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.max_grad_norm = max_grad_norm
        self.lr_schedule = lr_schedule
        self.global_steps = 0
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        self.overlap_comm = overlap_comm
        hot = model.hot                       # builds / validates the flat buffers (raises without CUDA)
        self.flat = hot.flat
        n = self.flat.numel
        dev = self.flat.flat.device
        self.exp_avg = torch.zeros((n,), device=dev, dtype=torch.float32)
        self.exp_avg_sq = torch.zeros((n,), device=dev, dtype=torch.float32)
        self.gnorm_sq = torch.zeros((1,), device=dev, dtype=torch.float64)
        flat_ids = {id(p) for _, p in self.flat.order}
        self.extra = [p for p in model.parameters() if id(p) not in flat_ids and p.requires_grad]
        self.extra_opt = (torch.optim.AdamW(self.extra, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
                          if self.extra else None)
        # gradients cross NVLink in bf16 by default, as under the reference's DeepSpeed bf16 configuration
        # (grad_reduce_dtype=torch.float32 keeps the exchange in fp32)
        self.grad_reduce_dtype = grad_reduce_dtype
        self._wire = (torch.empty((n,), device=dev, dtype=grad_reduce_dtype)
                      if (self.world > 1 and grad_reduce_dtype not in (None, torch.float32)) else None)
        self.reducer = GradReducer(self.flat.flat_grad, process_group, grad_reduce_dtype, self._wire)
        # 4-byte collective that aligns the ranks' compute streams at the start of backward (see backward())
        self._align = torch.zeros((1,), device=dev, dtype=torch.float32) if self.world > 1 else None
        self.align_ranks = True
        if self.world > 1:
            self._broadcast_params()

    def _broadcast_params(self):
        dist.broadcast(self.flat.flat, src=0, group=self.group)
        for p in self.extra:
            dist.broadcast(p.data, src=0, group=self.group)
        self.flat.invalidate_bf16()           # the broadcast wrote through raw storage: re-cast the bf16 copy
        self.flat.ensure()

    # ---- DeepSpeed-engine-like surface -----------------------------------------------------------
    def __call__(self, *a, **k):
        return self.module(*a, **k)

    def train(self, mode=True):
        self.module.train(mode)
        return self

    def eval(self):
        self.module.eval()
        return self

    def parameters(self):
        return self.module.parameters()

    def is_gradient_accumulation_boundary(self):
        return (self.micro_steps + 1) % self.gas == 0

    def backward(self, loss):
        hot = self.module._hot
        boundary = self.is_gradient_accumulation_boundary()
        if self.gas > 1:
            loss = loss / self.gas
        if not boundary:                      # accumulate locally; nothing is exchanged until the boundary micro-step
            hot.grad_ready_hook = None
            loss.backward()
            return
        self.reducer = GradReducer(self.flat.flat_grad, self.group, self.grad_reduce_dtype, self._wire)
        if self.world > 1 and self.overlap_comm:
            if self.align_ranks:
                # The ranks reach backward up to ~1 ms apart (different batches => different head sizes, host jitter).  An
                # all-reduce whose peer is late SPINS on its 16-24 SMs until the peer arrives, and the persistent GEMMs
                # running beside it lose those SMs (kernel timeline, tools/dp_trace.py: the first two gradient all-reduces
                # took 1.07 / 0.77 ms instead of 0.09 ms and the backward GEMMs 1.25 ms longer).  A 4-byte all-reduce —
                # one NCCL channel, one SM — on which the compute stream waits absorbs the skew here instead; the step
                # time is the slowest rank's either way, and every later collective finds its peers ready.
                dist.all_reduce(self._align, group=self.group, async_op=True).wait()
            hot.grad_ready_hook = lambda first, last: self.reducer.reduce_span(*self.flat.span(first, last))
        else:
            hot.grad_ready_hook = None
        try:
            loss.backward()
        finally:
            hot.grad_ready_hook = None
        if self.world > 1 and not self.overlap_comm:
            self.reducer.reduce_span(0, self.flat.numel)

    @staticmethod
    def schedule_iteration(updates_done):
        """DeepSpeed's `last_batch_iteration` in effect during parameter update number `updates_done + 1`: the scheduler
        starts at -1 and steps once AFTER every optimizer step (DeepSpeedEngine._take_model_step), so update 1 sees -1
        (WarmupLR.get_lr() -> min_lr), update 2 sees 0, update k sees k - 2."""
        return updates_done - 1

    def step(self):
        self.micro_steps += 1
        if self.micro_steps % self.gas != 0:
            return                            # not a boundary: gradients keep accumulating in the flat buffer
        sched_it = self.schedule_iteration(self.global_steps)
        self.global_steps += 1
        self.module._hot.check_device_errors(block=False)
        fp = self.flat
        self.reducer.wait()
        inv_world = 1.0 / self.world
        if self.world > 1:
            for p in self.extra:
                if p.grad is not None:
                    dist.all_reduce(p.grad, group=self.group)
                    p.grad.mul_(inv_world)
        lr = self.lr_schedule(sched_it) if self.lr_schedule is not None else self.lr
        # frozen parameters (requires_grad = False, e.g. freeze_llama_layers, modules_utils.py:45-54) are skipped entirely —
        # no update, no weight decay, not in the clipping norm — exactly as torch / DeepSpeed optimizers skip them
        spans = self._trainable_spans()
        gn = None
        if self.max_grad_norm and self.max_grad_norm > 0:
            self.gnorm_sq.zero_()
            for a, b in spans:
                ops.sumsq(fp.flat_grad[a:b], self.gnorm_sq)
            for p in self.extra:
                if p.grad is not None:       # extra grads are already averaged; flat ones are scaled in-kernel
                    self.gnorm_sq += (p.grad.double().pow(2).sum() * (self.world ** 2))
            gn = self.gnorm_sq
        for a, b in spans:
            ops.adamw(fp.flat[a:b], fp.flat_bf16[a:b], fp.flat_grad[a:b], self.exp_avg[a:b], self.exp_avg_sq[a:b], lr=lr,
                      betas=self.betas, eps=self.eps, weight_decay=self.weight_decay, step=self.global_steps, gnorm_sq=gn,
                      max_norm=self.max_grad_norm or 0.0, grad_scale=inv_world)
        fp.mark_bf16_fresh()
        if self.extra_opt is not None:
            if gn is not None:
                coef = torch.clamp(self.max_grad_norm / (gn.sqrt().float() * inv_world + 1e-6), max=1.0)
                for p in self.extra:
                    if p.grad is not None:
                        p.grad.mul_(coef)
            for grp in self.extra_opt.param_groups:
                grp["lr"] = lr
            self.extra_opt.step()
        self.zero_grad()

    def _trainable_spans(self):
        """Coalesced [start, end) element ranges of the flat buffers that belong to trainable parameters (cached per
        pattern of requires_grad flags)."""
        key = tuple(p.requires_grad for _, p in self.flat.order)
        if key != getattr(self, "_spans_key", None):
            spans = []
            for name, p in self.flat.order:
                if not p.requires_grad:
                    continue
                a, b = self.flat.span(name, name)
                if spans and spans[-1][1] == a:
                    spans[-1][1] = b
                else:
                    spans.append([a, b])
            self._spans_key, self._spans = key, [(a, b) for a, b in spans]
        return self._spans

    def zero_grad(self):
        for _, p in self.flat.order:
            p.grad = None                    # the flat buffer is zeroed lazily by the next backward
        for p in self.extra:
            p.grad = None

    def grad_norm(self):
        """Global gradient norm of the last step (after averaging), as a python float (syncs)."""
        return math.sqrt(float(self.gnorm_sq.item())) / self.world

    @property
    def device(self):
        return self.flat.flat.device

    @staticmethod
    def freeze_gc():
        """Call once after model / data-loader set-up and a few warm-up steps.  The launch thread runs only a few
        milliseconds ahead of the device (one step = ~300 kernel launches + one 8-byte D2H read), so a full Python
        garbage collection over the long-lived object graph (tens of milliseconds) drains the queue and idles the GPU:
        measured as one 125 ms step in every ~10 steps of 65 ms.  gc.freeze() moves everything allocated so far into the
        permanent generation; young-object collection keeps running."""
        import gc
        gc.collect()
        gc.freeze()

    # ---- DeepSpeed-style checkpoint directory (misc_utils.py:67-98: `model.save_checkpoint(model_dir)` is called by ALL
    # ranks, files land in model_dir/global_step<N>/ and model_dir/latest names the tag) ---------------------------
    def save_checkpoint(self, save_dir, tag=None, client_state=None):
        import os
        tag = tag or f"global_step{self.global_steps}"
        rank = dist.get_rank(self.group) if self.world > 1 else 0
        if rank == 0:                         # replicated state: one writer
            os.makedirs(os.path.join(save_dir, tag), exist_ok=True)
            sd = self.state_dict()
            sd["client_state"] = client_state or {}
            sd["micro_steps"] = self.micro_steps
            torch.save(sd, os.path.join(save_dir, tag, "ggpt_engine_states.pt"))
            write_model_pt(self.module, save_dir)
            with open(os.path.join(save_dir, "latest"), "w") as f:
                f.write(tag)
        if self.world > 1:
            dist.barrier(self.group)
        return True

    def load_checkpoint(self, load_dir, tag=None, load_optimizer_states=True):
        """Returns (path, client_state) like DeepSpeedEngine.load_checkpoint; (None, None) when nothing is there."""
        import os
        if tag is None:
            latest = os.path.join(load_dir, "latest")
            if not os.path.exists(latest):
                return None, None
            tag = open(latest).read().strip()
        path = os.path.join(load_dir, tag, "ggpt_engine_states.pt")
        if not os.path.exists(path):
            return None, None
        sd = torch.load(path, map_location=self.device)
        if load_optimizer_states:
            self.load_state_dict(sd)
            self.micro_steps = int(sd.get("micro_steps", self.global_steps * self.gas))
        else:
            self.module.load_state_dict(sd["model"])
            self.flat.ensure()
        return path, sd.get("client_state", {})

    # ---- checkpoint surface (DDP-style files; misc_utils.py:105-121) -------------------------------
    def state_dict(self):
        return {"model": self.module.state_dict(), "exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq,
                "global_steps": self.global_steps,
                "extra_opt": self.extra_opt.state_dict() if self.extra_opt is not None else None}

    def load_state_dict(self, sd):
        self.module.load_state_dict(sd["model"])
        self.flat.ensure()
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self.global_steps = int(sd["global_steps"])
        if self.extra_opt is not None and sd.get("extra_opt") is not None:
            self.extra_opt.load_state_dict(sd["extra_opt"])
