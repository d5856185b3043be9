"""Data-parallel gradient exchange for the training-step configuration (second table row; 8f rank 2).

The reference trains on one GPU (train.py:137-139); this is the new multi-GPU piece: one process per GPU, replicas with
local BatchNorm statistics, ONE exchange per step — an all-reduce (sum, then / world) of the flat fp32 gradient
(2 912 707 elements = 11.65 MB for C_NETWORK) — followed by the global-norm clip of config.py:48-49
(`gradient_clip_val` 100.0, algorithm "norm"), which needs the post-reduce norm and therefore no second collective.

Gradients live in a few flat buckets, filled decoder-first (the order backward produces them), so each bucket's
all-reduce can be launched asynchronously while the rest of backward still runs; parameters' `.grad` are views into the
buckets, so no copy in or out.  `torch.distributed` is plumbing: NCCL over NVLink on the GPUs, gloo in the CPU tests.
train_engine.TrainStep's backward kernels write straight into these buffers (`flat=True`: ONE buffer whose 4 MB slices are the buckets, so the
fused clip + Adam-amsgrad kernel sees a single array); the class works on whatever wrote `.grad`.
"""
import torch
import torch.distributed as dist


def forward_stage(name):
    """Position of a C_NETWORK / R_NETWORK parameter in the forward pass (c_network.py:187-226): initial BN, encoder 0..6,
    LSTM, fc, then decoder stage i = {skip_attention 2i / 2i+1, decoder i, decoder_attention 2i / 2i+1}.  (`named_parameters()`
    follows construction order, which is not forward order.)"""
    head, _, rest = name.partition(".")
    idx = int(rest.split(".")[0]) if rest and rest.split(".")[0].isdigit() else 0
    if head == "initial_batchnorm":
        return 0
    if head == "encoder":
        return 1 + idx
    if head == "lstm":
        return 100
    if head == "fc":
        return 101
    if head == "skip_attention":
        return 200 + 10 * (idx // 2)
    if head == "decoder":
        return 200 + 10 * idx + 1
    if head == "decoder_attention":
        return 200 + 10 * (idx // 2) + 2
    return 1000


class GradBuckets:
    def __init__(self, params, bucket_bytes=4 << 20, device=None, stage_of=forward_stage, flat=False):
        """params: iterable of (name, Parameter), e.g. `net.named_parameters()`; buckets are laid out last-stage-first (the
        order backward produces gradients).  Complex parameters are not expected (the reference's parameters are all real)."""
        named = [(n, p) for n, p in params if p.requires_grad]
        if any(p.is_complex() or p.dtype != torch.float32 for _, p in named):
            raise TypeError("GradBuckets: fp32 real parameters only")
        self.order = sorted(named, key=lambda np_: -stage_of(np_[0]))
        device = device or (self.order[0][1].device if self.order else "cpu")
        cap = max(int(bucket_bytes) // 4, 1)
        plan, cur, used = [], [], 0
        for n, p in self.order:
            if cur and used + p.numel() > cap:
                plan.append(cur)
                cur, used = [], 0
            cur.append((n, p))
            used += p.numel()
        if cur:
            plan.append(cur)
        self.buckets, self.slices = [], {}
        # flat=True: ONE buffer holds every bucket back to back (self.flat), so the optimizer kernels see a single array
        self.flat = torch.zeros(sum(p.numel() for _, p in self.order), dtype=torch.float32, device=device) if flat else None
        base = 0
        for b, members in enumerate(plan):
            size = sum(p.numel() for _, p in members)
            flat = self.flat[base:base + size] if self.flat is not None else torch.zeros(size, dtype=torch.float32, device=device)
            base += size
            off = 0
            for n, p in members:
                view = flat[off:off + p.numel()].view_as(p)
                if p.grad is not None:
                    view.copy_(p.grad)
                p.grad = view                                    # autograd / the backward kernels accumulate in place
                self.slices[n] = (b, off, p.numel())
                off += p.numel()
            self.buckets.append(flat)
        self._pending = []

    @property
    def numel(self):
        return sum(b.numel() for b in self.buckets)

    def zero(self):
        for b in self.buckets:
            b.zero_()

    def launch(self, index, group=None):
        """Start the all-reduce of one bucket (call as soon as backward has finished writing it)."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            self._pending.append((index, dist.all_reduce(self.buckets[index], op=dist.ReduceOp.SUM, group=group, async_op=True)))

    def wait(self, group=None):
        """Wait for the launched buckets and reduce any that were not launched; the SUM stays in the buckets (the fused optimizer
        kernel applies 1 / world itself)."""
        world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        launched = {i for i, _ in self._pending}
        for _, w in self._pending:
            w.wait()
        self._pending = []
        if world > 1:
            for i, b in enumerate(self.buckets):
                if i not in launched:
                    dist.all_reduce(b, op=dist.ReduceOp.SUM, group=group)
        return self

    def finish(self, group=None):
        """Wait for the launched buckets, reduce any that were not launched, divide by the world size."""
        world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        launched = {i for i, _ in self._pending}
        for _, w in self._pending:
            w.wait()
        self._pending = []
        if world > 1:
            for i, b in enumerate(self.buckets):
                if i not in launched:
                    dist.all_reduce(b, op=dist.ReduceOp.SUM, group=group)
            for b in self.buckets:
                b.div_(world)
        return self

    def clip_by_global_norm(self, max_norm=100.0, eps=1e-6):
        """torch.nn.utils.clip_grad_norm_ semantics (what Lightning's gradient_clip_algorithm='norm' calls) on the flat
        buckets: coef = min(1, max_norm / (norm + eps)).  Returns the pre-clip norm (0-dim tensor)."""
        total = torch.sqrt(sum((b.double() ** 2).sum() for b in self.buckets)).float()
        coef = torch.clamp(max_norm / (total + eps), max=1.0)
        for b in self.buckets:
            b.mul_(coef)
        return total
