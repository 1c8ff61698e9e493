"""Data-parallel gradient exchange of the train step.

The reference wraps the detector in mmcv's MMDistributedDataParallel (built inside
mmdet.apis.train_detector, reached from mmdet3d/apis/train.py:27-34; `broadcast_buffers=False` as in
test.py:186-189): scenes shard by batch, every rank holds a full replica, the only collective of the
path is the gradient all-reduce (~12.4 MB fp32).  Here:

  * every parameter's .grad is a VIEW into one flat fp32 buffer, laid out bucket by bucket in the
    order the backward pass produces gradients (head first, SA1 last), so a bucket is one contiguous
    slice and one NCCL all-reduce -- no gradient copies, no per-tensor launches;
  * a post-accumulate hook per parameter counts the bucket's ready gradients; the last one launches
    the bucket's all-reduce on a communication stream, so the head / FP buckets are reduced while
    the backward pass still runs through the SA levels; `finish()` joins the stream;
  * gradients are not accumulated tensor by tensor: before backward every .grad is None, so autograd
    hands each parameter its gradient tensor as produced (no `grad += g` launch per parameter, ~190 of
    them); when a bucket is complete its gradients are copied into the flat slice by ONE multi-tensor
    copy and the .grad attributes are pointed back at the views for clipping and the optimizer;
  * everything is stream-ordered (no host sync), so a step that calls backward() + finish() can be
    captured in a CUDA graph together with the optimizer;
  * averaging: ReduceOp.AVG on NCCL, SUM then scale on backends without it (gloo in the CPU tests).

BatchNorm statistics stay per-rank (plain BN2d/BN1d in the configs, not SyncBN) and buffers are not
broadcast after construction, as in the reference.  The teacher EMA needs no communication: every
rank applies the same update to identical weights.
"""
import torch
import torch.distributed as dist


def default_buckets(model, bucket_bytes=4 << 20):
    """Parameters in REVERSE registration order (the order their gradients become ready, to a good
    approximation: modules are registered in forward order), cut into buckets of ~bucket_bytes."""
    params = [p for p in model.parameters() if p.requires_grad]
    buckets, cur, size = [], [], 0
    for p in reversed(params):
        cur.append(p)
        size += p.numel() * 4
        if size >= bucket_bytes:
            buckets.append(cur)
            cur, size = [], 0
    if cur:
        buckets.append(cur)
    return buckets


class FlatGradDDP:

    def __init__(self, model, process_group=None, bucket_bytes=4 << 20, buckets=None,
                 broadcast_parameters=True, overlap=True, flatten_parameters=False):
        """flatten_parameters: also re-home every parameter into ONE flat buffer with the layout of the
        gradient buffer (`flat_params`; the modules keep working on views).  AdamW, gradient clipping
        and the teacher EMA are elementwise over all parameters, so with `flat_parameter()` they
        become one launch each instead of a multi-tensor sweep over ~190 tensors."""
        self.model = model
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.buckets = buckets if buckets is not None else default_buckets(model, bucket_bytes)
        self.params = [p for b in self.buckets for p in b]
        assert len({id(p) for p in self.params}) == len(self.params), "a parameter appears in two buckets"
        dev = self.params[0].device
        assert all(p.dtype == torch.float32 and p.device == dev for p in self.params)
        self.device = dev
        pad = lambda n: (n + 63) // 64 * 64      # noqa: E731  every gradient starts 256-byte aligned
        self.flat = torch.zeros(sum(pad(p.numel()) for p in self.params), dtype=torch.float32, device=dev)
        self.slices, self._bucket_of, self._view, off = [], {}, {}, 0
        for bi, bucket in enumerate(self.buckets):
            start = off
            for p in bucket:
                self._view[id(p)] = self.flat[off:off + p.numel()].view_as(p)
                p.grad = self._view[id(p)]
                self._bucket_of[id(p)] = bi
                off += pad(p.numel())
            self.slices.append(self.flat[start:off])
        self.flat_params = None
        if flatten_parameters:
            self.flat_params = torch.zeros_like(self.flat)
            for p in self.params:
                view = self._view[id(p)]
                off = view.storage_offset() - self.flat.storage_offset()
                dst = self.flat_params[off:off + p.numel()].view_as(p)
                dst.copy_(p.data)
                p.data = dst
        self._ready = [[] for _ in self.buckets]
        self._events = [[] for _ in self.buckets]
        self._pending = [len(b) for b in self.buckets]
        self._launched = [False] * len(self.buckets)
        self.overlap = overlap and dev.type == "cuda" and self.world > 1
        self.comm_stream = torch.cuda.Stream(device=dev) if self.overlap else None
        backend = dist.get_backend(process_group) if self.world > 1 else None
        self._avg = backend == "nccl"
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]
        if self.world > 1 and broadcast_parameters:
            self.broadcast()

    # ---- flat optimizer view -----------------------------------------------------------------------
    def flat_parameter(self):
        """One nn.Parameter over `flat_params` whose .grad is the flat gradient buffer: hand it to the
        optimizer instead of the ~190 parameter tensors (same elementwise update; the padding between
        tensors stays 0: zero value, zero gradient)."""
        assert self.flat_params is not None, "construct with flatten_parameters=True"
        fp = torch.nn.Parameter(self.flat_params, requires_grad=True)
        fp.grad = self.flat
        return fp

    def offsets(self):
        """{id(parameter): offset in the flat buffers} (for TeacherEMA(flat=...))."""
        return {id(p): self._view[id(p)].storage_offset() - self.flat.storage_offset() for p in self.params}

    def clip_grad_norm_(self, max_norm):
        """torch.nn.utils.clip_grad_norm_ over the flat gradient buffer (call after finish()): the
        2-norm of all gradients is the 2-norm of the buffer."""
        total = torch.linalg.vector_norm(self.flat)
        self.flat.mul_(torch.clamp(max_norm / (total + 1e-6), max=1.0))
        return total

    # ---- replica consistency --------------------------------------------------------------------
    def broadcast(self, src=0):
        """Identical replicas to start from (parameters and buffers, once)."""
        for t in list(self.model.parameters()) + list(self.model.buffers()):
            dist.broadcast(t.data, src, group=self.group)

    # ---- per step ---------------------------------------------------------------------------------
    def zero_grad(self):
        """Clear the flat buffer, detach every .grad (autograd then assigns instead of adding) and
        re-arm the buckets."""
        self.flat.zero_()
        for p in self.params:
            p.grad = None
        self._pending = [len(b) for b in self.buckets]
        self._launched = [False] * len(self.buckets)
        self._ready = [[] for _ in self.buckets]
        self._events = [[] for _ in self.buckets]

    def _gather(self, bi):
        """The bucket's fresh gradient tensors -> their views of the flat buffer (one multi-tensor
        copy); .grad of every parameter of the bucket becomes its view (zeros where no gradient came)."""
        ready = self._ready[bi]
        for ev in self._events[bi]:
            torch.cuda.current_stream(self.device).wait_event(ev)
        self._events[bi] = []
        if ready:
            torch._foreach_copy_([self._view[id(p)] for p in ready], [p.grad for p in ready])
        for p in self.buckets[bi]:
            p.grad = self._view[id(p)]
        self._ready[bi] = []

    def _reduce(self, bi):
        buf = self.slices[bi]
        if self._avg:
            dist.all_reduce(buf, op=dist.ReduceOp.AVG, group=self.group)
        else:
            dist.all_reduce(buf, group=self.group)
            buf.div_(self.world)
        self._launched[bi] = True

    def _on_grad(self, p):
        bi = self._bucket_of[id(p)]
        if p.grad is not self._view[id(p)]:
            self._ready[bi].append(p)
            if self.device.type == "cuda":
                # the gradient may have been produced on a forked stream (autograd runs a node's
                # backward on its forward stream): the gather must wait for it
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream(self.device))
                self._events[bi].append(ev)
        self._pending[bi] -= 1
        if self._pending[bi] != 0:
            return
        self._flush(bi)

    def _flush(self, bi):
        """Bucket complete: gather on the producing stream (the fresh gradient tensors are released
        right afterwards), then exchange on the communication stream."""
        self._gather(bi)
        if self.world == 1:
            self._launched[bi] = True
        elif self.overlap:
            self.comm_stream.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(self.comm_stream):
                self._reduce(bi)
        else:
            self._reduce(bi)

    def finish(self):
        """Call after backward(): reduces the buckets whose last gradient never arrived (parameters
        without a gradient this step) and makes the current stream wait for the exchange."""
        for bi, done in enumerate(self._launched):
            if not done:
                self._flush(bi)
        if self.overlap:
            torch.cuda.current_stream(self.device).wait_stream(self.comm_stream)

    def remove_hooks(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []
