"""Independent sub-graphs of a step on forked CUDA streams.

The captured train step is dominated by a few large kernels, but between them sit hundreds of small
ones (per-term loss arithmetic, the seven MiniPointNet + head chains of SidePooling ...).  Chains that
do not depend on each other are issued round-robin on a few streams forked from the caller's stream
and joined before their results are used: inside a CUDA graph they become parallel branches, and
autograd replays the same concurrency in the backward pass (a node's backward runs on its forward
stream).  Rules kept here:
  * every branch stream is re-forked from the caller's stream before use (this also orders any reuse of
    the allocator's blocks behind their last reader);
  * tensors made on the caller's stream and read inside a branch (and therefore by the branch's backward,
    which may run while the caller's stream frees and reallocates) are `record_stream`-ed;
  * one set of branch streams per (device, calling stream, width).
"""
import torch

_POOL = {}


def run_branches(funcs, like, shared=None, width=3):
    """funcs: callables without arguments; like: a tensor that tells the device; shared: tensors (or one
    list per branch) made on the current stream that the branches read.  Returns [f() for f in funcs]."""
    n = len(funcs)
    if not like.is_cuda or width <= 1 or n <= 1:
        return [f() for f in funcs]
    dev = like.device
    main = torch.cuda.current_stream(dev)
    key = (dev.index, main.cuda_stream, width)
    if key not in _POOL:
        _POOL[key] = [torch.cuda.Stream(device=dev) for _ in range(width)]
    streams = _POOL[key]
    for st in streams:
        st.wait_stream(main)
    per_branch = shared is not None and len(shared) == n and all(isinstance(s, (list, tuple)) for s in shared)
    outs = []
    for i, f in enumerate(funcs):
        st = streams[i % width]
        for t in ((shared[i] if per_branch else shared) or ()):
            if torch.is_tensor(t) and t.is_cuda:
                t.record_stream(st)
        with torch.cuda.stream(st):
            outs.append(f())
    for st in streams:
        main.wait_stream(st)
    return outs
