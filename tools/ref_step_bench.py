"""BASELINE.md B3, "the real bar": the reference-style train step on the same B200 -- the reference's
own formulation of the step (unfused modules, per-scene / per-box python loops, cuDNN / ATen
Conv-BN-ReLU, torch losses) running on the reference's OWN CUDA kernels compiled unmodified for
sm_100a (oracle/_ref: FPS, ball query, gather / group, three_nn / interpolate, points_in_boxes,
sort_vertices) -- timed next to this repo's step on the same inputs.

    python tools/ref_step_bench.py [--workload pretrain|stress|mean_teacher] [--steps 5]

Prints one JSON line: ms/step and scenes/s of both, eager launches (no CUDA graph on either side),
so the ratio isolates kernels + formulation; bench.py's graph-captured number is quoted beside it.
The twin classes are oracle/detectors_ref.py with their op backend switched from the CPU
restatement to oracle/ref_cuda.py (same function names)."""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def use_reference_cuda_kernels():
    from oracle import modules, nesie_head_ref, ref_cuda, side_pooling_ref
    assert ref_cuda.available() and ref_cuda.pib_available() and ref_cuda.sortv_available(), \
        "oracle/_ref libraries missing (built by __graft_entry__.build() where /root/reference exists)"
    for mod in (modules, nesie_head_ref, side_pooling_ref):
        mod.cpu = ref_cuda
    nesie_head_ref.NesieHeadOracle._k_sort_vertices = staticmethod(ref_cuda.sort_vertices)


def time_steps(wl, model, inp, static, steps, warmup):
    opt = torch.optim.AdamW(model.parameters(), lr=0.008, weight_decay=0.01)
    ts = []
    for it in range(warmup + steps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        loss = bench.step_loss(wl, model, inp, static=static)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 10.0)
        opt.step()
        if wl == "mean_teacher":
            model.after_train_iter(10 + it)
        torch.cuda.synchronize()
        if it >= warmup:
            ts.append(time.perf_counter() - t0)
    ts.sort()
    return ts[len(ts) // 2], float(loss)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="pretrain", choices=["pretrain", "stress", "mean_teacher"])
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--scenes", type=int, default=None)
    args = ap.parse_args()
    wl = args.workload
    S = args.scenes or bench.WORKLOADS[wl]["scenes"]
    dev = torch.device("cuda:0")
    torch.backends.cuda.matmul.allow_tf32 = False      # fp32 arithmetic on both sides
    torch.backends.cudnn.allow_tf32 = False
    inp = {k: v.to(dev) for k, v in bench.make_host_batch(wl, 0, S).items()}
    static = dict(sup_index=torch.arange(S // 2, device=dev), unsup_index=torch.arange(S // 2, S, device=dev))
    out = {"workload": bench.WORKLOADS[wl]["name"], "scenes_per_step": S, "launch": "eager"}
    for side in ("nesie_b200", "reference_style"):
        torch.manual_seed(0)
        if side == "reference_style":
            use_reference_cuda_kernels()
        model = bench.build_model(wl, oracle=(side == "reference_style")).to(dev)
        if wl == "mean_teacher":
            model.init_teacher()
        sec, loss = time_steps(wl, model, inp, static, args.steps, args.warmup)
        out[side] = {"ms_per_step": sec * 1e3, "scenes_per_s": S / sec, "loss": loss}
        del model
        torch.cuda.empty_cache()
    out["speedup"] = out["reference_style"]["ms_per_step"] / out["nesie_b200"]["ms_per_step"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
