#!/usr/bin/env python
"""bench.py -- Nesie-VoteNet train scenes/s on synthetic 40k-point ScanNet-shaped scenes.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workloads (BASELINE.json configs):
  pretrain      configs[2] (default): VoteNet pretrain step = PointNet2SASSG -> NesieHead.forward
                (vote module, vote aggregation SA, prediction convs, side2box, random box jitter,
                SidePooling quality head) -> NesieHead.loss (reference target assignment, all 8 loss
                terms) -> backward -> gradient all-reduce -> clip -> AdamW; 8 scenes per GPU.
  mean_teacher  configs[3]: student forward on 16 scenes (8 labeled + 8 unlabeled), teacher forward on
                the same 16 under swapped EMA weights, device-side pseudo-label filter + box transform,
                supervised + unsupervised losses, backward, all-reduce, AdamW, EMA update.
  stress        configs[4]: SAQE-shaped stress, 100k-point scenes, 4096 SA1 centres, 16 scenes per GPU,
                SAQE uncertainty weighting.
  pretrain_conv round-1 harness (quality scores from a 1x1 conv instead of SidePooling), kept as a
                named variant for comparison.
Scenes shard across ranks (nesie_b200.ddp.FlatGradDDP: bucketed NCCL all-reduce overlapped with the
backward pass) -> weak scaling.  Prints ONE JSON line (rank 0).

  value  : scenes/s with the batches already resident in HBM (CUDA events, max over ranks)
  e2e    : scenes/s through the same public API with the batches in pinned host memory: every step
           copies its inputs host->device and reads the loss back
  roofline: the hand-written kernel family with the largest summed time in the step, STEP-WEIGHTED:
           sum of algorithmic bytes of its launches / sum of their durations, every distinct launch
           shape timed live with CUDA events; the best single shape is reported beside it
  cpu_baseline: the same step on the host cores with the oracle port of the reference kernels (the
           reference has no CPU path of its own), median of 3 steps, plus the 1-scene forward of
           BASELINE configs[0]
`--impl reference` times that CPU port alone (all host threads).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "pretrain": dict(scenes=8, points=40000, gt_pad=16,
                     name="votenet_pretrain_fwd_bwd_adamw_b8_per_gpu_40kpts_18cls+side_pooling"),
    "pretrain_conv": dict(scenes=8, points=40000, gt_pad=16,
                          name="votenet_pretrain_fwd_bwd_adamw_b8_per_gpu_40kpts_18cls"),
    "mean_teacher": dict(scenes=16, points=40000, gt_pad=16,
                         name="nesie_mean_teacher_step_8lb_8ulb_per_gpu_40kpts_18cls"),
    "stress": dict(scenes=16, points=100000, gt_pad=16,
                   name="saqe_stress_fwd_bwd_adamw_b16_per_gpu_100kpts_4096seeds"),
}
# dram__bytes_read.sum + dram__bytes_write.sum of one launch of gemm_nt_tma_kernel at its largest
# HBM-bound shape (SA1 layer 3, 1048576 x 64 -> 128, 805.3 MB algorithmic + 8.4 MB of pooled extrema):
# 268.6 MB read + 521.2 MB written in the end-of-round ncu --set full capture (profiles/r02_ncu_notes.md;
# 749.8 MB in round 1's): no re-reads, the tail of the output is still in L2 when the kernel ends
GEMM_DRAM_TRAFFIC = 789.7e6


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return json.load(open(path)), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (profiling recipe's line)."""

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        clocks, reasons, mx = [], set(), None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                clocks.append(float(r[0]))
                mx = float(r[1])
            except (ValueError, IndexError):
                continue
            for nm, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        clocks.sort()
        return {"sm_mhz": clocks[len(clocks) // 2] if clocks else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(clocks)}


# ------------------------------------------------------------------------------------------------
# models and synthetic batches
# ------------------------------------------------------------------------------------------------
def mean_size_file():
    path = os.path.join(tempfile.mkdtemp(), "scannet_means.npz")
    np.savez(path, np.ones((18, 3), dtype=np.float32))
    return path


def build_model(workload, oracle=False):
    """The detector of a workload: product classes, or their CPU twins (oracle=True)."""
    from nesie_b200 import detectors as D
    if oracle:
        from oracle import detectors_ref as R
    if workload == "pretrain_conv":
        if oracle:
            from oracle.votenet_ref import VoteNetOracle
            return VoteNetOracle(quality_head="conv")
        from nesie_b200.votenet import VoteNetHarness
        return VoteNetHarness(quality_head="conv")
    head = D.nesie_head_cfg(mean_size_arr_path=mean_size_file())
    if workload == "pretrain":
        return (R.VoteNetRef if oracle else D.VoteNet)(bbox_head=head)
    if workload == "mean_teacher":
        return (R.VoteNetNesieRef if oracle else D.VoteNetNesie)(bbox_head=head, n_lb=120, n_ulb=1081)
    if workload == "stress":
        head["uncertainty"] = "saqe"
        backbone = dict(in_channels=4, num_points=(4096, 1024, 512, 256))
        return (R.VoteNetRef if oracle else D.VoteNet)(backbone=backbone, bbox_head=head)
    raise ValueError(workload)


def make_host_batch(workload, seed0, scenes=None):
    """One batch as a flat dict of CPU tensors with static shapes (pinned by the caller)."""
    from nesie_b200 import targets as T
    from nesie_b200.synthetic import make_batch
    cfg = WORKLOADS[workload]
    S = scenes or cfg["scenes"]
    cpu = torch.device("cpu")
    if workload == "pretrain_conv":
        from nesie_b200.votenet import VoteNetHarness
        pts, gb, gl = make_batch(S, cfg["points"], seed0=seed0)
        b, l, v = VoteNetHarness._pad_gt(gb, gl, cpu, pad_to=cfg["gt_pad"])
        return dict(pts=pts, gt_boxes=b, gt_labels=l, gt_valid=v)
    pts, gb, gl = make_batch(S, cfg["points"], seed0=seed0, origin="bottom")
    if workload in ("pretrain", "stress"):
        b, l, v = T.pad_gt(gb, gl, cpu, pad_to=cfg["gt_pad"])
        return dict(pts=pts, gt_boxes=b, gt_labels=l, gt_valid=v)
    # mean teacher: scenes [0, S/2) labeled, [S/2, S) unlabeled; student / teacher views of each
    from nesie_b200.detectors import BoxAug, transform_boxes
    g = torch.Generator().manual_seed(seed0)
    aug_s, aug_t = BoxAug.random(S, cpu, g), BoxAug.random(S, cpu, g)
    nl = S // 2
    b, l, v = T.pad_gt(gb[:nl], gl[:nl], cpu, pad_to=cfg["gt_pad"])
    b = transform_boxes(b, aug_s.index(torch.arange(nl))) * v.unsqueeze(-1)     # GT in the student frame
    pts_all = torch.cat([aug_s.apply_points(pts), aug_t.apply_points(pts)], dim=0)   # (2S, N, 4)
    out = dict(pts=pts_all, gt_boxes=b, gt_labels=l, gt_valid=v,
               ulb_pos=(torch.randperm(1081, generator=g)[:S - nl]).long())
    for tag, a in (("s", aug_s), ("t", aug_t)):
        out.update({f"aug_{tag}_hf": a.hf, f"aug_{tag}_vf": a.vf, f"aug_{tag}_rot": a.rot,
                    f"aug_{tag}_scale": a.scale, f"aug_{tag}_trans": a.trans})
    return out


def step_loss(workload, model, inp, fps=None, hook=None, static=None):
    """Forward + losses of one step on the tensors of a batch dict -> scalar loss."""
    kw = {}
    if fps is not None:
        kw = dict(fps_indices=fps, after_level=hook)
    if workload == "pretrain_conv":
        return model.train_step_loss_padded(inp["pts"], inp["gt_boxes"], inp["gt_labels"], inp["gt_valid"],
                                            **kw)[0]
    if workload in ("pretrain", "stress"):
        return sum(model.forward_train_padded(inp["pts"], inp["gt_boxes"], inp["gt_labels"],
                                              inp["gt_valid"], **kw).values())
    from nesie_b200.detectors import BoxAug
    S = inp["pts"].shape[0] // 2
    aug = {t: BoxAug(*(inp[f"aug_{t}_{k}"] for k in ("hf", "vf", "rot", "scale", "trans"))) for t in "st"}
    skw = tkw = None
    if fps is not None:
        skw = dict(fps_indices=[f[:S] for f in fps], after_level=hook)
        tkw = dict(fps_indices=[f[S:] for f in fps])
    losses = model.forward_train_padded(
        inp["pts"][:S], inp["pts"][S:], inp["gt_boxes"], inp["gt_labels"], inp["gt_valid"],
        static["sup_index"], static["unsup_index"], inp["ulb_pos"], aug["t"], aug["s"],
        student_kw=skw, teacher_kw=tkw)
    return sum(losses.values())


# ------------------------------------------------------------------------------------------------
# CPU reference arm (oracle port)
# ------------------------------------------------------------------------------------------------
def cpu_reference(workload, steps, warmup, scenes=None, forward_only_scene=False):
    """The oracle port of the step on the host cores -> (scenes/s, s/step, cores, scenes per step)."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    from oracle import cpu as oracle_cpu
    oracle_cpu.set_threads(cores)            # torchrun exports OMP_NUM_THREADS=1
    torch.manual_seed(0)
    model = build_model(workload, oracle=True)
    if workload == "mean_teacher":
        model.init_teacher()
    S = scenes or WORKLOADS[workload]["scenes"]
    inp = make_host_batch(workload, 9000, S)
    static = None
    if workload == "mean_teacher":
        static = dict(sup_index=torch.arange(S // 2), unsup_index=torch.arange(S // 2, S))
    if forward_only_scene:
        pts = inp["pts"][:1]
        times = []
        with torch.no_grad():
            for it in range(warmup + steps):
                t0 = time.perf_counter()
                model.predict(pts) if hasattr(model, "predict") else model.forward(pts)
                if it >= warmup:
                    times.append(time.perf_counter() - t0)
        times.sort()
        return 1.0 / times[len(times) // 2], times[len(times) // 2], cores, 1
    opt = torch.optim.AdamW(model.parameters(), lr=0.008, weight_decay=0.01)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        loss = step_loss(workload, model, inp, static=static)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 10.0)
        opt.step()
        if workload == "mean_teacher":
            model.after_train_iter(it)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    times.sort()
    sec = times[len(times) // 2]
    return S / sec, sec, cores, S


def cpu_sample_scenes(workload):
    """Bounded CPU sample (10-30 s of host work for 1 warm-up + 3 timed steps on the GPU box's 16
    cores): the full 8-scene batch for the pretrain step, 4 labeled + 4 unlabeled scenes for the
    mean-teacher step, 2 scenes of the 100k-point stress."""
    return {"stress": 2, "mean_teacher": 8}.get(workload, 8)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = args.workload
    steps, warmup = max(min(args.steps, 3), 1), min(args.warmup, 1)
    S = cpu_sample_scenes(wl)
    rate, sec, cores, S = cpu_reference(wl, steps, warmup, scenes=S)
    sample = (f"{S} scenes/step ({WORKLOADS[wl]['points']} pts each), full step on the host cores, "
              f"median of {steps} after {warmup} warm-up")
    line = {"impl": "reference", "metric": "train_scenes_per_s", "value": rate, "unit": "scenes/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": sec * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": WORKLOADS[wl]["name"], "scenes_per_step": S},
            "cpu_baseline": {"value": rate, "unit": "scenes/s", "cores": cores, "kind": "port",
                             "sample": sample},
            "e2e": {"value": rate, "unit": "scenes/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "note": "the reference has no CPU implementation of these ops; this is the oracle "
                    "port (C/OpenMP restatement of its CUDA kernels + torch-CPU MLPs)"}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# per-shape census of the GEMM kernels of one step (roofline)
# ------------------------------------------------------------------------------------------------
def gemm_census(trace, dev, min_reps=5, hbm_gbs=6459.6, bf16_tflops=1684.1):
    """trace: [(entry point, args)] of one eager step.  Times every distinct GEMM launch shape live
    (back to back, CUDA events) and returns per-family step-weighted numbers."""
    from nesie_b200 import _lib
    from nesie_b200 import linear_rows as lr
    fam = {"nesie_gemm_nt_3xtf32": "gemm_nt", "nesie_gemm_nt_3xtf32_fused": "gemm_nt",
           "nesie_gemm_wgrad_3xtf32": "gemm_wgrad", "nesie_gemm_wgrad_3xtf32_fused": "gemm_wgrad"}
    shapes = {}
    for name, a in trace:
        if name in fam:
            key = (fam[name], int(a[0]), int(a[1]), int(a[2]), 1, 0, 0)
            shapes[key] = shapes.get(key, 0) + 1
        elif name == "nesie_gemm_nt_3xtf32_pool":
            # (store the output?, pooling unit, rows per group bias): the launch is timed in its own mode
            key = ("gemm_nt", int(a[0]), int(a[1]), int(a[2]), int(a[6] is not None),
                   int(a[11]) if a[12] is not None else 0, int(a[17]) if a[16] is not None else 0)
            shapes[key] = shapes.get(key, 0) + 1
    out = {}
    for (family, R, N, K, store, pool_u, grp_k), count in sorted(shapes.items()):
        if R < 1:
            continue
        if family == "gemm_nt" and (not store or pool_u or grp_k):
            per = 4 * R * K + (4 * R * N if store else 0)
            ncopy = max(1, min(8, int(300e6 // max(per, 1)) + 1))
            xs = [torch.randn(R, K, device=dev) for _ in range(ncopy)]
            ys = [torch.empty(R, N, device=dev) if store else None for _ in range(ncopy)]
            img = lr._pack(torch.randn(N, K, device=dev), N, K, K, 1)
            pmax = torch.empty((R // pool_u, N), device=dev) if pool_u else None
            amax = torch.empty((R // pool_u, N), dtype=torch.uint8, device=dev) if pool_u else None
            grp = torch.randn((R // grp_k, N), device=dev) if grp_k else None

            def launch(i):
                _lib.call("nesie_gemm_nt_3xtf32_pool", R, N, K, _lib.ptr(xs[i % ncopy]), K, _lib.ptr(img),
                          _lib.ptr(ys[i % ncopy]), N, None, None, None, pool_u, _lib.ptr(pmax), _lib.ptr(amax),
                          None, None, _lib.ptr(grp), grp_k, _lib.stream())
        elif family == "gemm_nt":
            # operands rotate through enough copies that nothing is served from the 126 MB L2
            per = 4 * R * (K + N)
            ncopy = max(1, min(8, int(300e6 // max(per, 1)) + 1))
            xs = [torch.randn(R, K, device=dev) for _ in range(ncopy)]
            ys = [torch.empty(R, N, device=dev) for _ in range(ncopy)]
            img = lr._pack(torch.randn(N, K, device=dev), N, K, K, 1)

            def launch(i):
                _lib.call("nesie_gemm_nt_3xtf32", R, N, K, _lib.ptr(xs[i % ncopy]), K, _lib.ptr(img),
                          _lib.ptr(ys[i % ncopy]), N, _lib.stream())
        else:
            if N > 256 or K > 512:
                continue
            per = 4 * R * (K + N)
            ncopy = max(1, min(8, int(300e6 // max(per, 1)) + 1))
            gs = [torch.randn(R, N, device=dev) for _ in range(ncopy)]
            xs = [torch.randn(R, K, device=dev) for _ in range(ncopy)]
            ns = _lib.lib().nesie_gemm_wgrad_splits(R, N, K)
            parts = torch.empty((ns, N, K), device=dev)

            def launch(i):
                _lib.call("nesie_gemm_wgrad_3xtf32", R, N, K, _lib.ptr(gs[i % ncopy]), N,
                          _lib.ptr(xs[i % ncopy]), K, _lib.ptr(parts), ns, _lib.stream())
        reps = max(min_reps, ncopy)
        for i in range(2):
            launch(i)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(reps):
            launch(i)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps
        if os.environ.get("NESIE_BENCH_CENSUS_DUMP"):
            mode = ("" if store else " no-store") + (f" pool{pool_u}" if pool_u else "") + (f" grp{grp_k}" if grp_k else "")
            sys.stderr.write(f"[census] {family} R={R} K={K} N={N}{mode} x{count}: {ms * 1e3:.1f} us, "
                             f"{per / (ms * 1e-3) / 1e9:.0f} GB/s, {2.0 * R * N * K / (ms * 1e-3) / 1e12:.1f} TF/s fp32-equiv\n")
        f = out.setdefault(family, dict(bytes=0.0, ms=0.0, launches=0, flops=0.0, best=None, bound_ms=0.0))
        # per-launch lower bound: HBM time of the algorithmic bytes vs tensor time of the 3 TF32 MMAs per
        # fp32 product at half the measured bf16 rate
        f["bound_ms"] += count * max(per / (hbm_gbs * 1e9), 3 * 2.0 * R * N * K / (bf16_tflops / 2 * 1e12)) * 1e3
        f["bytes"] += count * per
        f["ms"] += count * ms
        f["launches"] += count
        f["flops"] += count * 2.0 * R * N * K
        gbs = per / (ms * 1e-3) / 1e9
        if per >= 64e6 and (f["best"] is None or gbs > f["best"]["gbs"]):
            f["best"] = dict(shape=[R, K, N], ms=ms, gbs=gbs, bytes=per)
        del xs
    return out


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="nesie_b200", choices=["nesie_b200", "reference"])
    ap.add_argument("--workload", default="pretrain", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-census", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    import nesie_b200 as nb
    from nesie_b200 import _lib
    from nesie_b200.ddp import FlatGradDDP

    wl = args.workload
    cfg = WORKLOADS[wl]
    S = cfg["scenes"]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    W, K = max(args.warmup, 3), args.steps

    torch.manual_seed(0)
    model = build_model(wl).to(dev)
    # flat gradient AND parameter buffers (bucketed all-reduce, broadcast of rank 0): AdamW, clipping and
    # the teacher EMA are one elementwise launch each over the flat buffers
    ddp = FlatGradDDP(model, flatten_parameters=True)
    if wl == "mean_teacher":
        model.init_teacher(flat=ddp)
    opt = torch.optim.AdamW([ddp.flat_parameter()], lr=0.008, weight_decay=0.01, fused=True, capturable=True)
    backbone = model.backbone
    static = None
    if wl == "mean_teacher":
        static = dict(sup_index=torch.arange(S // 2, device=dev), unsup_index=torch.arange(S // 2, S, device=dev))

    # a pool of distinct batches (so no step re-reads the previous step's inputs from L2)
    NB = 4
    host = [make_host_batch(wl, 1000 * rank + 100 * i) for i in range(NB)]
    host = [{k: v.pin_memory() for k, v in h.items()} for h in host]
    resident = [{k: v.to(dev) for k, v in h.items()} for h in host]
    h2d = sum(v.numel() * v.element_size() for v in host[0].values())
    # Input pipeline.  The FPS chain depends on coordinates only, so the chain of batch t+1 is computed
    # while batch t trains.  FPS is a serial, latency-bound kernel whose CTAs cannot share an SM with
    # the persistent GEMM CTAs, so WHERE it runs matters: it is forked inside the captured step after
    # SA level NESIE_BENCH_FPS_AT (default 3 = after the last SA level, beside the small-grid middle of
    # the step); split mode runs the long first FPS level of batch t+2 and the short remaining levels of
    # batch t+1 side by side on two branches (four input slots).
    fork_level = int(os.environ.get("NESIE_BENCH_FPS_AT", "3"))
    split = os.environ.get("NESIE_BENCH_FPS_SPLIT", "1") == "1"
    NSLOT = 4 if split else 3
    n_fps_scenes = resident[0]["pts"].shape[0]
    slots = []
    for _ in range(NSLOT):
        slots.append(dict(inp={k: torch.empty_like(v) for k, v in resident[0].items()},
                          fps=[torch.zeros((n_fps_scenes, n), dtype=torch.int32, device=dev)
                               for n in backbone.num_points],
                          ev_fps=torch.cuda.Event(), ev_step=torch.cuda.Event(),
                          ev_load=torch.cuda.Event()))
    s_loss = torch.zeros((), device=dev)
    side = torch.cuda.Stream()
    side2 = torch.cuda.Stream()
    copy_stream = torch.cuda.Stream()
    step_no = {"n": 0}

    def fps_body(slot):
        for dst, src in zip(slot["fps"], backbone.fps_chain(slot["inp"]["pts"])):
            dst.copy_(src)

    def fps_first(slot):
        slot["fps"][0].copy_(backbone.fps_chain(slot["inp"]["pts"], stop=1)[0])

    def fps_rest(slot):
        for dst, src in zip(slot["fps"][1:], backbone.fps_chain(slot["inp"]["pts"], given=slot["fps"][:1])):
            dst.copy_(src)

    def step_body(slot, nxt=None, nxt2=None):
        """One training step on `slot`; with `nxt`, the FPS chain of the next batch is forked onto the
        side stream after SA level `fork_level` and joined at the end of the step (split mode: the
        remaining levels of `nxt` and the first level of `nxt2` on two branches)."""
        cur = torch.cuda.current_stream()
        hook = None
        if os.environ.get("NESIE_BENCH_SKIP_FPS") == "1":     # diagnostic: the step without any FPS work
            nxt = nxt2 = None
        if nxt is not None:
            def hook(i):
                if i == fork_level:
                    side.wait_stream(cur)
                    if nxt2 is None:
                        with torch.cuda.stream(side):
                            fps_body(nxt)
                    else:
                        side2.wait_stream(cur)
                        with torch.cuda.stream(side):
                            fps_first(nxt2)
                        with torch.cuda.stream(side2):
                            fps_rest(nxt)
        ddp.zero_grad()
        loss = step_loss(wl, model, slot["inp"], slot["fps"], hook, static)
        loss.backward()
        ddp.finish()
        ddp.clip_grad_norm_(10.0)
        opt.step()
        if wl == "mean_teacher":
            model.after_train_iter(10 + step_no["n"])   # past the warm-up: constant momentum 0.001
        s_loss.copy_(loss.detach())
        if nxt is not None:
            cur.wait_stream(side)
            if nxt2 is not None:
                cur.wait_stream(side2)

    def load_inputs(slot, src):
        for k, d in slot["inp"].items():
            d.copy_(src[k], non_blocking=True)

    def nxt_of(j):
        return slots[(j + 1) % NSLOT]

    def nxt2_of(j):
        return slots[(j + 2) % NSLOT] if split else None

    # warm up eagerly on the side stream (also initialises NCCL), then capture each piece once
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for sl in slots:
            load_inputs(sl, resident[0])
            fps_body(sl)
        for _ in range(3):
            step_body(slots[0])
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    mode = "eager"
    step_fn = [lambda j=j: step_body(slots[j], nxt_of(j), nxt2_of(j)) for j in range(NSLOT)]
    fps_fn = [lambda sl=sl: fps_body(sl) for sl in slots]
    graphs = []
    if os.environ.get("NESIE_BENCH_GRAPH", "1") != "0":
        try:
            for j, sl in enumerate(slots):
                g1 = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g1):
                    step_body(sl, nxt_of(j), nxt2_of(j))
                graphs.append(g1)
            step_fn = [g.replay for g in graphs]
            mode = "cuda_graph"
        except Exception as e:  # noqa: BLE001 -- fall back to eager launches
            graphs = []
            torch.cuda.synchronize()
            sys.stderr.write(f"[bench] CUDA-graph capture failed, running eagerly: {e!r}\n")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def load_async(i, src):
        """copy stream: batch i -> slot i % NSLOT, once the slot's previous step has finished."""
        sl = slots[i % NSLOT]
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(sl["ev_step"])
            load_inputs(sl, src[i % NB])
            sl["ev_load"].record(copy_stream)

    def run_pipeline(nsteps, src, after_step=None):
        main_s = torch.cuda.current_stream()
        for sl in slots:
            sl["ev_step"].record(main_s)
        # prologue: batch 0 loaded and sampled, batch 1 loaded (split: + its first FPS level)
        load_async(0, src)
        with torch.cuda.stream(side):
            side.wait_event(slots[0]["ev_load"])
            fps_fn[0]()
            slots[0]["ev_fps"].record(side)
        load_async(1, src)
        if split:
            with torch.cuda.stream(side):
                side.wait_event(slots[1]["ev_load"])
                fps_first(slots[1])
                slots[1]["ev_fps"].record(side)
            load_async(2, src)
            main_s.wait_event(slots[1]["ev_fps"])
        main_s.wait_event(slots[0]["ev_fps"])
        for i in range(nsteps):
            sl, nx = slots[i % NSLOT], slots[(i + 1) % NSLOT]
            load_async(i + NSLOT - 1, src)                 # overlaps this step
            main_s.wait_event(nx["ev_load"])
            if split:
                main_s.wait_event(slots[(i + 2) % NSLOT]["ev_load"])
            step_fn[i % NSLOT]()                           # samples batch i+1 (and i+2) inside
            step_no["n"] += 1
            sl["ev_step"].record(main_s)
            if after_step is not None:
                after_step()
        main_s.wait_stream(side)
        main_s.wait_stream(copy_stream)

    def timed(fn):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        barrier()
        ms = a.elapsed_time(b)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    # ---- device-resident throughput -----------------------------------------------------------
    run_pipeline(W, resident)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms = timed(lambda: run_pipeline(K, resident))
    clocks = sampler.stop() if rank == 0 else None
    value = world * S * K / (ms / 1e3)

    if os.environ.get("NESIE_BENCH_TRACE") and rank == 0:
        # diagnostic: kernel timeline (name, stream, start, duration) of 3 pipelined steps
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            run_pipeline(3, resident)
            torch.cuda.synchronize()
        prof.export_chrome_trace(os.environ["NESIE_BENCH_TRACE"])

    # ---- end to end: pinned host -> device every step, loss read back every step ---------------
    # The loss of every step is copied to pinned host memory right behind the step; the host waits
    # for it one step later (the way a training loop logs), so the copy does not drain the pipeline.
    sinks = [torch.zeros(1).pin_memory() for _ in range(2)]
    sink_ev = [torch.cuda.Event(), torch.cuda.Event()]
    seen = {"n": 0, "last": 0.0}

    def read_loss():
        i = seen["n"]
        sinks[i & 1].copy_(s_loss.reshape(1), non_blocking=True)   # device -> host
        sink_ev[i & 1].record()
        if i > 0:
            sink_ev[(i - 1) & 1].synchronize()
            seen["last"] = float(sinks[(i - 1) & 1][0])
        seen["n"] = i + 1

    def e2e_run(n):
        run_pipeline(n, host, read_loss)
        if seen["n"] > 0:
            sink_ev[(seen["n"] - 1) & 1].synchronize()
            seen["last"] = float(sinks[(seen["n"] - 1) & 1][0])

    e2e_run(2)
    ms_e2e = timed(lambda: e2e_run(K))
    e2e_value = world * S * K / (ms_e2e / 1e3)
    final_loss = seen["last"]

    # ---- kernels of this repo per step + per-shape GEMM census (eager, outside the timed region) --
    _lib.TRACE = []
    l0 = _lib.LAUNCHES
    with torch.cuda.stream(side):
        fps_body(slots[0])
        step_body(slots[0])
    torch.cuda.synchronize()
    launches = (_lib.LAUNCHES - l0) * K
    trace, _lib.TRACE = _lib.TRACE, None
    pk, pk_kind = peaks()
    roofline = None
    if rank == 0 and not args.no_census:
        census = gemm_census(trace, dev, hbm_gbs=pk["hbm_gbs"], bf16_tflops=pk["bf16_tflops"])
        if census:
            name = max(census, key=lambda k: census[k]["ms"])
            c = census[name]
            achieved = c["bytes"] / (c["ms"] * 1e-3) / 1e9
            kern = {"gemm_nt": "gemm_nt_tma_kernel (3xTF32 tcgen05 row GEMM: forward + data gradient)",
                    "gemm_wgrad": "gemm_wgrad_tma_kernel (3xTF32 tcgen05 weight gradient)"}[name]
            roofline = {"kernel": kern, "bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"],
                        "unit": "GB/s", "frac": achieved / pk["hbm_gbs"], "traffic": GEMM_DRAM_TRAFFIC,
                        "traffic_launch": "SA1 layer 3 (1048576 x 64 -> 128): 805.3 MB algorithmic per launch",
                        "peak_source": pk_kind, "weighting": "step-weighted: sum of algorithmic bytes "
                        "4*R*(K+N) of the family's launches in one step / sum of their live-timed durations",
                        "launches_per_step": c["launches"], "kernel_ms_per_step": c["ms"],
                        "algorithmic_bytes_per_step": c["bytes"],
                        "tensor_tflops_fp32_equiv": c["flops"] / (c["ms"] * 1e-3) / 1e12,
                        "frac_of_max_hbm_tensor_bound": c["bound_ms"] / c["ms"],
                        "best_shape": c["best"],
                        "best_shape_frac": (c["best"]["gbs"] / pk["hbm_gbs"]) if c["best"] else None,
                        "families": {k: dict(ms_per_step=v["ms"], launches=v["launches"],
                                             gbs=v["bytes"] / (v["ms"] * 1e-3) / 1e9,
                                             frac=v["bytes"] / (v["ms"] * 1e-3) / 1e9 / pk["hbm_gbs"],
                                             frac_of_max_hbm_tensor_bound=v["bound_ms"] / v["ms"])
                                     for k, v in census.items()}}

    line = {"metric": "train_scenes_per_s", "value": value, "unit": "scenes/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["name"], "scenes_per_gpu": S, "points": cfg["points"],
                       "classes": 18, "parallelism": f"dp{world}",
                       "l2": "4 distinct resident batches cycled; per-step activations exceed L2",
                       "launch": mode, "ddp": f"{len(ddp.buckets)} gradient buckets, {ddp.flat.numel() * 4} bytes, "
                                              "all-reduce overlapped with backward",
                       "input_pipeline": f"batch t+2 copied and batch t+1 sampled (FPS chain) during step t; "
                                         f"FPS forked inside the captured step after SA level {fork_level}"},
            "e2e": {"value": e2e_value, "unit": "scenes/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / K},
            "gpu_launches": launches, "roofline": roofline, "clocks": clocks,
            "final_loss": final_loss}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        Sc = cpu_sample_scenes(wl)
        rate, sec, cores, Sc = cpu_reference(wl, 3, 1, scenes=Sc)
        line["cpu_baseline"] = {"value": rate, "unit": "scenes/s", "cores": cores, "kind": "port",
                                "sample": f"{Sc} scenes/step ({cfg['points']} pts each), full step, "
                                          f"median of 3 after 1 warm-up ({sec:.2f} s/step)"}
        if wl != "pretrain_conv":
            r1, s1, _, _ = cpu_reference(wl, 3, 1, scenes=2, forward_only_scene=True)
            line["cpu_baseline"]["forward_1scene"] = {
                "value": r1, "unit": "scenes/s", "ms": s1 * 1e3,
                "what": "BASELINE configs[0]: detector forward on 1 scene, batch 1, median of 3"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    # Orderly teardown: graphs first (they hold the captured NCCL work), then the process group.  A
    # watchdog ends the process if NCCL's destructor wedges, so a hang can never hold the box.
    sys.stdout.flush()
    sys.stderr.flush()
    threading.Thread(target=lambda: (time.sleep(90), os._exit(0)), daemon=True).start()
    del step_fn, graphs
    ddp.remove_hooks()
    barrier()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
