#!/usr/bin/env python
"""bench.py -- the reference's headline metric on B200: ROIAlign+ARD RoIs/s, forward+backward.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--layout nhwc|nchw|nchw_cl]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

A "step" is one pass of the RoI hot path over one batch of synthetic input at the shapes of BASELINE.json
configs[1] (VOC 15-5 ABR incremental step, batch 4 per GPU): teacher and student R-50-C4 feature maps
[4,1024,50,76] fp32 (800x1216 images, stride 16), 512 RoIs per image, POOLER_RESOLUTION 7, adaptive sampling
(sampling_ratio 0).  The work unit is the composite of SURVEY.md section 8d: one RoI through teacher ROIAlign
forward, student ROIAlign forward, ARD loss forward+backward (gamma=1) and student ROIAlign backward.
Work shards by image, so N GPUs run N independent batches (weak scaling, no data-path collective).

One JSON line on rank 0: `value` = RoIs/s with inputs resident in HBM (CUDA events over exactly K steps, max over
ranks); `e2e` = the same metric through the public Python API with HOST buffers (pinned H2D of both feature maps
and the RoIs, D2H of the loss and of the student feature-map gradient inside the timed region); `roofline` for the
dominant kernel from per-kernel CUDA events; `cpu_baseline` = the reference's CPU path on this box's host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOAD = dict(B=4, C=1024, H=50, W=76, rois_per_image=512, P=7, sampling_ratio=0, scale=1.0 / 16,
                image_w=1216, image_h=800)
METRIC = "ROIAlign+ARD RoIs/s fwd+bwd"
UNIT = "RoIs/s"


def make_workload(seed=0, rois=None):
    """Synthetic VOC-shaped batch: student = teacher + 0.1*noise; RoI centres uniform in the image,
    sides U(16,400) px, clipped to the image, 5 % degenerate (< 1 px)."""
    w = WORKLOAD
    rng = np.random.default_rng(seed)
    teacher = rng.standard_normal((w["B"], w["C"], w["H"], w["W"]), dtype=np.float32)
    student = teacher + np.float32(0.1) * rng.standard_normal(teacher.shape, dtype=np.float32)
    R = w["B"] * w["rois_per_image"] if rois is None else rois
    cx, cy = rng.uniform(0, w["image_w"], R), rng.uniform(0, w["image_h"], R)
    bw, bh = rng.uniform(16, 400, R), rng.uniform(16, 400, R)
    deg = rng.uniform(0, 1, R) < 0.05
    bw[deg] = rng.uniform(0, 1, deg.sum())
    x1, x2 = np.clip(cx - bw / 2, 0, w["image_w"] - 1), np.clip(cx + bw / 2, 0, w["image_w"] - 1)
    y1, y2 = np.clip(cy - bh / 2, 0, w["image_h"] - 1), np.clip(cy + bh / 2, 0, w["image_h"] - 1)
    img = np.repeat(np.arange(w["B"]), -(-R // w["B"]))[:R]
    roi = np.stack([img, x1, y1, x2, y2], 1).astype(np.float32)
    return teacher, student, roi


def algorithmic_bytes(R):
    """per kernel and for the composite unit (3*BCHW*s + 60R + 6*R*C*P^2*s)."""
    w = WORKLOAD
    fmap = w["B"] * w["C"] * w["H"] * w["W"] * 4
    pooled = R * w["C"] * w["P"] * w["P"] * 4
    return {"roi_align_fwd": fmap + 20 * R + pooled, "roi_align_bwd": pooled + 20 * R + fmap, "ard": 3 * pooled,
            "composite": 3 * fmap + 60 * R + 6 * pooled}


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def summary(self, t0, t1):
        if self.proc is not None:
            self.proc.terminate()
        rows = [r for t, r in self.rows if t0 - 0.05 <= t <= t1 + 0.15 and len(r) >= 6] or [r for _, r in self.rows if len(r) >= 6]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in rows)
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]), "reasons": reasons, "samples": len(rows)}


# ------------------------------------------------------------------------------------------------ CPU reference
def cpu_reference_sample(n_rois, threads=None):
    """The reference's CPU path for the composite unit on `n_rois` RoIs of the workload.  ROIAlign forward is the
    reference's own ROIAlign_forward_cpu (oracle/_ref, single-threaded by construction: csrc/cpu/ROIAlign_cpu.cpp:133)
    when it was compiled, else the C port; ARD is the PyTorch op sequence of distillation.py:86-130 on all host
    threads; ROIAlign backward has no CPU reference (csrc/ROIAlign.h:44) and is timed with the C port of the CUDA
    kernel.  Returns (seconds, description)."""
    import torch

    import oracle
    from oracle import ard_torch

    if threads:
        torch.set_num_threads(threads)
    w = WORKLOAD
    teacher, student, rois = make_workload(0, rois=n_rois)
    use_ref = oracle.ref_available()
    t0 = time.perf_counter()
    f_old = oracle.roi_align_forward(teacher, rois, w["scale"], w["P"], w["P"], w["sampling_ratio"], use_ref=use_ref)
    f_new = oracle.roi_align_forward(student, rois, w["scale"], w["P"], w["P"], w["sampling_ratio"], use_ref=use_ref)
    _, grad = ard_torch.ard_fwd_bwd(torch.from_numpy(f_old), torch.from_numpy(f_new), 1.0)
    oracle.roi_align_backward(grad.numpy(), rois, w["scale"], w["P"], w["P"], w["B"], w["C"], w["H"], w["W"],
                              w["sampling_ratio"])
    dt = time.perf_counter() - t0
    kind = "reference" if use_ref else "port"
    desc = ("%d RoIs of the same workload: ROIAlign_forward_cpu x2 (%s, 1 thread), PyTorch ARD fwd+bwd (%d threads), "
            "ROIAlign backward (C port of the CUDA kernel, 1 thread; the reference has no CPU backward)"
            % (n_rois, "oracle/_ref" if use_ref else "oracle port", torch.get_num_threads()))
    return dt, kind, desc, torch.get_num_threads()


def run_reference(args, rank, world):
    if rank != 0:
        return
    # calibrate the bounded sample so that K steps take about a minute in total (8..256 RoIs per step)
    cpu_reference_sample(8)
    per_roi = cpu_reference_sample(16)[0] / 16
    sample = int(max(8, min(256, 60.0 / max(args.steps, 1) / per_roi)))
    for _ in range(min(args.warmup, 1)):
        cpu_reference_sample(sample)
    t, kind, desc, cores = 0.0, "port", "", 1
    for _ in range(args.steps):
        dt, kind, desc, cores = cpu_reference_sample(sample)
        t += dt
    value = sample * args.steps / t
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(), sample_rois_per_step=sample),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": desc},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def workload_config():
    w = WORKLOAD
    return {"workload": "configs[1] VOC 15-5 ABR step, RoI hot path: teacher+student R-50-C4 maps [%d,%d,%d,%d] fp32, "
                        "%d RoIs/img, P=%d, sampling_ratio=%d; unit = teacher ROIAlign fwd + student ROIAlign fwd + "
                        "ARD fwd+bwd + student ROIAlign bwd" % (w["B"], w["C"], w["H"], w["W"], w["rois_per_image"],
                                                              w["P"], w["sampling_ratio"]),
            "batch_per_gpu": w["B"], "rois_per_gpu": w["B"] * w["rois_per_image"], "parallelism": "data-parallel by image",
            "l2": "per-step inputs+outputs 2.65 GB >> 126 MB L2, no explicit flush"}


# ------------------------------------------------------------------------------------------------ ours
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    from abr_iod_b200 import _lib
    from abr_iod_b200.distillation.distillation import _ard_launch
    from abr_iod_b200.distillation.distillation import calculate_attentive_roi_feature_distillation as ard
    from abr_iod_b200.layers import ROIAlign
    from abr_iod_b200.layers.roi_align import roi_align_backward, roi_align_forward

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    sampler = ClockSampler(local_rank) if rank == 0 else None  # started early: nvidia-smi needs a second to warm up
    w = WORKLOAD
    nhwc = args.layout == "nhwc"
    _lib.POOLED_CHANNELS_LAST = args.layout == "nchw_cl"  # contiguous maps, channels-last RoI features
    fmt = torch.channels_last if nhwc else torch.contiguous_format
    teacher_np, student_np, rois_np = make_workload(seed=rank)
    R = rois_np.shape[0]
    teacher = torch.from_numpy(teacher_np).to(dev).contiguous(memory_format=fmt)
    student = torch.from_numpy(student_np).to(dev).contiguous(memory_format=fmt)
    rois = torch.from_numpy(rois_np).to(dev)
    P, ratio, scale = w["P"], w["sampling_ratio"], w["scale"]
    names = ("roi_align_fwd_teacher", "roi_align_fwd_student", "ard", "roi_align_bwd")

    def step(ev=None):
        marks = []

        def mark():
            if ev is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                marks.append(e)
        mark()
        f_old, plan = roi_align_forward(teacher, rois, scale, P, P, ratio, return_plan=True)
        mark()
        f_new = roi_align_forward(student, rois, scale, P, P, ratio, plan=plan)  # same RoIs and map shape: plans reused
        mark()
        loss3, g = _ard_launch(f_old, f_new, 1.0, True)
        mark()
        gin = roi_align_backward(g, rois, scale, P, P, w["B"], w["C"], w["H"], w["W"], ratio, layout=_lib.roi_align_layout(student), plan=plan)
        mark()
        if ev is not None:
            ev.append(marks)
        return loss3, gin

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    launches0 = _lib.launch_count()
    events = []
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    start.record()
    for _ in range(args.steps):
        step(events)
    end.record()
    barrier()
    t1 = time.time()
    elapsed_ms = start.elapsed_time(end)
    launches = _lib.launch_count() - launches0
    clocks = sampler.summary(t0, t1) if sampler else None
    per_kernel = {n: sum(m[i].elapsed_time(m[i + 1]) for m in events) / len(events) for i, n in enumerate(names)}

    # ---- end to end through the public API with host buffers
    pool = ROIAlign((P, P), scale, ratio)
    h_teacher = torch.from_numpy(teacher_np).contiguous(memory_format=fmt).pin_memory()
    h_student = torch.from_numpy(student_np).contiguous(memory_format=fmt).pin_memory()
    h_rois = torch.from_numpy(rois_np).pin_memory()
    # Two streams, steps alternate between them: every step still pays its own H2D of both maps + RoIs and its own D2H of
    # the loss and the gradient map, but consecutive steps overlap (copy engines in both directions + SMs), the way a
    # double-buffered input pipeline feeds a training loop.
    streams = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
    h_grad = [torch.empty_like(h_student).pin_memory() for _ in streams]
    h_loss = [torch.empty((), dtype=torch.float32).pin_memory() for _ in streams]

    def e2e_step(i):
        k = i % len(streams)
        with torch.cuda.stream(streams[k]):
            t = h_teacher.to(dev, non_blocking=True)
            s = h_student.to(dev, non_blocking=True).requires_grad_(True)
            r = h_rois.to(dev, non_blocking=True)
            with torch.no_grad():
                f_old = pool(t, r)
            f_new = pool(s, r)
            loss = ard(f_old, f_new, 1.0)
            loss.backward()
            h_loss[k].copy_(loss.detach(), non_blocking=True)
            h_grad[k].copy_(s.grad, non_blocking=True)

    for i in range(4):
        e2e_step(i)
    barrier()
    e2e_steps = max(4, min(args.steps, 40))
    es, ee = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for st_ in streams:
        st_.wait_stream(torch.cuda.current_stream(dev))
    es.record(streams[0])
    streams[1].wait_event(es)
    for i in range(e2e_steps):
        e2e_step(i)
    streams[0].wait_stream(streams[1])
    ee.record(streams[0])
    barrier()
    e2e_ms = es.elapsed_time(ee)

    if world > 1:
        t = torch.tensor([elapsed_ms, e2e_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms, e2e_ms = t.tolist()
        lt = torch.tensor([launches], device=dev, dtype=torch.int64)
        dist.all_reduce(lt)
        launches = int(lt.item())
    if rank != 0:
        return

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy)"
    else:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    ab = algorithmic_bytes(R)
    kbytes = {"roi_align_fwd_teacher": ab["roi_align_fwd"], "roi_align_fwd_student": ab["roi_align_fwd"], "ard": ab["ard"],
              "roi_align_bwd": ab["roi_align_bwd"]}
    dominant = max(per_kernel, key=per_kernel.get)
    achieved = kbytes[dominant] / (per_kernel[dominant] * 1e-3) / 1e9
    traffic, limiter = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        traffic = tj.get(args.layout, {}).get(dominant)
        # The dominant kernel is not DRAM-bound on this input (the map is L2-resident and RoIs overlap ~23x): next to the
        # required HBM roofline, report it against the on-chip resource that does bound it -- bytes per launch from ncu,
        # peak from the committed microbenchmarks, time from this run.
        lim = tj.get("l2_limits", {}).get(dominant) if args.layout == "nhwc" else None
        if lim:
            got = lim["bytes_per_launch"] / (per_kernel[dominant] * 1e-3) / 1e9
            limiter = {"resource": lim["resource"], "achieved": round(got, 1), "peak": lim["peak_gbs"], "unit": "GB/s",
                       "frac": round(got / lim["peak_gbs"], 4), "source": lim["source"]}
    ms_per_step = elapsed_ms / args.steps
    value = world * R / (ms_per_step * 1e-3)
    kernels = {n: {"ms": round(per_kernel[n], 4), "GB/s": round(kbytes[n] / (per_kernel[n] * 1e-3) / 1e9, 1),
                   "frac": round(kbytes[n] / (per_kernel[n] * 1e-3) / 1e9 / peak, 4)} for n in names}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": dict(workload_config(), layout=args.layout),
        "e2e": {"value": world * R / (e2e_ms / e2e_steps * 1e-3), "unit": UNIT,
                "h2d_bytes_per_step": int(h_teacher.numel() * 4 + h_student.numel() * 4 + h_rois.numel() * 4),
                "d2h_bytes_per_step": int(h_grad[0].numel() * 4 + 4), "steps": e2e_steps,
                "pipelining": "2 streams, consecutive steps overlap"},
        "gpu_launches": launches, "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": dominant, "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                     "frac": round(achieved / peak, 4), "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": kbytes[dominant], "limiter": limiter},
        "composite": {"algorithmic_bytes_per_step": ab["composite"],
                      "GB/s": round(ab["composite"] / (ms_per_step * 1e-3) / 1e9, 1),
                      "frac_of_hbm_peak": round(ab["composite"] / (ms_per_step * 1e-3) / 1e9 / peak, 4)},
        "kernels": kernels,
    }
    if world == 1 and not args.no_cpu:
        dt, kind, desc, cores = cpu_reference_sample(args.cpu_rois)
        line["cpu_baseline"] = {"value": args.cpu_rois / dt, "unit": UNIT, "cores": cores, "kind": kind, "sample": desc}
    if world == 1 and not args.no_secondary:
        line["secondary"] = secondary_metrics(dev)
    print(json.dumps(line))


def _time_call(fn, graph, reps=10):
    """Seconds per call of `fn` (CUDA events on the current stream, after warm-up).  graph=False: issued from Python call
    by call (includes the wrapper's allocations and launches whenever the host is the slower side); graph=True: the same
    call captured once in a CUDA graph and replayed -- the library only enqueues work on the caller's stream, so a whole
    call is capturable -- which is the device time of the launch sequence."""
    import torch

    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if not graph:
        s.record()
        for _ in range(reps):
            fn()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) * 1e-3 / reps
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        keep_alive = fn()  # noqa: F841  (outputs and workspaces stay allocated in the graph's pool)
    g.replay()
    torch.cuda.synchronize()
    s.record()
    for _ in range(reps):
        g.replay()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) * 1e-3 / reps


def secondary_metrics(dev):
    """BASELINE.json's other two numbers, measured in the same run: batched RPN NMS boxes/s (config 3) and ABR paste
    imgs/s (config 4, one GPU's shard)."""
    import torch

    from abr_iod_b200.layers import nms_batched
    from inputs import make_boxes

    out = {}
    rng = np.random.default_rng(3)
    # RPN-shaped calls: per image N boxes, IoU 0.7, post-NMS cut; the RPN hands NMS the output of a
    # sorted top-k (modeling/rpn/inference.py:94-95), so the scores arrive in descending order.  One unsorted line too.
    for n, batch, keep_n, is_sorted in ((6000, 4, 1000, True), (6000, 16, 1000, True), (12000, 4, 2000, True), (12000, 16, 2000, True),
                                        (6000, 4, 1000, False)):
        data = [make_boxes(rng, n, 1216, 800) for _ in range(batch)]
        if is_sorted:
            data = [(np.ascontiguousarray(b[np.argsort(-s, kind="stable")]), np.ascontiguousarray(np.sort(s)[::-1])) for b, s in data]
        boxes = [torch.from_numpy(b).to(dev) for b, _ in data]
        scores = [torch.from_numpy(s).to(dev) for _, s in data]
        key = "nms_boxes_per_s_n%d_b%d_keep%d_%s" % (n, batch, keep_n, "sorted" if is_sorted else "unsorted")
        out[key] = round(batch * n / _time_call(lambda: nms_batched(boxes, scores, 0.7, keep_n), graph=False))
        out[key + "_graph"] = round(batch * n / _time_call(lambda: nms_batched(boxes, scores, 0.7, keep_n), graph=True))
    out.update(rpn_metrics(dev))
    out.update(box_post_metrics(dev))
    out.update(paste_metrics(dev))
    return out


def box_post_metrics(dev):
    """Box-head post-processing (second half) on config-2 test shapes: batch 4, 1000 proposals per
    image, 21 classes (VOC), score 0.05, NMS 0.5, 100 detections per image; images/s with resident head outputs."""
    import torch

    from abr_iod_b200.modeling.roi_heads.box_head import box_postprocess

    rng = np.random.default_rng(7)
    N, n, C = 4, 1000, 21
    w, h = 1216, 800
    centers = rng.uniform([0.1 * w, 0.1 * h], [0.9 * w, 0.9 * h], (N, 12, 2))
    which = rng.integers(0, 12, (N, n))
    c = np.take_along_axis(centers, which[..., None].repeat(2, -1), 1) + rng.normal(0, 12, (N, n, 2))
    wh = rng.uniform(40, 300, (N, n, 2))
    props = np.clip(np.concatenate([c - wh / 2, c + wh / 2], -1), 0, [w - 1, h - 1, w - 1, h - 1]).astype(np.float32).reshape(-1, 4)
    logits = rng.normal(0, 1, (N * n, C)).astype(np.float32)
    logits[np.arange(N * n), 1 + which.reshape(-1) % (C - 1)] += rng.uniform(0, 6, N * n).astype(np.float32)
    reg = rng.normal(0, 0.5, (N * n, 4 * C)).astype(np.float32)
    lt, rt, pt = (torch.from_numpy(a).to(dev) for a in (logits, reg, props))
    call = lambda: box_postprocess(lt, rt, pt, [n] * N, [(w, h)] * N, 0.05, 0.5, 100)  # noqa: E731
    # (the wrapper reads the per-image counts once per call, so it cannot be graph-captured as a whole)
    return {"box_post_imgs_per_s_b4_r1000_c21": round(N / _time_call(call, graph=False))}


def rpn_metrics(dev):
    """The RPN proposal path around NMS on config-2 shapes: batch 4, 15 anchors on a 50x76 map
    (57,000 anchors per image), train (12000 -> 2000) and test (6000 -> 1000) settings; images/s for the whole device
    sequence (radix select, sort + decode, NMS, gather) with resident head outputs."""
    import torch

    from abr_iod_b200.modeling.rpn import rpn_proposals
    from inputs import make_anchors

    rng = np.random.default_rng(6)
    N, A, H, W = 4, 15, 50, 76
    anchors = torch.from_numpy(make_anchors(H, W, 16)).to(dev)
    obj = torch.from_numpy((rng.standard_normal((N, A, H, W)) * 2).astype(np.float32)).to(dev)
    reg = torch.from_numpy((rng.standard_normal((N, 4 * A, H, W)) * 0.3).astype(np.float32)).to(dev)
    sizes = [(W * 16, H * 16)] * N
    out = {}
    for pre, post in ((12000, 2000), (6000, 1000)):
        key = "rpn_proposals_imgs_per_s_b4_a57000_pre%d_post%d" % (pre, post)
        out[key] = round(N / _time_call(lambda: rpn_proposals(obj, reg, anchors, sizes, pre, post, 0.7, 0), graph=False))
        out[key + "_graph"] = round(N / _time_call(lambda: rpn_proposals(obj, reg, anchors, sizes, pre, post, 0.7, 0), graph=True))
    return out


def paste_metrics(dev, n_proto=2000, batch=16, rounds=8):
    """Config 4 on one GPU's shard: 2000 synthetic prototypes (sides U(71,300)), batches of 16 images 375x500 at the
    reference's 25/25/50 mixup/mosaic/untouched policy.  imgs/s for the whole call (host planning in the reference's
    draw order + pinned H2D + ONE paste launch) and for the kernel alone."""
    import random

    import torch
    from PIL import Image

    from abr_iod_b200.data.abr_paste import BoxRehearsalPaster

    rng = np.random.default_rng(4)
    protos = []
    for i in range(n_proto):
        h, w = int(rng.integers(71, 301)), int(rng.integers(71, 301))
        protos.append(("%d_%05d.jpg" % (int(rng.integers(1, 16)), i),
                       np.broadcast_to(rng.integers(0, 256, (1, 1, 3), dtype=np.uint8), (h, w, 3)).copy()))
    paster = BoxRehearsalPaster(protos, batch_size=batch, device=dev)
    images = [Image.fromarray(rng.integers(0, 256, (375, 500, 3), dtype=np.uint8)) for _ in range(batch)]
    targets = []
    for _ in range(batch):
        x1, y1 = rng.uniform(0, 250, 3), rng.uniform(0, 180, 3)
        targets.append(np.stack([x1, y1, x1 + rng.uniform(30, 200, 3), y1 + rng.uniform(30, 150, 3), rng.integers(16, 21, 3)], 1))
    random.seed(0)
    torch.manual_seed(0)
    paster.paste_batch(images, targets)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    plans_all = []
    for _ in range(rounds):
        plans = [paster.plan_transform(im, g) for im, g in zip(images, targets)]
        plans_all.append(plans)
        paster.execute(plans)
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # kernel + H2D only (plans precomputed)
    s.record()
    for plans in plans_all:
        paster.execute(plans)
    e.record()
    torch.cuda.synchronize()
    return {"paste_imgs_per_s_plan+h2d+kernel": round(batch * rounds / t_all),
            "paste_imgs_per_s_h2d+kernel": round(batch * rounds / (s.elapsed_time(e) * 1e-3))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--layout", default="nhwc", choices=["nhwc", "nchw", "nchw_cl"],
                    help="nhwc: channels-last maps (default); nchw: contiguous maps and RoI features (an unmodified reference "
                         "model); nchw_cl: contiguous maps, channels-last RoI features (_lib.POOLED_CHANNELS_LAST)")
    ap.add_argument("--cpu-rois", type=int, default=192, help="RoIs in the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist

            dist.destroy_process_group()


if __name__ == "__main__":
    main()
