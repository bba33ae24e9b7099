#!/usr/bin/env python
"""bench.py -- the reference's headline metric on B200: ROIAlign+ARD RoIs/s, forward+backward.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload configs0|configs1_p7|fpn|paste] [--route fused|separate]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

Default workload = BASELINE.json configs[0], the configuration the metric is quoted on: teacher and student R-50-C4
feature maps [2,1024,38,63] fp32 (two 600x1000 VOC images, stride 16), 512 RoIs per image, POOLER_RESOLUTION 14 (the
code default, config/defaults.py:239), adaptive sampling (sampling_ratio 0).  A "step" is one pass of the RoI hot path
over one batch: the composite unit of SURVEY.md section 8d -- teacher ROIAlign forward, student ROIAlign forward, ARD
loss forward+backward (gamma=1) and student ROIAlign backward -- which this library runs as ONE call (abr_roi_ard_fused:
plan, teacher+student pooling with the ARD channel sums in its epilogue, per-RoI coefficients, backward that forms the
ARD gradient on the fly).  Work shards by image, so N GPUs run N independent batches (weak scaling, no data-path
collective).

One JSON line on rank 0: `value` = RoIs/s with inputs resident in HBM (CUDA events over exactly K steps, max over
ranks); `e2e` = the same metric through the public Python API with HOST buffers (pinned H2D of both feature maps and the
RoIs, D2H of the loss and of the student feature-map gradient inside the timed region); `roofline` for the dominant
kernel from CUDA events recorded around the kernels on their stream; `cpu_baseline` = the reference's CPU path on this
box's host cores; `workloads` = the round-1 P=7 line, config 5 (FPN) and `secondary` = configs 3 and 4 with their CPU
baselines.
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

WORKLOADS = {
    # BASELINE.json configs[0] (SURVEY 8d "Config 1"): the quoted configuration
    "configs0": dict(B=2, C=1024, H=38, W=63, rois_per_image=512, P=14, sampling_ratio=0, scale=1.0 / 16, image_w=1000,
                     image_h=600, name="configs[0] ROIAlign 14x14 + ARD, synthetic VOC batch (2 imgs 600x1000, R-50-C4 1024ch "
                                       "stride 16, 512 RoIs/img)"),
    # round-1 headline: configs[1] shapes (VOC 15-5 ABR step, batch 4 per GPU) at the shipped POOLER_RESOLUTION 7
    "configs1_p7": dict(B=4, C=1024, H=50, W=76, rois_per_image=512, P=7, sampling_ratio=0, scale=1.0 / 16, image_w=1216,
                        image_h=800, name="configs[1] shapes: VOC 15-5 ABR step RoI path, batch 4 (800x1216), 512 RoIs/img, P=7"),
}
METRIC = "ROIAlign+ARD RoIs/s fwd+bwd"
UNIT = "RoIs/s"


def make_workload(w, seed=0, rois=None):
    """Synthetic VOC-shaped batch: student = teacher + 0.1*noise; RoI centres uniform in the image,
    sides U(16,400) px, clipped to the image, 5 % degenerate (< 1 px)."""
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


def algorithmic_bytes(w, R):
    """per kernel and for the composite unit (3*BCHW*s + 60R + 6*R*C*P^2*s)."""
    fmap = w["B"] * w["C"] * w["H"] * w["W"] * 4
    pooled = R * w["C"] * w["P"] * w["P"] * 4
    return {"roi_align_fwd": fmap + 20 * R + pooled, "roi_align_bwd": pooled + 20 * R + fmap, "ard": 3 * pooled,
            "composite": 3 * fmap + 60 * R + 6 * pooled, "fmap": fmap, "pooled": pooled}


_JSON_OUT = None  # the real stdout, when fd 1 has been pointed at stderr for the duration of the run (main)


def emit(line):
    """The ONE JSON line of the contract, on the process's original stdout."""
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


INPUT_SETS = 4  # device copies of the (teacher, student) maps used in rotation: 4 x 39 MB = 157 MB > the 126 MB L2


def workload_config(w, route):
    ab = algorithmic_bytes(w, w["B"] * w["rois_per_image"])
    return {"workload": "%s: teacher+student maps [%d,%d,%d,%d] fp32 channels-last, %d RoIs/img, P=%d, sampling_ratio=%d; "
                        "unit = teacher ROIAlign fwd + student ROIAlign fwd + ARD fwd+bwd + student ROIAlign bwd"
                        % (w["name"], w["B"], w["C"], w["H"], w["W"], w["rois_per_image"], w["P"], w["sampling_ratio"]),
            "route": route, "batch_per_gpu": w["B"], "rois_per_gpu": w["B"] * w["rois_per_image"],
            "parallelism": "data-parallel by image",
            "l2": "inputs larger than L2: %d copies of the two feature maps (%.0f MB in total) are used in rotation, so the "
                  "maps a step reads were last touched %d steps (%.0f GB of traffic) earlier; the pooled tensors a step "
                  "writes and re-reads are %.2f GB; no explicit flush"
                  % (INPUT_SETS, INPUT_SETS * 2 * ab["fmap"] / 1e6, INPUT_SETS, INPUT_SETS * ab["composite"] / 1e9,
                     2 * ab["pooled"] / 1e9)}


# ------------------------------------------------------------------------------------------------ clocks / placement
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def summary(self, t0, t1, t_load0=None):
        """Median SM clock and the throttle reasons seen in [t0, t1] (the timed region); with t_load0 also over the whole
        stretch of identical load that starts there (settle loop + warm-up + timed steps) -- a 20-step region is ~20 ms,
        one or two nvidia-smi samples."""
        if self.proc is not None:
            self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1 + 0.05 and len(r) >= 6]
        load = [r for t, r in self.rows if t_load0 is not None and t_load0 + 0.1 <= t <= t1 + 0.05 and len(r) >= 6]
        note = "sampled inside the timed region"
        if not rows:
            rows, note = [r for _, r in self.rows if len(r) >= 6], "no sample fell inside the timed region: all samples of the run"
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in rows)
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in rows)]
        out = {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]), "reasons": reasons, "samples": len(rows), "note": note}
        if load:
            lsm = sorted(float(r[0]) for r in load)
            out["load_phase"] = {"sm_mhz": lsm[len(lsm) // 2], "sm_mhz_min": lsm[0], "samples": len(load),
                                 "reasons": [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in load)],
                                 "note": "settle loop + warm-up + timed steps: the same work, back to back"}
        return out


def bind_to_gpu_numa(index):
    """Pin this process (and the pinned host buffers it allocates afterwards: first touch) to the NUMA node of GPU
    `index`, so that N ranks do not all stage their copies through node 0.  Returns a description for the JSON line."""
    try:
        import torch

        p = torch.cuda.get_device_properties(index)
        bus = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read())
        if node < 0:
            return {"numa_node": None, "note": "the platform reports no NUMA affinity for %s" % bus}
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return {"numa_node": node, "cpus_bound": len(cpus), "pci": bus}
    except Exception as e:  # noqa: BLE001  (placement is best effort; the record says what happened)
        return {"numa_node": None, "note": "not bound: %s" % e}


# ------------------------------------------------------------------------------------------------ CPU reference
def cpu_reference_sample(w, n_rois, threads=None):
    """The reference's CPU path for the composite unit on `n_rois` RoIs of the workload.  ROIAlign forward is the
    reference's own ROIAlign_forward_cpu (oracle/_ref, single-threaded by construction: csrc/cpu/ROIAlign_cpu.cpp:133)
    when it was compiled, else the C port; ARD is the PyTorch op sequence of distillation.py:86-130 on all host
    threads; ROIAlign backward has no CPU reference (csrc/ROIAlign.h:44) and is timed with the C port of the CUDA
    kernel.  Returns (seconds, kind, description, threads)."""
    import torch

    import oracle
    from oracle import ard_torch

    if threads:
        torch.set_num_threads(threads)
    teacher, student, rois = make_workload(w, 0, rois=n_rois)
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
    w = WORKLOADS[args.workload if args.workload in WORKLOADS else "configs0"]
    # calibrate the bounded sample so that K steps take about a minute in total (4..256 RoIs per step)
    cpu_reference_sample(w, 4)
    per_roi = cpu_reference_sample(w, 8)[0] / 8
    sample = int(max(4, min(256, 60.0 / max(args.steps, 1) / per_roi)))
    for _ in range(min(args.warmup, 1)):
        cpu_reference_sample(w, sample)
    t, kind, desc, cores = 0.0, "port", "", 1
    for _ in range(args.steps):
        dt, kind, desc, cores = cpu_reference_sample(w, sample)
        t += dt
    value = sample * args.steps / t
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(w, "reference CPU path"), sample_rois_per_step=sample),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": desc},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ------------------------------------------------------------------------------------------------ the RoI path
class RoiPath:
    """One workload resident on one GPU, runnable on the fused route (ONE C-ABI call per step) or on the separate ops."""

    def __init__(self, w, dev, seed, route):
        import torch

        from abr_iod_b200 import _lib

        self.w, self.dev, self.route, self.torch, self._lib = w, dev, route, torch, _lib
        self.teacher_np, self.student_np, self.rois_np = make_workload(w, seed=seed)
        cl = torch.channels_last
        self.teacher = torch.from_numpy(self.teacher_np).to(dev).contiguous(memory_format=cl)
        self.student = torch.from_numpy(self.student_np).to(dev).contiguous(memory_format=cl)
        # the timed steps rotate over INPUT_SETS copies of the maps (the same values: parity is unaffected), so that no step
        # finds its inputs in L2 from the step before
        self.map_sets = [(self.teacher, self.student)] + [(self.teacher.clone(memory_format=torch.preserve_format),
                                                            self.student.clone(memory_format=torch.preserve_format))
                                                           for _ in range(INPUT_SETS - 1)]
        self.step_no = 0
        self.rois = torch.from_numpy(self.rois_np).to(dev)
        self.R = self.rois_np.shape[0]
        P, C = w["P"], w["C"]
        if route == "fused":
            L = _lib.lib()
            self.f_old = torch.empty((self.R, C, P, P), device=dev).contiguous(memory_format=cl)
            self.f_new = torch.empty_like(self.f_old)
            self.gmap = torch.empty_like(self.student)
            self.loss3 = torch.empty(3, device=dev)
            self.ws_bytes = int(L.abr_roi_ard_fused_workspace_bytes(self.R, C, P, P))
            self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=dev)

    def step(self, events=None):
        w, _lib, torch = self.w, self._lib, self.torch
        P, ratio, scale = w["P"], w["sampling_ratio"], w["scale"]
        teacher, student = self.map_sets[self.step_no % len(self.map_sets)]
        self.step_no += 1
        if self.route == "fused":
            # buffers are reused from step to step; RoIs are re-planned every step (a new batch has new RoIs)
            _lib.check(_lib.lib().abr_roi_ard_fused(
                teacher.data_ptr(), student.data_ptr(), self.rois.data_ptr(), self.f_old.data_ptr(),
                self.f_new.data_ptr(), self.gmap.data_ptr(), self.loss3.data_ptr(), w["B"], w["C"], w["H"], w["W"], self.R,
                P, P, scale, ratio, 1.0, 1.0, _lib.ABR_F32, _lib.ABR_NHWC, 1, self.ws.data_ptr(), self.ws_bytes, 0,
                _lib.stream_ptr(self.dev)))
            return self.loss3, self.gmap
        from abr_iod_b200.distillation.distillation import _ard_launch
        from abr_iod_b200.layers.roi_align import roi_align_backward, roi_align_forward

        marks = []

        def mark():
            if events is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                marks.append(e)
        mark()
        f_old, plan = roi_align_forward(teacher, self.rois, scale, P, P, ratio, return_plan=True)
        mark()
        f_new = roi_align_forward(student, self.rois, scale, P, P, ratio, plan=plan)  # same RoIs: plans reused
        mark()
        loss3, g = _ard_launch(f_old, f_new, 1.0, True)
        mark()
        gin = roi_align_backward(g, self.rois, scale, P, P, w["B"], w["C"], w["H"], w["W"], ratio, layout=_lib.ABR_NHWC, plan=plan)
        mark()
        if events is not None:
            events.append(marks)
        return loss3, gin

    def kernel_bytes(self):
        """Algorithmic bytes attributed to each timed stage of the route."""
        ab = algorithmic_bytes(self.w, self.R)
        if self.route == "fused":
            return {"plan": 20 * self.R,
                    "pool_teacher_student": 2 * ab["roi_align_fwd"],          # teacher fwd + student fwd
                    "ard_coefficients": 0,
                    "backward": ab["ard"] + ab["roi_align_bwd"]}              # ARD fwd+bwd + student bwd
        return {"roi_align_fwd_teacher": ab["roi_align_fwd"], "roi_align_fwd_student": ab["roi_align_fwd"], "ard": ab["ard"],
                "roi_align_bwd": ab["roi_align_bwd"]}

    def timed(self, steps, warmup, barrier):
        """(ms per step, {stage: ms}, launches, t0, t1) over exactly `steps` steps after `warmup` untimed ones."""
        torch, _lib = self.torch, self._lib
        for _ in range(warmup):
            self.step()
        barrier()
        launches0 = _lib.launch_count()
        events = []
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if self.route == "fused":
            # the five event records of a call cost ~15 us of stream time: the stages are timed on every `every`-th step of
            # the timed region (at least five of them), not on all
            self.stage_every = max(1, steps // 5)
            _lib.stage_timing_begin(min(steps, 4096), self.stage_every)
        t0 = time.time()
        start.record()
        for _ in range(steps):
            self.step(events)
        end.record()
        barrier()
        t1 = time.time()
        ms = start.elapsed_time(end) / steps
        if self.route == "fused":
            self.stage_samples, per_kernel = _lib.stage_timing_end()
        else:
            self.stage_every, self.stage_samples = 1, len(events)
            names = ("roi_align_fwd_teacher", "roi_align_fwd_student", "ard", "roi_align_bwd")
            per_kernel = {n: sum(m[i].elapsed_time(m[i + 1]) for m in events) / len(events) for i, n in enumerate(names)}
        return ms, per_kernel, _lib.launch_count() - launches0, t0, t1

    def e2e(self, steps, barrier):
        """The same unit through the public autograd API with HOST buffers: per step pinned H2D of both maps and the
        RoIs, D2H of the loss and of the student map's gradient.  Three streams, steps rotate over them: every step
        pays its own copies, consecutive steps overlap (copy engines both ways + SMs) like a prefetching input
        pipeline (two streams leave no slack: H2D 0.73 + kernels 1.12 + D2H 0.36 ms is 2 x the kernel time).  Returns (ms per step, h2d bytes, d2h bytes)."""
        torch, w, dev = self.torch, self.w, self.dev
        from abr_iod_b200.distillation.distillation import calculate_attentive_roi_feature_distillation as ard
        from abr_iod_b200.distillation.distillation import pooled_attentive_roi_distillation
        from abr_iod_b200.layers import ROIAlign

        P, ratio, scale = w["P"], w["sampling_ratio"], w["scale"]
        cl = torch.channels_last
        h_teacher = torch.from_numpy(self.teacher_np).contiguous(memory_format=cl).pin_memory()
        h_student = torch.from_numpy(self.student_np).contiguous(memory_format=cl).pin_memory()
        h_rois = torch.from_numpy(self.rois_np).pin_memory()
        streams = [torch.cuda.Stream(dev) for _ in range(int(os.environ.get('ABR_E2E_STREAMS', '3')))]
        h_grad = [torch.empty_like(h_student).pin_memory() for _ in streams]
        h_loss = [torch.empty((), dtype=torch.float32).pin_memory() for _ in streams]
        pool = ROIAlign((P, P), scale, ratio)

        def one(i):
            k = i % len(streams)
            with torch.cuda.stream(streams[k]):
                t = h_teacher.to(dev, non_blocking=True)
                s = h_student.to(dev, non_blocking=True).requires_grad_(True)
                r = h_rois.to(dev, non_blocking=True)
                if self.route == "fused":
                    _, _, loss = pooled_attentive_roi_distillation(t, s, r, (P, P), scale, ratio, 1.0)
                else:
                    with torch.no_grad():
                        f_old = pool(t, r)
                    loss = ard(f_old, pool(s, r), 1.0)
                loss.backward()
                h_loss[k].copy_(loss.detach(), non_blocking=True)
                h_grad[k].copy_(s.grad, non_blocking=True)

        for i in range(2 * len(streams)):
            one(i)
        barrier()
        es, ee = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for st_ in streams:
            st_.wait_stream(torch.cuda.current_stream(dev))
        es.record(streams[0])
        for st_ in streams[1:]:
            st_.wait_event(es)
        t_issue = time.perf_counter()
        for i in range(steps):
            one(i)
        t_issue = (time.perf_counter() - t_issue) * 1e3 / steps  # host time to ISSUE a step (Python API, autograd, allocator)
        for st_ in streams[1:]:
            streams[0].wait_stream(st_)
        ee.record(streams[0])
        barrier()
        h2d = int(h_teacher.numel() * 4 + h_student.numel() * 4 + h_rois.numel() * 4)
        d2h = int(h_grad[0].numel() * 4 + 4)
        e2e_ms = es.elapsed_time(ee) / steps
        # what this host's PCIe links sustain for pinned 128 MB copies, measured in the same run: each way alone, and both
        # ways at once (what the pipelined leg asks for); every rank copies at the same time (barrier), so at N GPUs the
        # figures are per GPU under the load of all N
        big_h = [torch.empty(128 << 20, dtype=torch.uint8).pin_memory() for _ in range(2)]
        big_d = [torch.empty(128 << 20, dtype=torch.uint8, device=dev) for _ in range(2)]
        s_up, s_dn = streams[0], streams[1]
        peak = {}

        def copy_rate(up, down):
            best = [0.0, 0.0]
            for _ in range(3):
                barrier()
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
                if up:
                    with torch.cuda.stream(s_up):
                        ev[0].record()
                        big_d[0].copy_(big_h[0], non_blocking=True)
                        ev[1].record()
                if down:
                    with torch.cuda.stream(s_dn):
                        ev[2].record()
                        big_h[1].copy_(big_d[1], non_blocking=True)
                        ev[3].record()
                torch.cuda.synchronize()
                if up:
                    best[0] = max(best[0], big_h[0].numel() / (ev[0].elapsed_time(ev[1]) * 1e-3) / 1e9)
                if down:
                    best[1] = max(best[1], big_h[1].numel() / (ev[2].elapsed_time(ev[3]) * 1e-3) / 1e9)
            return best

        peak["h2d"] = round(copy_rate(True, False)[0], 1)
        peak["d2h"] = round(copy_rate(False, True)[1], 1)
        both = copy_rate(True, True)
        peak["h2d_while_d2h"], peak["d2h_while_h2d"] = round(both[0], 1), round(both[1], 1)
        self.pcie_peak, self.e2e_streams = peak, len(streams)
        return e2e_ms, h2d, d2h, t_issue


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy)"
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


def kernel_table(per_kernel, kbytes, peak):
    return {n: {"ms": round(per_kernel[n], 4), "algorithmic_bytes": kbytes[n],
                "GB/s": round(kbytes[n] / (per_kernel[n] * 1e-3) / 1e9, 1) if per_kernel[n] > 0 else None,
                "frac": round(kbytes[n] / (per_kernel[n] * 1e-3) / 1e9 / peak, 4) if per_kernel[n] > 0 else None}
            for n in per_kernel}


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    placement = bind_to_gpu_numa(local_rank)
    sampler = ClockSampler(local_rank) if rank == 0 else None  # started early: nvidia-smi needs a second to warm up

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.workload in ("fpn", "paste"):
        return run_other_workload(args, rank, world, dev, barrier, sampler)
    w = WORKLOADS[args.workload]
    path = RoiPath(w, dev, seed=rank, route=args.route)
    R = path.R
    warm = max(args.warmup, 3)
    # the driver's default run is short (tens of milliseconds): precede the timed region by ~0.3 s of the same work so
    # that clocks and the clock sampler have settled; these steps are warm-up, outside the K timed steps
    t_settle = time.time()
    while time.time() - t_settle < 0.3:
        path.step()
    ms_per_step, per_kernel, launches, t0, t1 = path.timed(args.steps, warm, barrier)
    clocks = sampler.summary(t0, t1, t_settle) if sampler else None
    e2e_steps = max(4, min(args.steps, 40))
    e2e_ms, h2d, d2h, issue_ms = path.e2e(e2e_steps, barrier)

    if world > 1:
        t = torch.tensor([ms_per_step, e2e_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_per_step, e2e_ms = t.tolist()
        lt = torch.tensor([launches], device=dev, dtype=torch.int64)
        dist.all_reduce(lt)
        launches = int(lt.item())
    if rank != 0:
        return

    peak, peak_src = measured_peak()
    ab = algorithmic_bytes(w, R)
    kbytes = path.kernel_bytes()
    timed_kernels = {k: v for k, v in per_kernel.items() if kbytes[k] > 20 * R}
    dominant = max(timed_kernels, key=timed_kernels.get)
    achieved = kbytes[dominant] / (per_kernel[dominant] * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("%s/%s" % (args.workload, args.route), {}).get(dominant)
    value = world * R / (ms_per_step * 1e-3)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(w, args.route),
        "e2e": {"value": world * R / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "pipelining": "%d streams, consecutive steps overlap" % path.e2e_streams,
                "h2d_GBs_per_gpu": round(h2d / (e2e_ms * 1e-3) / 1e9, 2), "d2h_GBs_per_gpu": round(d2h / (e2e_ms * 1e-3) / 1e9, 2),
                "h2d_GBs_aggregate": round(world * h2d / (e2e_ms * 1e-3) / 1e9, 2), "host_placement": placement,
                "host_issue_ms_per_step": round(issue_ms, 3),
                "pcie_peak_GBs_measured": dict(path.pcie_peak, note="128 MB pinned copies per GPU, all %d ranks copying at the "
                                               "same time; *_while_*: both directions at once" % world),
                "h2d_frac_of_pcie_peak": round(h2d / (e2e_ms * 1e-3) / 1e9 / max(path.pcie_peak["h2d"], 1e-9), 3),
                "copy_GBs_per_gpu_both_ways": round((h2d + d2h) / (e2e_ms * 1e-3) / 1e9, 2),
                "probe_GBs_per_gpu_both_ways": round(path.pcie_peak["h2d_while_d2h"] + path.pcie_peak["d2h_while_h2d"], 1),
                "limiter": ("the device step (copies hidden behind the kernels)" if e2e_ms < 1.15 * ms_per_step else
                            "host issue rate (Python API + autograd per step)" if issue_ms > 0.85 * e2e_ms else
                            "host<->device copies: the host side shared by the ranks (compare copy_GBs_per_gpu_both_ways "
                            "with probe_GBs_per_gpu_both_ways, plain pinned copies under the same N-rank load)")},
        "gpu_launches": launches, "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": dominant, "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                     "frac": round(achieved / peak, 4), "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": kbytes[dominant],
                     "note": "algorithmic bytes of SURVEY 8d attributed to the stage (fused route: pooling kernel = teacher "
                             "fwd + student fwd; backward kernel = ARD fwd+bwd + student bwd, whose pooled-gradient tensor "
                             "is never materialised)"},
        "composite": {"algorithmic_bytes_per_step": ab["composite"],
                      "GB/s": round(ab["composite"] / (ms_per_step * 1e-3) / 1e9, 1),
                      "frac_of_hbm_peak": round(ab["composite"] / (ms_per_step * 1e-3) / 1e9 / peak, 4)},
        "kernels": kernel_table(per_kernel, kbytes, peak),
        "kernel_timing": {"events": "CUDA events recorded by the library on the caller's stream between the stages, inside the "
                                    "timed region", "every_nth_step": path.stage_every, "steps_sampled": path.stage_samples},
    }
    if world == 1 and not args.no_cpu:
        dt, kind, desc, cores = cpu_reference_sample(w, args.cpu_rois)
        line["cpu_baseline"] = {"value": args.cpu_rois / dt, "unit": UNIT, "cores": cores, "kind": kind, "sample": desc}
    if world == 1 and not args.no_secondary:
        extra = {}
        other = "configs1_p7" if args.workload == "configs0" else "configs0"
        for route in ("fused", "separate"):
            p2 = RoiPath(WORKLOADS[other], dev, seed=0, route=route)
            ms2, pk2, _, _, _ = p2.timed(min(args.steps, 50), 3, barrier)
            ab2 = algorithmic_bytes(p2.w, p2.R)
            extra["%s/%s" % (other, route)] = {
                "workload": WORKLOADS[other]["name"], "value": p2.R / (ms2 * 1e-3), "unit": UNIT, "ms_per_step": ms2,
                "composite_frac_of_hbm_peak": round(ab2["composite"] / (ms2 * 1e-3) / 1e9 / peak, 4),
                "kernels": kernel_table(pk2, p2.kernel_bytes(), peak)}
            del p2
        if args.route == "fused":
            p3 = RoiPath(w, dev, seed=0, route="separate")
            ms3, pk3, _, _, _ = p3.timed(min(args.steps, 50), 3, barrier)
            extra["%s/separate" % args.workload] = {
                "workload": w["name"], "value": p3.R / (ms3 * 1e-3), "unit": UNIT, "ms_per_step": ms3,
                "composite_frac_of_hbm_peak": round(ab["composite"] / (ms3 * 1e-3) / 1e9 / peak, 4),
                "kernels": kernel_table(pk3, p3.kernel_bytes(), peak)}
            del p3
        torch.cuda.empty_cache()
        extra["configs4_fpn"] = fpn_metrics(dev, peak, steps=min(args.steps, 50))
        extra["ard_call_of_the_real_step"] = small_call_metrics(dev)
        line["workloads"] = extra
        line["reference_cuda"] = reference_cuda_metrics(w, dev)
        torch.cuda.empty_cache()
        line["secondary"] = secondary_metrics(dev)
    emit(line)


def small_call_metrics(dev):
    """The distillation call at the size the real step makes it (batch 4, 64 soften RoIs per image = 256 RoIs,
    P = 7, maps [4,1024,50,76]): one abr_roi_ard_fused call issued from Python vs. the same call replayed from a CUDA
    graph (the library only enqueues work on the caller's stream, so the whole call is capturable) -- at this size the
    host's issue time matters as much as the kernels."""
    w = dict(WORKLOADS["configs1_p7"], rois_per_image=64)
    path = RoiPath(w, dev, seed=0, route="fused")
    t_py = _time_call(path.step, graph=False, reps=50)
    t_graph = _time_call(path.step, graph=True, reps=50)
    return {"workload": "abr_roi_ard_fused, %d RoIs (64/img), P=7, maps [4,1024,50,76]" % path.R,
            "us_per_call_issued_from_python": round(t_py * 1e6, 1), "us_per_call_cuda_graph_replay": round(t_graph * 1e6, 1),
            "RoIs_per_s_cuda_graph": round(path.R / t_graph)}


# ------------------------------------------------------------------------------------------------ reference kernels, same box
def reference_cuda_metrics(w, dev, steps=10):
    """The reference's own CUDA kernels recompiled as they are for sm_100a (oracle/_ref/libabr_ref_cuda.so: csrc/cuda/
    ROIAlign_cuda.cu, nms.cu) and its PyTorch ARD (distillation.py:86-130, autograd) on THIS B200, on the same
    tensors in the reference's layout (contiguous NCHW): "the bar is the reference kernels recompiled as-is" (SURVEY 2.2).
    The comparator is test infrastructure; the product never links it."""
    import torch

    import oracle
    from oracle import ard_torch

    if not oracle.ref_cuda_available():
        return {"unavailable": "oracle/_ref/libabr_ref_cuda.so was not built (needs /root/reference at build time)"}
    L = oracle.ref_cuda_lib()
    teacher_np, student_np, rois_np = make_workload(w, seed=0)
    t = torch.from_numpy(teacher_np).to(dev)
    s = torch.from_numpy(student_np).to(dev)
    r = torch.from_numpy(rois_np).to(dev)
    B, C, H, W, P, ratio, scale = w["B"], w["C"], w["H"], w["W"], w["P"], w["sampling_ratio"], w["scale"]
    R = r.shape[0]
    f_old = oracle.ref_cuda_roi_align_forward(t, r, scale, P, P, ratio)
    f_new = oracle.ref_cuda_roi_align_forward(s, r, scale, P, P, ratio)
    _, g = ard_torch.ard_fwd_bwd(f_old, f_new, 1.0)

    def timeit(fn):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / steps

    out = {"roi_align_forward_ms": timeit(lambda: L.ref_cuda_roi_align_forward_nocopy(t.data_ptr(), r.data_ptr(), B, C, H, W, R, P, P, scale, ratio)),
           "roi_align_backward_ms": timeit(lambda: L.ref_cuda_roi_align_backward_nocopy(g.data_ptr(), r.data_ptr(), B, C, H, W, R, P, P, scale, ratio)),
           "ard_pytorch_fwd_bwd_ms": timeit(lambda: ard_torch.ard_fwd_bwd(f_old, f_new, 1.0))}
    step_ms = 2 * out["roi_align_forward_ms"] + out["ard_pytorch_fwd_bwd_ms"] + out["roi_align_backward_ms"]
    out.update({"step_ms": step_ms, "value": R / (step_ms * 1e-3), "unit": UNIT,
                "what": "2 x ROIAlign_forward_cuda + PyTorch ARD forward+backward (autograd) + ROIAlign_backward_cuda, contiguous NCHW fp32, "
                        "the reference's launch shapes, device-timed on this GPU"})
    rng = np.random.default_rng(3)
    b6, s6 = rpn_shaped_boxes(rng, 6000)
    bt, st = torch.from_numpy(b6).to(dev), torch.from_numpy(s6).to(dev)
    ms = timeit(lambda: oracle.ref_cuda_nms(bt, st, 0.7))
    out["nms_cuda_boxes_per_s_n6000_b1"] = round(6000 / (ms * 1e-3))
    return out


# ------------------------------------------------------------------------------------------------ config 5 (FPN)
FPN = dict(B=2, C=256, image_w=1344, image_h=800, scales=(0.25, 0.125, 0.0625, 0.03125), rois_per_image=512, P=7, sampling_ratio=2)


def make_fpn(dev, seed):
    import torch

    from abr_iod_b200.structures.bounding_box import BoxList

    f = FPN
    rng = np.random.default_rng(100 + seed)
    feats = [torch.randn(f["B"], f["C"], int(f["image_h"] * s), int(f["image_w"] * s), device=dev).contiguous(
        memory_format=torch.channels_last).requires_grad_(True) for s in f["scales"]]
    boxes = []
    for _ in range(f["B"]):
        n = f["rois_per_image"]
        side = np.exp(rng.uniform(np.log(16), np.log(800), n))          # log-uniform 16..800 px: all four levels are hit
        aspect = rng.uniform(0.5, 2.0, n)
        bw, bh = side * np.sqrt(aspect), side / np.sqrt(aspect)
        cx, cy = rng.uniform(0, f["image_w"], n), rng.uniform(0, f["image_h"], n)
        b = np.stack([np.clip(cx - bw / 2, 0, f["image_w"] - 1), np.clip(cy - bh / 2, 0, f["image_h"] - 1),
                      np.clip(cx + bw / 2, 0, f["image_w"] - 1), np.clip(cy + bh / 2, 0, f["image_h"] - 1)], 1).astype(np.float32)
        boxes.append(BoxList(torch.from_numpy(b).to(dev), (f["image_w"], f["image_h"]), "xyxy"))
    return feats, boxes


def fpn_step_fn(dev, seed):
    import torch

    from abr_iod_b200.modeling.poolers import Pooler

    f = FPN
    feats, boxes = make_fpn(dev, seed)
    pooler = Pooler((f["P"], f["P"]), f["scales"], f["sampling_ratio"])
    R = f["B"] * f["rois_per_image"]
    gout = torch.randn(R, f["C"], f["P"], f["P"], device=dev).contiguous(memory_format=torch.channels_last)

    def step():
        for x in feats:
            x.grad = None
        out = pooler(feats, boxes)
        out.backward(gout)
        return out

    maps = sum(x.numel() for x in feats) * 4
    pooled = R * f["C"] * f["P"] * f["P"] * 4
    return step, R, 2 * (maps + 20 * R + pooled)  # SURVEY 8d multi-level: fwd + bwd, sum over levels of B*C*H_l*W_l*s


def fpn_metrics(dev, peak, steps=50):
    """BASELINE.json configs[4] on one GPU's shard: 2 images 800x1344, 4-level FPN maps (256 ch),
    512 RoIs per image with log-uniform sizes, P=7, sampling_ratio 2; Pooler forward + backward through the public API
    (LevelMapper on the device, all levels in one launch per direction)."""
    import torch

    step, R, nbytes = fpn_step_fn(dev, 0)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(steps):
        step()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / steps
    graph_ms = None
    try:  # the same forward + backward replayed from a CUDA graph: the device time without the Python / autograd issue gaps
        graph_ms = _time_call(step, graph=True, reps=steps) * 1e3
    except Exception as ex:  # noqa: BLE001
        graph_ms = "capture failed: %s" % str(ex)[:80]
    return {"workload": "configs[4] COCO-shape FPN multi-level ROIAlign 7x7 fwd+bwd: levels [2,256,200,336] .. [2,256,25,42], "
                        "512 RoIs/img, sampling_ratio 2, per-GPU shard of batch 16 / 8 GPUs",
            "value": R / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "ms_per_step_cuda_graph": graph_ms,
            "algorithmic_bytes_per_step": nbytes,
            "roofline": {"bound": "hbm", "achieved": round(nbytes / (ms * 1e-3) / 1e9, 1), "peak": peak, "unit": "GB/s",
                         "frac": round(nbytes / (ms * 1e-3) / 1e9 / peak, 4),
                         "note": "includes the Python wrapper, LevelMapper kernel, plan kernel and the zero-fill of the four "
                                 "gradient maps (183 MB); the call is launch/latency-bound at this size"}}


def run_other_workload(args, rank, world, dev, barrier, sampler):
    """--workload fpn / paste under torchrun: every rank runs its own shard (no data-path collective), RoIs/s resp.
    imgs/s summed over ranks (weak scaling)."""
    import torch
    import torch.distributed as dist

    peak, _ = measured_peak()
    if args.workload == "fpn":
        step, units, nbytes = fpn_step_fn(dev, rank)
        metric, unit = "FPN multi-level ROIAlign 7x7 RoIs/s fwd+bwd", UNIT
    else:
        step, units = paste_step_fn(dev, rank)
        nbytes, metric, unit = None, "ABR paste imgs/s (plan + H2D + one launch)", "imgs/s"
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    s.record()
    for _ in range(args.steps):
        step()
    e.record()
    barrier()
    t1 = time.time()
    ms = s.elapsed_time(e) / args.steps
    if args.workload == "paste":
        ms = (t1 - t0) * 1e3 / args.steps  # host planning is part of the call: wall clock between the barriers
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    if rank != 0:
        return
    line = {"metric": metric, "value": world * units / (ms * 1e-3), "unit": unit, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.workload == "fpn" else "u8", "data": "synthetic",
            "config": {"workload": args.workload, "units_per_gpu_per_step": units}, "clocks": sampler.summary(t0, t1) if sampler else None}
    if nbytes:
        line["roofline"] = {"bound": "hbm", "achieved": round(nbytes / (ms * 1e-3) / 1e9, 1), "peak": peak, "unit": "GB/s",
                            "frac": round(nbytes / (ms * 1e-3) / 1e9 / peak, 4)}
    emit(line)


# ------------------------------------------------------------------------------------------------ configs 3 and 4
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


def rpn_shaped_boxes(rng, n, H=38, W=63, stride=16):
    """anchors of a 38x63 grid x 15 anchors perturbed by N(0, 0.1) deltas, decoded and clipped;
    scores sigmoid(N(0, 2)); the n best by score, in descending score order (what the RPN's sorted top-k hands to NMS)."""
    from inputs import make_anchors

    anchors = make_anchors(H, W, stride)
    d = rng.normal(0, 0.1, anchors.shape).astype(np.float32)
    wa, ha = anchors[:, 2] - anchors[:, 0] + 1, anchors[:, 3] - anchors[:, 1] + 1
    cx, cy = anchors[:, 0] + 0.5 * wa + d[:, 0] * wa, anchors[:, 1] + 0.5 * ha + d[:, 1] * ha
    w_, h_ = wa * np.exp(d[:, 2]), ha * np.exp(d[:, 3])
    boxes = np.stack([cx - 0.5 * w_, cy - 0.5 * h_, cx + 0.5 * w_ - 1, cy + 0.5 * h_ - 1], 1)
    boxes = np.clip(boxes, 0, [W * stride - 1, H * stride - 1, W * stride - 1, H * stride - 1]).astype(np.float32)
    scores = (1.0 / (1.0 + np.exp(-rng.normal(0, 2, len(boxes))))).astype(np.float32)
    order = np.argsort(-scores, kind="stable")[:n]
    return np.ascontiguousarray(boxes[order]), np.ascontiguousarray(scores[order])


def nms_metrics(dev):
    """BASELINE.json configs[2]: RPN proposal NMS, N = 6000 boxes per image, IoU 0.7, post-NMS cut
    2000, batch 1..16 (and N = 12000 -> 2000); boxes/s = batch * N / time including order check and compaction.
    Beside it the reference's nms_cpu (csrc/cpu/nms_cpu.cpp, compiled in oracle/_ref) on one host core, per image, and
    the keep lists of both compared in this run."""
    import torch

    import oracle
    from abr_iod_b200.layers import nms_batched

    rng = np.random.default_rng(3)
    out = {"note": "boxes/s counts all N boxes per image; with a post-NMS cut and score-sorted input the kernels first run "
                   "a prefix pass over the 2*max_proposals best boxes and skip the full pass for images it settles "
                   "(DESIGN 3.3); nms_cpu always looks at all N"}
    use_ref = oracle.ref_available()
    for n, keep_n in ((6000, 2000), (12000, 2000)):
        images = [rpn_shaped_boxes(rng, n) for _ in range(16)]
        # CPU baseline: per image, single thread (nms_cpu has no batch form; the RPN calls it once per image)
        t0 = time.perf_counter()
        cpu_keep = [oracle.nms(b, s, 0.7, "cpu", use_ref=use_ref) for b, s in images[:2]]
        cpu_s = (time.perf_counter() - t0) / 2
        out["nms_cpu_boxes_per_s_n%d" % n] = {"value": round(n / cpu_s), "unit": "boxes/s", "cores": 1, "per_image_ms": round(cpu_s * 1e3, 2),
                                              "kind": "reference" if use_ref else "port"}
        boxes = [torch.from_numpy(b).to(dev) for b, _ in images]
        scores = [torch.from_numpy(s).to(dev) for _, s in images]
        keep, n_keep = nms_batched(boxes[:2], scores[:2], 0.7, keep_n, cpu_tie_rule=True)  # nms_cpu's '>=' rule
        exact = all(np.array_equal(keep[i, : int(n_keep[i])].cpu().numpy(), cpu_keep[i][:keep_n]) for i in range(2))
        out["nms_keep_lists_equal_nms_cpu_n%d" % n] = bool(exact)
        for batch in (1, 2, 4, 8, 16):
            key = "nms_boxes_per_s_n%d_b%d_keep%d" % (n, batch, keep_n)
            call = lambda: nms_batched(boxes[:batch], scores[:batch], 0.7, keep_n)  # noqa: E731
            out[key] = round(batch * n / _time_call(call, graph=False))
            out[key + "_graph"] = round(batch * n / _time_call(call, graph=True))
    # unsorted input (the box head's per-class segments): the rank sort runs first
    from inputs import make_boxes

    data = [make_boxes(rng, 6000, 1216, 800) for _ in range(4)]
    boxes = [torch.from_numpy(b).to(dev) for b, _ in data]
    scores = [torch.from_numpy(s).to(dev) for _, s in data]
    out["nms_boxes_per_s_n6000_b4_keep1000_unsorted_graph"] = round(4 * 6000 / _time_call(lambda: nms_batched(boxes, scores, 0.7, 1000), graph=True))
    return out


def secondary_metrics(dev):
    """BASELINE.json's other two numbers, measured in the same run: batched RPN NMS boxes/s (config 3) and ABR paste
    imgs/s (config 4, one GPU's shard), each with its CPU baseline, plus the rows either side of the path."""
    out = {}
    out.update(nms_metrics(dev))
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


def make_paste_inputs(seed, n_proto=2000, batch=16):
    """Config 4 inputs: 2000 synthetic prototypes (uint8 RGB, sides U(71,300)), 16 images 375x500 with 3 GT boxes each."""
    from PIL import Image

    rng = np.random.default_rng(4 + seed)
    protos = []
    for i in range(n_proto):
        h, w = int(rng.integers(71, 301)), int(rng.integers(71, 301))
        protos.append(("%d_%05d.jpg" % (int(rng.integers(1, 16)), i),
                       np.broadcast_to(rng.integers(0, 256, (1, 1, 3), dtype=np.uint8), (h, w, 3)).copy()))
    images = [Image.fromarray(rng.integers(0, 256, (375, 500, 3), dtype=np.uint8)) for _ in range(batch)]
    targets = []
    for _ in range(batch):
        x1, y1 = rng.uniform(0, 250, 3), rng.uniform(0, 180, 3)
        targets.append(np.stack([x1, y1, x1 + rng.uniform(30, 200, 3), y1 + rng.uniform(30, 150, 3), rng.integers(16, 21, 3)], 1))
    return protos, images, targets


def paste_step_fn(dev, seed, batch=16):
    import random

    import torch

    from abr_iod_b200.data.abr_paste import BoxRehearsalPaster

    protos, images, targets = make_paste_inputs(seed, batch=batch)
    paster = BoxRehearsalPaster(protos, batch_size=batch, device=dev)
    random.seed(seed)
    torch.manual_seed(seed)

    def step():
        paster.execute([paster.plan_transform(im, g) for im, g in zip(images, targets)])

    return step, batch


def _reference_voc_abr():
    """The reference's own PascalVOCDataset_ABR class from the staged sources (baseline/_ref, tools/stage_reference.py),
    loaded by file path with the imports this image lacks stubbed (as tests/golden/make_golden.py does); None when the
    sources are not staged."""
    import importlib.util
    import types

    from refmods import reference_root

    root = reference_root()
    if root is None:
        return None
    if root not in sys.path:
        sys.path.insert(0, root)
    apex, amp = types.ModuleType("apex"), types.ModuleType("apex.amp")
    amp.float_function = lambda f: f
    apex.amp = amp
    sys.modules.setdefault("apex", apex)
    sys.modules.setdefault("apex.amp", amp)
    C = types.ModuleType("maskrcnn_benchmark._C")
    sys.modules.setdefault("maskrcnn_benchmark._C", C)
    data, tr = types.ModuleType("maskrcnn_benchmark.data"), types.ModuleType("maskrcnn_benchmark.data.transforms")
    tr.Compose = object
    data.transforms = tr
    sys.modules["maskrcnn_benchmark.data"], sys.modules["maskrcnn_benchmark.data.transforms"] = data, tr
    tools, em = types.ModuleType("tools"), types.ModuleType("tools.extract_memory")
    em.Mem = object
    tools.extract_memory = em
    sys.modules["tools"], sys.modules["tools.extract_memory"] = tools, em
    spec = importlib.util.spec_from_file_location("ref_voc_abr", os.path.join(root, "maskrcnn_benchmark/data/datasets/voc_abr.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _paste_reference_worker(args):
    """The reference's own transform_current_data_with_ABR (data/datasets/voc_abr.py:821-858, with _start_mixup /
    _start_boxes_mosaic and PIL file reads of the prototypes, as its DataLoader workers run it) on `rounds` batches.
    Returns seconds, or None when the reference sources are not staged."""
    import random
    import tempfile
    import types

    import torch
    from PIL import Image

    seed, rounds, batch = args
    torch.set_num_threads(1)
    mod = _reference_voc_abr()
    if mod is None:
        return None
    from maskrcnn_benchmark.structures.bounding_box import BoxList

    protos, images, targets = make_paste_inputs(0, batch=batch)
    tmp = tempfile.mkdtemp(prefix="abr_protos_")
    for name, arr in protos:  # file-per-box memory, as tools/extract_memory.py:213-236 writes it (PNG: lossless, same decode cost class)
        Image.fromarray(arr).save(os.path.join(tmp, name.replace(".jpg", ".png")))
    ds = mod.PascalVOCDataset_ABR.__new__(mod.PascalVOCDataset_ABR)
    ds.PrototypeBoxSelection = types.SimpleNamespace(current_mem_path=tmp, first_mem_path=tmp)
    ds.BoxRehearsal_path = [n.replace(".jpg", ".png") for n, _ in protos]
    ds.boxes_index = list(range(len(protos)))
    ds.batch_size, ds.bg_size = batch, 0
    tl = []
    for im, g in zip(images, targets):
        t = BoxList(torch.tensor(g[:, :4]), im.size, mode="xyxy")
        t.add_field("labels", torch.tensor(g[:, 4]).long())
        tl.append(t)
    random.seed(seed)
    torch.manual_seed(seed)
    t0 = time.perf_counter()
    for _ in range(rounds):
        for im, t in zip(images, tl):
            ds.transform_current_data_with_ABR(im, t)
    return time.perf_counter() - t0


def _paste_cpu_worker(args):
    """One worker of the CPU baseline: the numpy restatement of voc_abr.py's transform on `rounds` batches."""
    import random

    import torch

    from oracle import paste as opaste

    seed, rounds, batch = args
    torch.set_num_threads(1)
    protos, images, targets = make_paste_inputs(0, batch=batch)
    st = opaste.BoxRehearsalState([(n, opaste.as_pil(a)) for n, a in protos], batch)
    random.seed(seed)
    torch.manual_seed(seed)
    arrays = [np.asarray(im) for im in images]
    t0 = time.perf_counter()
    for _ in range(rounds):
        for a, g in zip(arrays, targets):
            opaste.transform_current_data_with_abr(st, opaste.as_pil(a), g)
    return time.perf_counter() - t0


def paste_metrics(dev, batch=16, rounds=8):
    """BASELINE.json configs[3] on one GPU's shard: 2000 synthetic prototypes, batches of 16 images
    375x500 at the reference's 25/25/50 mixup/mosaic/untouched policy.  imgs/s for the whole call (host planning in
    the reference's draw order + pinned H2D + ONE paste launch) and for H2D + kernel alone; beside it the CPU path
    (numpy restatement of voc_abr.py:555-858, oracle/paste.py) on 1 core and on 4 worker processes (the reference runs
    the transform inside 4 DataLoader workers, config/defaults.py:83)."""
    import multiprocessing as mp
    import random

    import torch

    from abr_iod_b200.data.abr_paste import BoxRehearsalPaster

    protos, images, targets = make_paste_inputs(0, batch=batch)
    paster = BoxRehearsalPaster(protos, batch_size=batch, device=dev)
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
    s.record()
    for plans in plans_all:  # kernel + H2D only (plans precomputed)
        paster.execute(plans)
    e.record()
    torch.cuda.synchronize()
    out = {"paste_imgs_per_s_plan+h2d+kernel": round(batch * rounds / t_all),
           "paste_imgs_per_s_h2d+kernel": round(batch * rounds / (s.elapsed_time(e) * 1e-3))}
    t1 = _paste_cpu_worker((0, rounds, batch))
    out["paste_cpu_imgs_per_s_1core"] = {"value": round(batch * rounds / t1), "unit": "imgs/s", "cores": 1, "kind": "port",
                                         "sample": "%d batches of %d images, numpy restatement of voc_abr.py (oracle/paste.py)" % (rounds, batch)}
    try:
        ctx = mp.get_context("spawn")
        with ctx.Pool(4) as pool:
            pool.map(_paste_cpu_worker, [(i, 1, batch) for i in range(4)])  # start-up and imports outside the timing
            tw = max(pool.map(_paste_cpu_worker, [(i, rounds, batch) for i in range(4)]))  # the slowest worker's own clock
            out["paste_cpu_imgs_per_s_4workers"] = {"value": round(4 * batch * rounds / tw), "unit": "imgs/s", "cores": 4, "kind": "port"}
            # the reference's own code (staged sources): 1 core here, 4 workers in the pool
            tr1 = _paste_reference_worker((0, rounds, batch))
            if tr1 is not None:
                out["paste_reference_imgs_per_s_1core"] = {
                    "value": round(batch * rounds / tr1), "unit": "imgs/s", "cores": 1, "kind": "reference",
                    "sample": "%d batches of %d images through PascalVOCDataset_ABR.transform_current_data_with_ABR "
                              "(voc_abr.py:821-858), prototypes read from one file per box" % (rounds, batch)}
                tw = max(pool.map(_paste_reference_worker, [(i, rounds, batch) for i in range(4)]))  # (set-up is outside each worker's clock)
                out["paste_reference_imgs_per_s_4workers"] = {"value": round(4 * batch * rounds / tw), "unit": "imgs/s", "cores": 4,
                                                              "kind": "reference"}
    except Exception as ex:  # noqa: BLE001
        out["paste_cpu_imgs_per_s_4workers"] = {"value": None, "note": "worker pool failed: %s" % ex}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="configs0", choices=["configs0", "configs1_p7", "fpn", "paste"],
                    help="configs0: BASELINE.json configs[0], P=14 (default, the quoted configuration); configs1_p7: the round-1 "
                         "headline shapes; fpn / paste: configs[4] / configs[3] shards for multi-GPU runs")
    ap.add_argument("--route", default="fused", choices=["fused", "separate"],
                    help="fused: abr_roi_ard_fused, one call per step (default); separate: the four ops one by one")
    ap.add_argument("--cpu-rois", type=int, default=64, help="RoIs in the bounded CPU-baseline sample")
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
        # stdout carries the ONE JSON line and nothing else: libraries that write to fd 1 themselves (NCCL prints its version
        # banner there) are sent to stderr for the duration of the run
        global _JSON_OUT
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
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
