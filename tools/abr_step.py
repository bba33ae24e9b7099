#!/usr/bin/env python
"""BASELINE.json configs[1] -- the VOC 15-5 ABR incremental step under DDP.

    python tools/abr_step.py --arm ours [--steps 10] [--channels-last]              # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/abr_step.py --arm ours

A plain-torch HOST for one step of tools/train_incremental.py:55-181 of the reference: teacher (16 classes, eval, no
grad) and student (21 classes) R-50-C4 Faster R-CNN, batch 4 per GPU of synthetic 800x1216 images with 3 ground-truth
boxes each, ard + id losses (alpha 0.5, beta 1, gamma 1 -- scripts/run_SI.sh:22), SGD step, the student wrapped in
DistributedDataParallel (tools/train_incremental.py:230-235: NCCL gradient all-reduce, broadcast_buffers=False).

The trunk (conv1 .. layer3), the res5 head (layer4) and the linear predictors are torchvision / torch.nn modules --
they are NOT part of the accelerated path.  Everything between them is the REFERENCE's own detection code, imported
unmodified from the staged sources (baseline/_ref, tools/stage_reference.py): AnchorGenerator, RPNPostProcessor,
RPNLossComputation, Matcher, BalancedPositiveNegativeSampler, BoxCoder, Pooler, FastRCNNLossComputation,
calculate_roi_distillation_losses, calculate_attentive_roi_feature_distillation.  Two arms run the SAME host:

  --arm reference   maskrcnn_benchmark._C = the reference's own CUDA kernels (oracle/_ref/libabr_ref_cuda.so, compiled in
                    place for sm_100a) and the reference's Python as it is (per-image NMS loops with host syncs, per-level
                    Pooler loop, per-box torch.cat of the soften pick, PyTorch ARD / id losses through autograd);
  --arm ours        abr_iod_b200.compat.install() + patch_loaded(): the same names now resolve to this library
                    (batched RPN proposals, fused matching / losses, ROIAlign, fused ARD kernel) -- zero edits to the host;
  --arm ours_fused  additionally the distillation RoI work goes through pooled_attentive_roi_distillation (teacher +
                    student pooling + ARD loss + its backward in one call).

Prints one JSON line (rank 0): imgs/s = world * batch / step time (CUDA events, max over ranks), the step time without
the gradient all-reduce (DDP no_sync) and hence the EXPOSED (non-overlapped) all-reduce time.
TEST / MEASUREMENT INFRASTRUCTURE: not imported by the product."""
import argparse
import json
import os
import random
import sys
import time
import types

import numpy as np
import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

OLD_CLASSES = ["c%d" % i for i in range(15)]   # VOC 15-5: 15 old classes ...
NEW_CLASSES = ["n%d" % i for i in range(5)]    # ... + 5 new ones (+ background = 16 teacher / 21 student outputs)


def reference_C():
    """A `maskrcnn_benchmark._C` made of the reference's own CUDA kernels (the five functions the hot path binds,
    csrc/vision.cpp:9-24)."""
    import oracle

    m = types.ModuleType("maskrcnn_benchmark._C")

    def roi_align_forward(inp, rois, scale, ph, pw, ratio):
        return oracle.ref_cuda_roi_align_forward(inp, rois, scale, ph, pw, ratio)

    def roi_align_backward(grad, rois, scale, ph, pw, b, c, h, w, ratio):
        return oracle.ref_cuda_roi_align_backward(grad, rois, scale, ph, pw, b, c, h, w, ratio)

    def nms(dets, scores, thr):
        if dets.numel() == 0:
            return torch.empty((0,), dtype=torch.int64, device="cpu")
        return oracle.ref_cuda_nms(dets, scores, thr)

    m.roi_align_forward, m.roi_align_backward, m.nms = roi_align_forward, roi_align_backward, nms
    return m


def import_reference(arm):
    """The reference's modules on top of the arm's `_C` (see the module docstring).  Returns a namespace."""
    from refmods import reference_root

    root = reference_root()
    if root is None:
        raise SystemExit("reference sources not staged: run tools/stage_reference.py where /root/reference exists")
    sys.path.insert(0, root)
    for name, typ in (("float", float), ("int", int), ("bool", bool)):  # the reference pins numpy 1.21.5 (requirements.txt:8)
        if not hasattr(np, name):                                       # and spells np.float (anchor_generator.py:224)
            setattr(np, name, typ)
    apex, amp = types.ModuleType("apex"), types.ModuleType("apex.amp")
    amp.float_function = lambda f: f
    apex.amp = amp
    sys.modules["apex"], sys.modules["apex.amp"] = apex, amp
    import abr_iod_b200.compat as compat

    if arm == "reference":
        C = reference_C()
        sys.modules["maskrcnn_benchmark._C"] = C
        import maskrcnn_benchmark

        maskrcnn_benchmark._C = C
    else:
        compat.install()
    ns = types.SimpleNamespace()
    import maskrcnn_benchmark.distillation.distillation as distillation
    import maskrcnn_benchmark.modeling.balanced_positive_negative_sampler as sampler
    import maskrcnn_benchmark.modeling.box_coder as box_coder
    import maskrcnn_benchmark.modeling.matcher as matcher
    import maskrcnn_benchmark.modeling.poolers as poolers
    import maskrcnn_benchmark.modeling.roi_heads.box_head.inference as box_inference  # noqa: F401
    import maskrcnn_benchmark.modeling.roi_heads.box_head.loss as box_loss
    import maskrcnn_benchmark.modeling.rpn.anchor_generator as anchor_generator
    import maskrcnn_benchmark.modeling.rpn.inference as rpn_inference
    import maskrcnn_benchmark.modeling.rpn.loss as rpn_loss
    import maskrcnn_benchmark.structures.bounding_box as bounding_box
    import maskrcnn_benchmark.structures.image_list as image_list

    ns.swapped = compat.patch_loaded() if arm != "reference" else []
    ns.distillation, ns.sampler, ns.box_coder, ns.matcher, ns.poolers = distillation, sampler, box_coder, matcher, poolers
    ns.box_loss, ns.anchor_generator, ns.rpn_inference, ns.rpn_loss = box_loss, anchor_generator, rpn_inference, rpn_loss
    ns.BoxList, ns.to_image_list = bounding_box.BoxList, image_list.to_image_list
    return ns


def reference_soften_pick(all_proposals, BoxList):
    """GeneralizedRCNN.generate_soften_proposal's pick as the reference does it (generalized_rcnn.py:125-163): sort by
    objectness, random.sample 64 of the best 128, assemble boxes and scores with one torch.cat per box."""
    picked = []
    for proposals in all_proposals:
        inds = [proposals.get_field("objectness").sort(descending=True)[1]]
        proposals = proposals[inds]
        n = len(proposals)
        bbox, score = proposals.bbox, proposals.get_field("objectness")
        if n < 64:
            index = random.sample(range(0, n, 1), n)
        elif n < 128:
            index = random.sample(range(0, n, 1), 64)
        else:
            index = random.sample(range(0, 128, 1), 64)
        for i, element in enumerate(index):
            if i == 0:
                sel_bbox = bbox[element].view(-1, 4)
                sel_score = score[element].view(-1, 1)
            else:
                sel_bbox = torch.cat((sel_bbox, bbox[element].view(-1, 4)), 0)
                sel_score = torch.cat((sel_score, score[element].view(-1, 1)), 1)
        out = BoxList(sel_bbox.view(-1, 4), proposals.size, proposals.mode)
        out.add_field("objectness", sel_score.view(-1))
        picked.append(out)
    return picked


class FasterRCNNC4(nn.Module):
    """R-50-C4 Faster R-CNN: torchvision trunk / res5 head / linear predictors around the reference's detection modules."""

    def __init__(self, ns, num_classes, n_old, arm, pre_nms=(12000, 6000), post_nms=(2000, 1000)):
        super().__init__()
        import torchvision
        from torchvision.ops.misc import FrozenBatchNorm2d

        r50 = torchvision.models.resnet50(weights=None, norm_layer=FrozenBatchNorm2d)
        for m in r50.modules():  # random weights + frozen identity batch-norm: damp the residual branches so that 16 blocks
            if isinstance(m, torchvision.models.resnet.Bottleneck):  # deep the activations stay finite (same FLOPs)
                m.bn3.weight.fill_(0.2)
        self.stem = nn.Sequential(r50.conv1, r50.bn1, r50.relu, r50.maxpool)
        self.layer1, self.layer2, self.layer3 = r50.layer1, r50.layer2, r50.layer3
        for p in list(self.stem.parameters()) + list(self.layer1.parameters()):  # FREEZE_CONV_BODY_AT 2 (config/defaults.py)
            p.requires_grad_(False)
        self.res5 = r50.layer4                    # ResNet50Conv5ROIFeatureExtractor's head (stride 2 in its first block)
        self.avgpool = nn.AdaptiveAvgPool2d(1)  # roi_box_predictors.py:16
        self.cls_score = nn.Linear(2048, num_classes)
        self.bbox_pred = nn.Linear(2048, num_classes * 4)
        self.rpn_conv = nn.Conv2d(1024, 1024, 3, padding=1)
        self.rpn_cls = nn.Conv2d(1024, 15, 1)
        self.rpn_box = nn.Conv2d(1024, 60, 1)
        for m in (self.rpn_conv, self.rpn_cls, self.rpn_box):
            nn.init.normal_(m.weight, std=0.01)
            nn.init.zeros_(m.bias)
        nn.init.normal_(self.cls_score.weight, std=0.01)
        nn.init.normal_(self.bbox_pred.weight, std=0.001)
        self.ns, self.arm, self.n_old = ns, arm, n_old
        self.anchor_generator = ns.anchor_generator.AnchorGenerator((32, 64, 128, 256, 512), (0.5, 1.0, 2.0), (16,), 0)
        rpn_coder = ns.box_coder.BoxCoder(weights=(1.0, 1.0, 1.0, 1.0))
        RPNPost = ns.rpn_inference.RPNPostProcessor
        self.box_selector_train = RPNPost(pre_nms_top_n=pre_nms[0], post_nms_top_n=post_nms[0], nms_thresh=0.7, min_size=0,
                                          box_coder=rpn_coder, fpn_post_nms_top_n=post_nms[0])
        self.box_selector_test = RPNPost(pre_nms_top_n=pre_nms[1], post_nms_top_n=post_nms[1], nms_thresh=0.7, min_size=0,
                                         box_coder=rpn_coder, fpn_post_nms_top_n=post_nms[1])
        self.rpn_loss = ns.rpn_loss.RPNLossComputation(
            ns.matcher.Matcher(0.7, 0.3, allow_low_quality_matches=True), ns.sampler.BalancedPositiveNegativeSampler(256, 0.5),
            rpn_coder, ns.rpn_loss.generate_rpn_labels)
        self.pooler = ns.poolers.Pooler(output_size=(7, 7), scales=(1.0 / 16,), sampling_ratio=0)
        self.box_loss = ns.box_loss.FastRCNNLossComputation(
            ns.matcher.Matcher(0.5, 0.5, allow_low_quality_matches=False), ns.sampler.BalancedPositiveNegativeSampler(512, 0.25),
            ns.box_coder.BoxCoder(weights=(10.0, 10.0, 5.0, 5.0)), False, "id" if n_old else None,
            OLD_CLASSES if n_old else [])

    # ---- pieces
    def features(self, images):
        return self.layer3(self.layer2(self.layer1(self.stem(images))))

    def rpn(self, image_list, feats, targets=None):
        t = F.relu(self.rpn_conv(feats))
        objectness, box_regression = [self.rpn_cls(t)], [self.rpn_box(t)]
        anchors = self.anchor_generator(image_list, [feats])
        losses = {}
        with torch.no_grad():
            selector = self.box_selector_train if self.training else self.box_selector_test
            proposals = selector(anchors, objectness, box_regression, targets)
        if self.training:
            lo, lb = self.rpn_loss(anchors, objectness, box_regression, targets)
            losses = {"loss_objectness": lo, "loss_rpn_box_reg": lb}
        return proposals, losses

    def head(self, pooled):
        x = self.avgpool(self.res5(pooled)).flatten(1)
        return self.cls_score(x), self.bbox_pred(x)

    def soften_label(self, feats, proposals):
        """ROIBoxHead.calculate_soften_label (box_head.py:60-78)."""
        pooled = self.pooler([feats], proposals)
        logits, reg = self.head(pooled)
        return (logits, reg.view(-1, logits.size(1), 4)), pooled

    # ---- the two roles
    @torch.no_grad()
    def generate_soften_proposal(self, images, image_sizes):
        """GeneralizedRCNN.generate_soften_proposal (generalized_rcnn.py:121-167), teacher in eval mode."""
        image_list = self.ns.to_image_list(images)
        image_list.image_sizes = image_sizes
        feats = self.features(image_list.tensors)
        all_proposals, _ = self.rpn(image_list, feats)
        if self.arm == "reference":
            picked = reference_soften_pick(all_proposals, self.ns.BoxList)
        else:
            from abr_iod_b200.modeling.detector.soften import select_soften_proposals

            picked = select_soften_proposals(all_proposals)
        if self.arm == "ours_fused":
            return None, picked, feats, None  # pooled together with the student's features later
        soften, pooled = self.soften_label(feats, picked)
        return soften, picked, feats, pooled

    def forward(self, images, image_sizes, targets, teacher_out, teacher=None, alpha=0.5, beta=1.0, gamma=1.0):
        """The student's share of one step (train_incremental.py:88-127): detection losses on its own proposals, then the
        head on the teacher's soften proposals, the id distillation loss and the ARD loss.  Returns the total loss."""
        ns = self.ns
        soften, soften_proposals, feats_teacher, pooled_teacher = teacher_out
        image_list = ns.to_image_list(images)
        image_list.image_sizes = image_sizes
        feats = self.features(image_list.tensors)
        proposals, losses = self.rpn(image_list, feats, targets)
        with torch.no_grad():
            sampled = self.box_loss.subsample(proposals, targets)
        pooled = self.pooler([feats], sampled)
        logits, reg = self.head(pooled)
        lc, lb = self.box_loss([logits], [reg])
        losses.update(loss_classifier=lc, loss_box_reg=lb)
        if self.arm == "ours_fused":
            from abr_iod_b200.distillation.distillation import pooled_attentive_roi_distillation
            rois = self.pooler.convert_to_roi_format(soften_proposals)
            pooled_teacher, pooled_student, ard = pooled_attentive_roi_distillation(feats_teacher, feats, rois, (7, 7), 1.0 / 16, 0, gamma)
            with torch.no_grad():
                t_logits, t_reg = teacher.head(pooled_teacher)
                soften = (t_logits, t_reg.view(-1, t_logits.size(1), 4))
            s_logits, s_reg = self.head(pooled_student)
            target = (s_logits, s_reg.view(-1, s_logits.size(1), 4))
        else:
            target, pooled_student = self.soften_label(feats, soften_proposals)
            ard = ns.distillation.calculate_attentive_roi_feature_distillation(pooled_teacher, pooled_student, gamma=gamma)
        dist_loss = alpha * ns.distillation.calculate_roi_distillation_losses(soften, target, dist="id", soften_proposal=None)
        dist_loss = dist_loss + beta * ard
        return sum(losses.values()) + dist_loss


def make_batch(dev, rank, step, batch=4, h=800, w=1216, channels_last=False, BoxList=None):
    g = torch.Generator(device="cpu").manual_seed(1000 * rank + step)
    images = torch.rand((batch, 3, h, w), generator=g).mul_(2.0).sub_(1.0)  # uniform noise, normalised range
    images = images.pin_memory().to(dev, non_blocking=True)
    if channels_last:
        images = images.contiguous(memory_format=torch.channels_last)
    rng = np.random.default_rng(1000 * rank + step)
    targets = []
    for _ in range(batch):
        x1, y1 = rng.uniform(0, w - 300, 3), rng.uniform(0, h - 300, 3)
        boxes = np.stack([x1, y1, x1 + rng.uniform(60, 290, 3), y1 + rng.uniform(60, 290, 3)], 1).astype(np.float32)
        t = BoxList(torch.from_numpy(boxes).to(dev), (w, h), mode="xyxy")
        t.add_field("labels", torch.from_numpy(rng.integers(16, 21, 3)).to(dev))  # new classes 16..20
        targets.append(t)
    return images, [(h, w)] * batch, targets


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arm", default="ours", choices=["reference", "ours", "ours_fused"])
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--channels-last", action="store_true", help="trunk and features in channels-last storage (a one-line "
                    "change of the host: model.to(memory_format=torch.channels_last)); the RoI ops then take their NHWC kernels")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ns = import_reference(args.arm)
    if args.arm != "reference" and args.channels_last:
        from abr_iod_b200 import _lib

        _lib.POOLED_CHANNELS_LAST = True
    torch.manual_seed(0)
    random.seed(rank)
    teacher = FasterRCNNC4(ns, 1 + len(OLD_CLASSES), 0, args.arm).to(dev).eval()
    student = FasterRCNNC4(ns, 1 + len(OLD_CLASSES) + len(NEW_CLASSES), len(OLD_CLASSES), args.arm).to(dev).train()
    if args.channels_last:
        teacher, student = teacher.to(memory_format=torch.channels_last), student.to(memory_format=torch.channels_last)
    for p in teacher.parameters():
        p.requires_grad_(False)
    model = student
    if world > 1:
        model = nn.parallel.DistributedDataParallel(student, device_ids=[local_rank], broadcast_buffers=False)  # train_incremental.py:230-235
    params = [p for p in student.parameters() if p.requires_grad]
    opt = torch.optim.SGD(params, lr=0.001, momentum=0.9, weight_decay=1e-4)
    n_params = sum(p.numel() for p in params)

    def one_step(i, sync=True):
        images, sizes, targets = make_batch(dev, rank, i, args.batch, channels_last=args.channels_last, BoxList=ns.BoxList)
        teacher_out = teacher.generate_soften_proposal(images, sizes)
        ctx = model.no_sync() if (world > 1 and not sync) else _null()
        with ctx:
            loss = model(images, sizes, targets, teacher_out, teacher)
            opt.zero_grad(set_to_none=True)
            loss.backward()
        opt.step()
        return loss

    def timed(n, sync=True):
        for i in range(args.warmup):
            one_step(i, sync)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        s.record()
        for i in range(n):
            loss = one_step(args.warmup + i, sync)
        e.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3 / n
        ms = max(s.elapsed_time(e) / n, wall)  # host-bound steps: the wall clock between the barriers is the step time
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms, float(loss.detach())

    from abr_iod_b200 import _lib

    launches0 = _lib.launch_count() if args.arm != "reference" else 0
    ms, loss = timed(args.steps)
    launches = (_lib.launch_count() - launches0) if args.arm != "reference" else 0
    ms_nosync = timed(args.steps, sync=False)[0] if world > 1 else None
    if rank == 0:
        line = {"metric": "ABR incremental step imgs/s", "arm": args.arm, "value": world * args.batch / (ms * 1e-3), "unit": "imgs/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "scaling": "weak",
                "config": {"workload": "configs[1] VOC 15-5 ABR incremental step: teacher (16 cls) + student (21 cls) R-50-C4 Faster R-CNN, "
                                       "batch %d/GPU of 800x1216 synthetic images, ard+id losses, SGD step, DDP" % args.batch,
                           "channels_last": bool(args.channels_last), "trainable_parameters": n_params,
                           "gradient_bytes_allreduced": 4 * n_params if world > 1 else 0},
                "loss": loss, "swapped_entry_points": len(ns.swapped), "library_launches_per_step": launches / (args.steps + args.warmup),
                "ms_per_step_without_allreduce": ms_nosync,
                "allreduce_exposed_ms": (ms - ms_nosync) if ms_nosync is not None else None}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


if __name__ == "__main__":
    main()
