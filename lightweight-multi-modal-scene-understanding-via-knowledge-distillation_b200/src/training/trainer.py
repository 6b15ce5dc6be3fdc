"""Training loop (drop-in for the reference's ``src/training/trainer.py``) with the
teacher->student distillation step, on-device metrics and data-parallel training.

Same public surface as the reference -- ``SegmentationMetrics`` and
``Trainer(model, train_loader, val_loader, device, lr, weight_decay, save_dir,
class_weights, num_epochs)`` with ``train / train_epoch / validate /
save_checkpoint / load_checkpoint / update_history`` and the same checkpoint and
history formats -- plus keyword-only extensions (``teacher``, ``kd_*``,
``amp_dtype``).  What changes underneath the reference's step (trainer.py:81-93):

  * loss: one fused kernel gives CE (+ KL + feature-mimic MSE when a teacher is
    given) AND the gradients w.r.t. logits / mimic taps, so the backward starts
    from those tensors directly;
  * no ``loss.item()`` per step and no ``.cpu()`` + per-pixel Python loop for the
    metrics: loss and the confusion matrix accumulate on the device and are read
    once per epoch;
  * AdamW is one kernel over a flat buffer; under torch.distributed the gradients
    are all-reduced as that one flat bucket over NCCL.
"""
from __future__ import annotations

import json
import os
from typing import Dict, Optional

import numpy as np
import torch
import torch.distributed as dist
import torch.nn as nn

from .. import native as _native
from .. import ops
from .optim import FlatAdamW
from .parallel import allreduce_gradients_, world as _world

try:  # progress bars are cosmetic
    from tqdm import tqdm
except ImportError:  # pragma: no cover
    def tqdm(it, **_):
        return it

MIMIC_TAPS = ("lidar_feat", "camera_feat")        # fusion_module.py:260-262 intermediates


class SegmentationMetrics:
    """Confusion-matrix mIoU with the reference's interface (trainer.py:9-37).
    CUDA inputs are counted by ``kdf_confusion_matrix`` into a device-resident
    matrix (no host sync until ``compute()``); ``confusion`` stays a numpy view for
    code that reads it."""

    def __init__(self, num_classes=2, ignore_index=-1):
        self.num_classes = num_classes
        self.ignore_index = ignore_index
        self.reset()

    def reset(self):
        self._host = np.zeros((self.num_classes, self.num_classes), dtype=np.int64)
        self._dev: Optional[torch.Tensor] = None

    def update(self, preds, targets):
        if not preds.is_cuda:
            # host tensors (reference tooling evaluates on the CPU, trainer.py:18-26) are counted by the same kernel:
            # they are copied to the device first -- there is no host implementation of the count
            if not torch.cuda.is_available():
                raise RuntimeError("SegmentationMetrics.update counts on a CUDA device (no CPU fallback)")
            dev = self._dev.device if self._dev is not None else torch.device("cuda", torch.cuda.current_device())
            preds, targets = preds.to(dev), targets.to(dev)
        if preds.shape[1] > self.num_classes:
            # reference semantics: predictions outside [0, num_classes) are skipped (trainer.py:25)
            pred_idx = preds.argmax(dim=1)
            keep = pred_idx < self.num_classes
            targets = torch.where(keep, targets, torch.full_like(targets, self.ignore_index))
            preds = preds[:, :self.num_classes]
        if self._dev is None or self._dev.device != preds.device:
            self._dev = torch.zeros(self.num_classes, self.num_classes, dtype=torch.int64, device=preds.device)
        ops.confusion_matrix_(self._dev, preds.float() if preds.dtype not in (torch.float32, torch.bfloat16) else preds,
                              targets, self.ignore_index)

    @property
    def confusion(self) -> np.ndarray:
        if self._dev is not None:
            return self._host + self._dev.cpu().numpy()
        return self._host

    def all_reduce(self):
        if self._dev is not None and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self._dev)

    def compute(self):
        conf = self.confusion
        ious = []
        for i in range(self.num_classes):
            tp = conf[i, i]
            denom = conf[:, i].sum() + conf[i, :].sum() - tp
            ious.append(tp / denom if denom > 0 else 0.0)
        return {"class_iou": ious, "miou": float(np.mean(ious))}


class _Criterion(nn.Module):
    """``nn.CrossEntropyLoss(ignore_index=-1, weight=w)`` (trainer.py:55) on the fused loss kernel."""

    def __init__(self, weight: Optional[torch.Tensor], ignore_index: int = -1):
        super().__init__()
        self.ignore_index = ignore_index
        self.register_buffer("weight", weight)

    def forward(self, logits, target):
        loss, _ = ops.KDLossFn.apply(logits, None, target, self.weight, 1.0, 0.0, 0.0, self.ignore_index)
        return loss


class Trainer:
    def __init__(self, model, train_loader, val_loader, device,
                 lr=1e-3, weight_decay=1e-3, save_dir="checkpoints",
                 class_weights=None, num_epochs=20, *,
                 teacher: Optional[nn.Module] = None, kd_temperature: float = 4.0,
                 kd_alpha: float = 0.5, kd_beta: float = 1.0,
                 amp_dtype: Optional[torch.dtype] = None, verbose: bool = True,
                 use_cuda_graph: bool = False, graph_warmup_steps: int = 3, overlap_teacher: bool = True,
                 overlap_allreduce: bool = True):
        self.model = model
        self.train_loader = train_loader
        self.val_loader = val_loader
        self.device = torch.device(device)
        self.num_epochs = num_epochs
        self.rank, self.world_size = _world()
        self.verbose = verbose and self.rank == 0
        if self.device.type != "cuda":
            raise RuntimeError("this Trainer drives the B200 kernels and needs a CUDA device (no CPU fallback)")

        if class_weights is not None:
            class_weights = torch.tensor(class_weights, dtype=torch.float32, device=self.device)
            if self.verbose:
                print(f"Using class weights: {class_weights.tolist()}")
        self.class_weights = class_weights
        self.criterion = _Criterion(class_weights, ignore_index=-1).to(self.device)
        self.optimizer = FlatAdamW(model.parameters(), lr=lr, weight_decay=weight_decay)
        self.scheduler = torch.optim.lr_scheduler.CosineAnnealingLR(self.optimizer, T_max=num_epochs, eta_min=1e-5)

        self.teacher = teacher
        if teacher is not None:
            teacher.eval()
            for p in teacher.parameters():
                p.requires_grad_(False)
            _native.mark_frozen(teacher)          # nothing writes to it: its eval-mode caches never go stale
        self.kd_temperature, self.kd_alpha, self.kd_beta = kd_temperature, kd_alpha, kd_beta
        cls = getattr(getattr(model, "head", None), "cls", None)
        self._num_classes = int(getattr(cls, "out_channels", 0) or 0)       # 0: unknown head, the loss call counts itself
        self.amp_dtype = amp_dtype

        self.save_dir = save_dir
        if self.rank == 0:
            os.makedirs(save_dir, exist_ok=True)
        self.best_miou = 0.0
        self.history_path = os.path.join(save_dir, "training_history.json")
        self.history = {"train_loss": [], "train_miou": [], "val_loss": [], "val_miou": [], "lr": []}
        self.last_loss_terms: Optional[torch.Tensor] = None
        # CUDA-graph replay of the whole step (forward + loss + backward + all-reduce + AdamW): the step is
        # ~800 launches of mostly small kernels, so at B200 speeds the host cannot issue them fast enough
        self.overlap_teacher = overlap_teacher and teacher is not None
        self._side = None
        self.overlap_allreduce = overlap_allreduce and self.world_size > 1
        self._plan, self._arrival, self._order_hooks = None, None, []
        self._overlap_armed = self._overlap_fired = False
        self._comm_stream = None
        self.use_cuda_graph = use_cuda_graph
        self.graph_warmup_steps = graph_warmup_steps
        # one captured graph per input shape (the last, partial batch of an epoch gets its own instead of evicting the
        # full-batch one twice per epoch); beyond ``max_graphs`` shapes a step runs eagerly
        self._graphs: Dict[tuple, dict] = {}
        self.max_graphs = 2
        self._eager_steps = 0
        self.sync_replicas()

    # ------------------------------------------------------------------ data-parallel replicas
    @torch.no_grad()
    def sync_replicas(self):
        """Every rank continues from rank 0's parameters, optimizer moments and buffers (BatchNorm running statistics,
        batch counters).  Replicas that start from different random initialisations would apply the averaged gradient
        to different points and never meet; called at construction and after ``load_checkpoint``."""
        if self.world_size <= 1:
            return
        opt = self.optimizer
        for t in (opt.flat_param, opt.exp_avg, opt.exp_avg_sq):
            dist.broadcast(t, src=0)
        step = torch.tensor([opt._step], dtype=torch.int64, device=self.device)
        dist.broadcast(step, src=0)
        opt._step = int(step.item())
        for b in self.model.buffers():
            dist.broadcast(b, src=0)
        _native.bump_generation()

    def release_graphs(self):
        """Drop the captured step graphs (and their static buffers).  Call before
        ``torch.distributed.destroy_process_group()``: a communicator whose all-reduce still sits in a live
        captured graph cannot be torn down."""
        self._graphs.clear()
        torch.cuda.synchronize(self.device)

    # ------------------------------------------------------------------ one optimisation step
    def _autocast(self):
        return torch.autocast("cuda", dtype=self.amp_dtype, enabled=self.amp_dtype is not None)

    def _step_impl(self, imgs, pts, seg, update_hyper: bool):
        self.optimizer.detach_grads()                    # gradients arrive as fresh tensors, gathered below in one copy
        counted = None
        with self._autocast(), _native.deferred_batch_counters():
            if self.teacher is not None:
                # The frozen teacher's forward does not depend on the student's: it runs on a side stream (a parallel
                # branch of the captured graph), so that the two camera branches -- chains of small kernels that leave
                # most SMs idle at their heads and tails -- fill each other's gaps.  The cell ordering both LiDAR
                # encoders share is built first, on the main stream.
                side = self._side_stream()
                main = torch.cuda.current_stream(self.device)
                if side is not None:
                    self._prepare_shared(pts)
                    side.wait_stream(main)
                    with torch.cuda.stream(side), torch.no_grad():
                        # the label histogram the loss normaliser needs depends on the labels only: taken here, on the
                        # side stream, it is off the critical path between the student's logits and the loss kernel
                        if self._num_classes:
                            counted = ops.kd_label_count(seg, self._num_classes, ignore_index=-1)
                        t_logits, t_mid = self.teacher(imgs, pts, return_intermediates=True)
                else:
                    with torch.no_grad():
                        t_logits, t_mid = self.teacher(imgs, pts, return_intermediates=True)
                logits, mid = self.model(imgs, pts, return_intermediates=True)
                if side is not None:
                    main.wait_stream(side)
                    for t in [t_logits] + [t_mid[k] for k in MIMIC_TAPS] + ([counted] if counted is not None else []):
                        t.record_stream(main)
                if counted is not None and logits.shape[1] != self._num_classes:
                    counted = None                               # a head with another class count: count inside the call
                s_feats = [mid[k] for k in MIMIC_TAPS]
                t_feats = [t_mid[k] for k in MIMIC_TAPS]
                alpha, beta = self.kd_alpha, self.kd_beta
            else:
                logits = self.model(imgs, pts)
                t_logits, s_feats, t_feats, alpha, beta = None, [], [], 0.0, 0.0
        terms, d_logits, d_feats = ops.kd_loss_fwd_bwd(
            logits, t_logits, seg, self.class_weights, s_feats, t_feats,
            T=self.kd_temperature, alpha=alpha, beta=beta, ignore_index=-1, counted_ws=counted)
        plan = self._overlap_plan() if self.world_size > 1 else None
        if plan is not None:
            self._overlap_armed = True
        torch.autograd.backward([logits] + list(s_feats), [d_logits] + list(d_feats))
        if plan is not None and self._overlap_fired:
            # the early bucket is already being reduced on the communication stream (launched from the sentinel
            # parameter's hook while the rest of the backward ran): only the late, small bucket is left
            self._overlap_armed = self._overlap_fired = False
            self.optimizer.gather_grads_(plan["late"])
            lo, hi = plan["late_span"]
            allreduce_gradients_(self.optimizer.flat_grad[lo:hi])
            torch.cuda.current_stream(self.device).wait_stream(self._comm_stream)
        else:
            self._overlap_armed = False
            self.optimizer.gather_grads_()                              # one multi-tensor copy into the flat bucket
            allreduce_gradients_(self.optimizer.flat_grad)              # one flat NCCL bucket (no-op at world 1)
        self.optimizer.step(grad_scale=1.0 / self.world_size, update_hyper=update_hyper)
        return terms, logits.detach()

    # ------------------------------------------------------------------ gradient all-reduce overlapped with the backward
    # The flat bucket is latency-bound (2.1 MB), so what matters is WHEN it is launched, not how it is split: the first
    # data-parallel step records the order in which the parameters' gradients arrive; the parameters whose gradients
    # arrive last (the first layers of the camera encoder, at most ``overlap_tail_bytes``) form the late bucket, all
    # others the early one.  From then on a hook on the early bucket's last-arriving parameter gathers that bucket and
    # launches its all-reduce on a communication stream -- a parallel branch of the captured step graph -- while the
    # backward of the remaining layers runs; after the backward only the small late bucket is reduced in line.
    overlap_tail_bytes = 96 * 1024

    def _overlap_plan(self):
        if not self.overlap_allreduce:
            return None
        if self._plan is not None:
            return self._plan or None
        opt = self.optimizer
        if self._arrival is None:                                      # first step: record the arrival order
            self._arrival = []
            self._order_hooks = [p.register_post_accumulate_grad_hook(lambda _p, i=i: self._arrival.append(i))
                                 for i, p in enumerate(opt._params)]
            return None
        for h in self._order_hooks:
            h.remove()
        self._order_hooks = []
        order, late, acc = self._arrival, [], 0
        for i in reversed(order):
            nbytes = opt._params[i].numel() * 4
            if acc + nbytes > self.overlap_tail_bytes:
                break
            late.append(i)
            acc += nbytes
        never = [i for i in range(len(opt._params)) if i not in set(order)]      # parameters that get no gradient
        early = [i for i in order if i not in set(late)]
        late_all = late + never
        e_span, l_span = opt.span(early), opt.span(late_all) if late_all else None
        if not early or not late_all or e_span is None or l_span is None:
            self._plan = {}                                            # buckets not contiguous in the flat layout: one bucket
            return None
        sentinel = early[-1]                                           # the early bucket is complete when this gradient lands
        self._plan = {"early": early, "late": late_all, "early_span": e_span, "late_span": l_span, "sentinel": sentinel}
        self._comm_stream = torch.cuda.Stream(self.device)
        opt._params[sentinel].register_post_accumulate_grad_hook(self._launch_early_bucket)
        return self._plan

    def _launch_early_bucket(self, _param):
        if not self._overlap_armed or self._overlap_fired:
            return
        plan = self._plan
        with torch.no_grad():
            self.optimizer.gather_grads_(plan["early"])
            here = torch.cuda.current_stream(self.device)
            self._comm_stream.wait_stream(here)
            lo, hi = plan["early_span"]
            with torch.cuda.stream(self._comm_stream):
                allreduce_gradients_(self.optimizer.flat_grad[lo:hi])
        self._overlap_fired = True

    def _side_stream(self):
        if not self.overlap_teacher:
            return None
        if self._side is None:
            self._side = torch.cuda.Stream(self.device)
        return self._side

    def _prepare_shared(self, pts):
        """Cell ordering of the sweep, built once on the current stream before the teacher forks off (both LiDAR
        encoders then hit the cache)."""
        enc = getattr(getattr(self.model, "lidar_encoder", None), "encoder", None)
        if enc is not None and hasattr(enc, "_geom") and pts.is_cuda and pts.dtype == torch.float32:
            from .. import point_mlp
            point_mlp.cached_build_order(pts, enc._geom, tuple(enc.grid_size))

    def _capture(self, key, imgs, pts, seg):
        static = {"image": torch.empty_like(imgs), "points": torch.empty_like(pts), "seg": torch.empty_like(seg)}
        for k, v in (("image", imgs), ("points", pts), ("seg", seg)):
            static[k].copy_(v)
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        k0 = _native.launch_stats["kernels"]
        with torch.cuda.graph(graph):
            out = self._step_impl(static["image"], static["points"], static["seg"], update_hyper=False)
        kernels = _native.launch_stats["kernels"] - k0                  # our kernels inside one replay
        _native.launch_stats["kernels"] = k0
        self._graphs[key] = {"graph": graph, "static": static, "out": out, "kernels": kernels}
        self._graph_kernels = kernels
        return self._graphs[key]

    def training_step(self, imgs: torch.Tensor, pts: torch.Tensor, seg: torch.Tensor):
        """zero_grad -> forward(s) -> fused loss+grad -> backward -> (all-reduce) -> AdamW.
        Returns (loss terms f32[8] on device, student logits).  Never syncs the host.  With
        ``use_cuda_graph`` the first ``graph_warmup_steps`` calls run eagerly, then the whole step is
        captured once per input shape and replayed (outputs are then static buffers, valid until the
        next call)."""
        _native.bump_generation()            # parameters / running statistics move (also inside a replayed graph)
        if not self.use_cuda_graph:
            out = self._step_impl(imgs, pts, seg, update_hyper=True)
            self.last_loss_terms = out[0]
            return out
        key = (imgs.shape, pts.shape, seg.shape, imgs.dtype, self.model.training)
        entry = self._graphs.get(key)
        fresh = False
        if entry is None:
            if self._eager_steps < self.graph_warmup_steps or len(self._graphs) >= self.max_graphs:
                self._eager_steps += 1
                out = self._step_impl(imgs, pts, seg, update_hyper=True)
                self.last_loss_terms = out[0]
                return out
            entry, fresh = self._capture(key, imgs, pts, seg), True
        if not fresh:
            st = entry["static"]
            st["image"].copy_(imgs, non_blocking=True)
            st["points"].copy_(pts, non_blocking=True)
            st["seg"].copy_(seg, non_blocking=True)
        opt = self.optimizer
        opt._step += 1
        opt.set_hyper(opt.param_groups[0]["lr"], opt._step)      # device-side (lr, step) the captured AdamW reads
        entry["graph"].replay()
        _native.launch_stats["kernels"] += entry["kernels"]
        self.last_loss_terms = entry["out"][0]
        return entry["out"]

    # ------------------------------------------------------------------ epochs
    def _to_device(self, batch):
        return (batch["image"].to(self.device, non_blocking=True),
                batch["points"].to(self.device, non_blocking=True),
                batch["segmentation"].to(self.device, non_blocking=True))

    def train_epoch(self):
        self.model.train()
        metrics = SegmentationMetrics(num_classes=2)
        total = torch.zeros((), dtype=torch.float32, device=self.device)
        for batch in tqdm(self.train_loader, desc="Train", disable=not self.verbose):
            imgs, pts, seg = self._to_device(batch)
            terms, logits = self.training_step(imgs, pts, seg)
            total += terms[0]
            metrics.update(logits.detach(), seg)
        if self.world_size > 1:
            dist.all_reduce(total)
            total /= self.world_size
            metrics.all_reduce()
        return total.item() / max(len(self.train_loader), 1), metrics.compute()

    def validate(self):
        self.model.eval()
        metrics = SegmentationMetrics(num_classes=2)
        total = torch.zeros((), dtype=torch.float32, device=self.device)
        with torch.no_grad():
            for batch in tqdm(self.val_loader, desc="Val", disable=not self.verbose):
                imgs, pts, seg = self._to_device(batch)
                with self._autocast():
                    logits = self.model(imgs, pts)
                total += self.criterion(logits, seg)
                metrics.update(logits, seg)
        if self.world_size > 1:
            dist.all_reduce(total)
            total /= self.world_size
            metrics.all_reduce()
        return total.item() / max(len(self.val_loader), 1), metrics.compute()

    # ------------------------------------------------------------------ checkpoints / history (reference formats)
    def save_checkpoint(self, epoch, val_miou, is_best=False):
        if self.rank != 0:
            return
        ckpt = {"epoch": epoch, "model_state": self.model.state_dict(),
                "optimizer_state": self.optimizer.state_dict(),
                "scheduler_state": self.scheduler.state_dict(), "val_miou": val_miou}
        torch.save(ckpt, os.path.join(self.save_dir, "latest.pth"))
        if is_best:
            torch.save(ckpt, os.path.join(self.save_dir, "best.pth"))

    def load_checkpoint(self, path):
        ckpt = torch.load(path, map_location=self.device)
        self.model.load_state_dict(ckpt["model_state"])
        self.optimizer.load_state_dict(ckpt["optimizer_state"])
        if "scheduler_state" in ckpt:
            self.scheduler.load_state_dict(ckpt["scheduler_state"])
        self.best_miou = ckpt.get("val_miou", 0.0)
        start_epoch = ckpt.get("epoch", 0) + 1
        _native.bump_generation()
        self.sync_replicas()
        if self.verbose:
            print(f"Resumed from {path}, starting at epoch {start_epoch}, best mIoU {self.best_miou:.4f}")
        return start_epoch

    def update_history(self, train_loss, train_miou, val_loss, val_miou, lr):
        for k, v in zip(("train_loss", "train_miou", "val_loss", "val_miou", "lr"),
                        (train_loss, train_miou, val_loss, val_miou, lr)):
            self.history[k].append(v)
        if self.rank == 0:
            with open(self.history_path, "w") as f:
                json.dump(self.history, f, indent=2)

    def train(self, start_epoch=0):
        log = print if self.verbose else (lambda *a, **k: None)
        log(f"\nStarting training from epoch {start_epoch + 1}/{self.num_epochs}")
        log("=" * 60)
        for epoch in range(start_epoch, self.num_epochs):
            log(f"\nEpoch {epoch + 1}/{self.num_epochs}")
            log("-" * 60)
            for loader in (self.train_loader, self.val_loader):
                sampler = getattr(loader, "sampler", None)
                if hasattr(sampler, "set_epoch"):                # DistributedSampler: a new shuffle every epoch
                    sampler.set_epoch(epoch)
            train_loss, train_metrics = self.train_epoch()
            val_loss, val_metrics = self.validate()
            self.scheduler.step()
            lr = self.optimizer.param_groups[0]["lr"]
            train_miou, val_miou = train_metrics["miou"], val_metrics["miou"]
            log("\nResults:")
            log(f"  Train Loss: {train_loss:.4f} | Train mIoU: {train_miou:.4f}")
            log(f"  Val Loss:   {val_loss:.4f} | Val mIoU:   {val_miou:.4f}")
            log(f"  Learning Rate: {lr:.6f}")
            log("\n  Per-class IoU (Val):")
            for name, iou in zip(("Background", "Drivable"), val_metrics["class_iou"]):
                log(f"    {name:12s}: {iou:.4f}")
            self.update_history(train_loss, train_miou, val_loss, val_miou, lr)
            is_best = val_miou > self.best_miou
            if is_best:
                self.best_miou = val_miou
                log(f"  New best mIoU: {val_miou:.4f}")
            self.save_checkpoint(epoch, val_miou, is_best=is_best)
        log("\n" + "=" * 60)
        log(f"Training completed! Best validation mIoU: {self.best_miou:.4f}")
        log("=" * 60)
        return self.best_miou
