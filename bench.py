#!/usr/bin/env python
"""Headline benchmark: teacher->student KD training frames/s on synthetic PandaSet-shaped frames.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --gpus 1 --steps 3 --warmup 1      # CPU arm (oracle port of the reference)

Workload (BASELINE.json configs[1]; weak scaling for N>1 = configs[3]): frozen concat/256 teacher
(eval) -> weighted/128 student (train), 2 classes, class weights [0.4, 3.5], KD loss (CE + T^2 KL +
feature-mimic MSE), AdamW, bf16 activations with fp32 points / index math / statistics, 32 frames
per GPU, 170k-point Pandar64-shaped sweeps, 256x256 images, 64x64 BEV grid.

One JSON line on stdout (rank 0).  `value` = frames/s with inputs resident in HBM; `e2e` = the same
step driven from pinned HOST batches through Trainer.training_step with the H2D copies and a D2H read
of the loss terms inside the timed region; `roofline` = the dominant hand-written kernel timed alone
with CUDA events at the step's shapes; `cpu_baseline` = the oracle port of the reference's CPU path
on a bounded sample (rank 0, N=1 only).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "lightweight-multi-modal-scene-understanding-via-knowledge-distillation_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

METRIC = "kd_train_frames_per_sec"
UNIT = "frames/s"
CLASS_WEIGHTS = [0.4, 3.5]                    # train_with_fusion_ablation.py:47
FALLBACK_PEAK_GBS = 6650.0                    # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=("native", "reference"))
    ap.add_argument("--batch", type=int, default=32, help="frames per GPU (weak scaling)")
    ap.add_argument("--points", type=int, default=170_000)
    ap.add_argument("--fp32", action="store_true", help="fp32 activations instead of bf16")
    ap.add_argument("--student", default="weighted", choices=("weighted", "concat", "minimal"),
                    help="student fusion ablation (BASELINE.json configs[2] sweeps all three at --batch 64)")
    ap.add_argument("--cpu-batch", type=int, default=2, help="frames per CPU-baseline step (bounded sample)")
    ap.add_argument("--cpu-steps", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-kernels", action="store_true", help="skip the per-kernel roofline section")
    ap.add_argument("--no-graph", action="store_true", help="issue the step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--no-overlap-allreduce", action="store_true",
                    help="one in-line gradient bucket after the backward instead of the overlapped early + late buckets (N > 1)")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the other BASELINE configurations (ablation sweep at batch 64, fp32 step, projection microbench, "
                         "reference on CUDA, strong scaling at global batch 256) that the default run appends as `extras`")
    return ap.parse_args()


def workload_config(args, n_gpus):
    width = 256 if args.student == "concat" else 128
    return {"workload": f"teacher(concat/256, eval) -> student({args.student}/{width}, train) KD training step, 2-class, "
                        f"{'fp32' if args.fp32 else 'bf16'} activations, {args.batch} frames/GPU, "
                        f"{args.points}-pt sweeps, 256x256 images, 64x64 BEV",
            "global_batch": args.batch * n_gpus, "frames_per_gpu": args.batch, "frames_per_step": args.batch * n_gpus,
            "points_per_frame": args.points,
            "parallelism": f"dp{n_gpus}", "loss": "0.5*CE + 0.5*T^2*KL(T=4) + 1.0*MSE(lidar_feat, camera_feat)",
            "optimizer": "AdamW lr 1e-3 wd 1e-3 (flat, one kernel)",
            "allreduce": ("none (1 GPU)" if n_gpus == 1 else "one in-line flat bucket" if getattr(args, "no_overlap_allreduce", False) else
                          "flat bucket in two parts: early part launched from a gradient hook on a communication stream while the "
                          "backward of the first camera stages runs, late part (<= 96 KB) in line"),
            "launch": "eager" if args.no_graph else "whole step replayed as one CUDA graph",
            "l2": "per-step working set (>= 10 GB of activations, 113 MB of inputs) far exceeds the 126 MB L2; no flush"}


# ============================================================================= CPU arm (oracle port)
def cpu_step_runner(batch, points, seed=0, fusion="weighted"):
    """The reference's CPU path, restated by the oracle: teacher fwd (eval) + student fwd/bwd (train) +
    KD loss + torch AdamW, eager fp32 on all host threads."""
    from oracle import kd_oracle, model_oracle
    from oracle.weights import make_state_dict, synthetic_frames
    torch.set_num_threads(os.cpu_count() or 1)
    sd_s = model_oracle.clone_state(make_state_dict(5, fusion_type=fusion), requires_grad=True)
    sd_t = model_oracle.clone_state(make_state_dict(6, fusion_type="concat", random_running_stats=True))
    params = [v for v in sd_s.values() if v.requires_grad]
    opt = torch.optim.AdamW(params, lr=1e-3, weight_decay=1e-3)
    w = torch.tensor(CLASS_WEIGHTS)
    img, pts, lab = synthetic_frames(seed, batch, points)

    def step():
        opt.zero_grad()
        with torch.no_grad():
            tl, tm = model_oracle.model_forward(img, pts, sd_t, fusion_type="concat", train=False)
        sl, sm = model_oracle.model_forward(img, pts, sd_s, fusion_type=fusion, train=True)
        out = kd_oracle.kd_loss(sl, tl, lab, w, [sm[k] for k in kd_oracle.MIMIC_TAPS], [tm[k] for k in kd_oracle.MIMIC_TAPS])
        out["loss"].backward()
        opt.step()
        return float(out["loss"].detach())
    return step


def time_cpu(batch, points, steps, warmup, fusion="weighted"):
    step = cpu_step_runner(batch, points, fusion=fusion)
    for _ in range(warmup):
        step()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
    sec = statistics.median(ts)
    return {"value": batch / sec, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
            "sample": f"{steps} timed + {warmup} warm-up KD steps of {batch} frames x {points} pts on the host "
                      f"(oracle port of the reference's eager fp32 path, torch {torch.__version__}, "
                      f"{torch.get_num_threads()} threads), median {sec * 1e3:.0f} ms/step",
            "ms_per_step": sec * 1e3}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    cb = time_cpu(args.cpu_batch, args.points, max(1, args.steps), max(0, args.warmup), args.student)
    cfg = workload_config(args, args.gpus)
    cfg["sample"] = f"each step is a bounded sample of {args.cpu_batch} frames of the same workload"
    # what this arm really ran, machine-readable: ONE host process whatever --gpus says, a bounded sample per step,
    # fp32 (the reference has no bf16 path: SURVEY.md section 0.3)
    cfg.update({"frames_per_step": args.cpu_batch, "same_config": False, "activations": "fp32", "processes": 1,
                "kind": "port (oracle restatement of the reference's eager CPU path; the reference has no packaging to install)"})
    line = {"impl": "reference", "kind": "port", "same_config": False, "frames_per_step": args.cpu_batch,
            "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ============================================================================= clocks
class ClockSampler:
    """`nvidia-smi -lms` on this rank's GPU, started before the warm-up steps (the tool needs a few hundred ms to start,
    longer on an 8-GPU box) and bracketed by mark_begin()/mark_end() around the timed region: the clocks reported are
    the samples whose timestamps fall inside the bracket; when the timed region is shorter than the sampling
    period they are the samples under load around it (warm-up + timed), and `window` says which."""
    FIELDS = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")
    NAMES = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")

    def __init__(self, device):
        self.proc = None
        self.t_begin = self.t_end = None
        try:
            uuid = str(torch.cuda.get_device_properties(device).uuid)
            if not uuid.startswith("GPU-"):
                uuid = "GPU-" + uuid
            self.cmd = ["nvidia-smi", "-i", uuid, f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "50"]
        except Exception:
            self.cmd = None

    def __enter__(self):
        if self.cmd:
            try:
                self.proc = subprocess.Popen(self.cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            except OSError:
                self.proc = None
        return self

    def mark_begin(self):
        import datetime
        self.t_begin = datetime.datetime.now()

    def mark_end(self):
        import datetime
        self.t_end = datetime.datetime.now()

    def __exit__(self, *exc):
        import datetime
        self.result = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "window": None}
        if not self.proc:
            return
        time.sleep(0.1)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        rows = []
        for ln in out.splitlines():
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(parts[0], "%Y/%m/%d %H:%M:%S.%f")
                rows.append((ts, float(parts[1]), float(parts[2]), parts[4:8]))
            except ValueError:
                continue
        inside = [r for r in rows if self.t_begin and self.t_end and self.t_begin <= r[0] <= self.t_end]
        near = [r for r in rows if self.t_begin and r[0] >= self.t_begin - datetime.timedelta(seconds=0.3)]
        use, window = (inside, "timed region") if inside else (near, "warm-up + timed region (timed region shorter than the sampling period)")
        if use:
            reasons = set()
            for _, _, _, flags in use:
                for name, val in zip(self.NAMES, flags):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            self.result = {"sm_mhz": statistics.median(r[1] for r in use), "sm_max_mhz": max(r[2] for r in use),
                           "reasons": sorted(reasons), "samples": len(use), "window": window}


# ============================================================================= native arm
OVERLAP_ALLREDUCE = True


def build_models(device, fp32, use_graph=False, student_fusion="weighted"):
    from src.models.camera_encoder import TwinLiteEncoder
    from src.models.fusion_module import CompleteSegmentationModel
    from src.models.lidar_encoder import LiDAREncoder
    from src.training.trainer import Trainer

    def make(ft, oc):
        return CompleteSegmentationModel(TwinLiteEncoder(return_multiscale=True),
                                         LiDAREncoder("spatial", grid_size=(64, 64), use_vectorized=True), num_classes=2,
                                         fusion_type=ft, fusion_out_channels=oc,
                                         camera_fpn_stages=["stage3", "stage4", "stage5"], camera_fpn_channels=128,
                                         output_mode="same").to(device)
    torch.manual_seed(0)                                  # identical replicas on every rank
    student, teacher = make(student_fusion, 256 if student_fusion == "concat" else 128), make("concat", 256)
    trainer = Trainer(student, [], [], device, lr=1e-3, weight_decay=1e-3, class_weights=CLASS_WEIGHTS,
                      save_dir=os.path.join(ROOT, "gpurun_out", "bench_ckpt"), teacher=teacher,
                      amp_dtype=None if fp32 else torch.bfloat16, verbose=False, use_cuda_graph=use_graph,
                      overlap_allreduce=OVERLAP_ALLREDUCE)
    student.train()
    return trainer


def peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_PEAK_GBS, "fallback (B200_PROFILING.md)"


def time_kernel(fn, iters=20, warm=3):
    """Average device time of fn() in ms: CUDA events on the launching (current) stream."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def ncu_traffic():
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of each hand-written kernel at the bench
    shapes, from the committed `ncu --set full` captures (profiles/r01_ncu_traffic.json names the capture files)."""
    out = {}
    for name in ("r01_ncu_traffic.json", "r02_ncu_traffic.json", "r02b_ncu_traffic.json"):          # the later capture wins for kernels it covers
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                out.update(json.load(f)["bytes_per_launch"])
        except Exception:
            pass
    return out


def kernel_rooflines(args, device, fp32):
    """Each hand-written kernel of the step alone at the step's shapes.  Algorithmic bytes follow SURVEY.md 8(d) and
    DESIGN.md section 4 (stated per kernel in `note`); inputs per launch are >> L2 for the per-point kernels, and a
    256 MB buffer is rewritten between launches for the small (per-pixel) ones."""
    from src import native, ops, point_mlp
    from src.data_loading.synthetic_frames import make_frames
    B, N, C, H, W = args.batch, args.points, 128, 64, 64
    s = 4 if fp32 else 2
    dt = torch.float32 if fp32 else torch.bfloat16
    pts = make_frames(B, N, seed=123, device=device)["points"]
    geom = ops.bev_range_constants([-50, -50, -5, 50, 50, 3])
    peak, peak_src = peak_gbs()
    traffic = ncu_traffic()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    out = []
    f32 = dict(device=device, dtype=torch.float32)
    st = native.stream_ptr(device)
    p = native.ptr
    Mpts = B * N

    def add(name, ms, alg_bytes, note, in_step=True):
        out.append({"kernel": name, "ms": ms, "algorithmic_bytes": alg_bytes, "achieved": alg_bytes / (ms * 1e-3) / 1e9,
                    "peak": peak, "unit": "GB/s", "frac": alg_bytes / (ms * 1e-3) / 1e9 / peak, "bound": "hbm",
                    "traffic": traffic.get(name), "in_step": in_step, "note": note})

    def with_flush(fn):
        def g():
            flush.fill_(1)
            fn()
        return g
    t_flush = time_kernel(lambda: flush.fill_(1))

    cell, count, order, offsets = point_mlp.bev_build_order(pts, geom, (H, W))
    v = (cell >= 0).float().mean().item()

    if not fp32:
        # ---- fused point MLP (tcgen05): the four layer kernels of the student (the teacher runs the two forward ones)
        q, r = torch.randn(64, 4, **f32) * 0.02, torch.randn(64, **f32) * 0.1
        W2 = (torch.randn(128, 64, **f32) / 8).to(dt)
        W3 = (torch.randn(128, 128, **f32) / 11).to(dt)
        sc, sh = torch.rand(128, **f32) + 0.5, torch.randn(128, **f32) * 0.1
        flat = pts.view(Mpts, 4)
        z2, _ = point_mlp.mlp_layer_fwd_raw(0, flat, q, r, W2)
        add("mlp_layer_fwd_kernel<0>", time_kernel(lambda: point_mlp.mlp_layer_fwd_raw(0, flat, q, r, W2), 10), Mpts * (16 + C * s),
            "layers 1+2: 16 B point in, C*s pre-BatchNorm row out per point")
        add("mlp_layer_fwd_kernel<1>", time_kernel(lambda: point_mlp.mlp_layer_fwd_raw(1, z2, sc, sh, W3), 10), Mpts * 2 * C * s,
            "layer 3: C*s row in, C*s row out per point")
        add("mlp_eval3_kernel", time_kernel(lambda: point_mlp.mlp_eval3_fwd(flat, q, r, W2, sc, sh, W3), 10), Mpts * (16 + C * s),
            "the teacher's whole point MLP (running statistics) in one kernel: 16 B point in, C*s pre-BatchNorm-3 row out per point")
        z3, _ = point_mlp.mlp_layer_fwd_raw(1, z2, sc, sh, W3)
        dy = (torch.randn(Mpts, 128, device=device) * (torch.rand(Mpts, 128, device=device) < 0.05)).to(dt)
        gs, ga, gb = torch.rand(128, **f32) + 0.5, torch.randn(128, **f32) * 0.01, torch.randn(128, **f32) * 0.01
        add("mlp_layer_bwd_kernel<1>", time_kernel(lambda: ops.mlp_layer_bwd(1, dy, z3, gs, ga, gb, z2, sc, sh, W3), 10), Mpts * 4 * C * s,
            "layer 3 backward: dy, z, z_prev rows in, dy_prev row out per point (dgrad + wgrad + BatchNorm backward)")
        add("mlp_layer_bwd_kernel<0>", time_kernel(lambda: ops.mlp_layer_bwd(0, dy, z3, gs, ga, gb, flat, q, r, W2), 10), Mpts * (2 * C * s + 16),
            "layers 2+1 backward: dy, z rows + 16 B point in per point; 64x5 sums out")
        del dy
        # ---- projection of the un-normalised last layer
        add("bev_build_order (index+scan+fill)", time_kernel(lambda: point_mlp.bev_build_order(pts, geom, (H, W)), 10),
            B * (16 * N + 4 * N + 4 * N + 4 * N + 4 * v * N + 8 * H * W),
            "16N points in; cell ids, arrival ranks out and in again, order out; occupancy + offsets per cell")
        add("bev_reduce_affine_kernel", time_kernel(lambda: point_mlp.bev_reduce_affine(z3, sc, sh, order, offsets, B, N, (H, W), True), 10),
            B * (C * s * v * N + 4 * v * N + 2 * C * s * H * W),
            "C*s*v*N rows of valid points read once + 4*v*N ids; grid and extreme written once (2*C*s*HW)")
        grid, grid_z = point_mlp.bev_reduce_affine(z3, sc, sh, order, offsets, B, N, (H, W), True)
        gg = torch.randn(B, H, W, C, device=device, dtype=dt)
        add("bev_bwd_affine_kernel",
            time_kernel(lambda: point_mlp.bev_bwd_affine(gg, z3, grid, grid_z, order, offsets, cell, B, N, (H, W), zero_outside=False), 10),
            B * (3 * C * s * H * W + 2 * C * s * v * N + 8 * v * N),
            "grad/grid/extreme rows per cell (3*C*s*HW) + rows of valid points read once (C*s*v*N; the second sweep hits "
            "L1/L2) + a gradient row per valid point (C*s*v*N; rows of points outside are masked by the consumer) + ids")
        del z2, z3, gg, grid, grid_z

    # ---- layer-by-layer projection path (fp32 parity path; bf16 features here): the reference's scatter as one call
    feats = torch.rand(B, N, C, device=device, dtype=dt)
    grid = torch.empty(B, H, W, C, dtype=dt, device=device)
    cnt = torch.empty(B, H * W, dtype=torch.int32, device=device)
    cel = torch.empty(B, N, dtype=torch.int32, device=device)
    order2 = torch.empty(B, N, dtype=torch.int32, device=device)
    offs = torch.empty(B, H * W + 1, dtype=torch.int32, device=device)
    wsb = native.lib.kdf_bev_workspace_bytes(B, N, H, W)
    ws = torch.empty(wsb, dtype=torch.uint8, device=device)

    def proj_fwd():
        native.call("kdf_bev_project_fwd", p(pts), 4, p(feats), native.dtype_code(feats), B, N, C, *geom, H, W, 0,
                    p(grid), p(cnt), p(cel), None, p(order2), p(offs), p(ws), wsb, st)
    add("bev_project_fwd (index+scan+fill+max)", time_kernel(proj_fwd), B * (16 * N + C * s * v * N + C * s * H * W + 4 * H * W),
        "16N + C*s*v*N + C*s*HW + 4*HW per frame (SURVEY 8d)", in_step=fp32)
    gg = torch.rand(B, H * W, C, device=device, dtype=dt)
    gf = torch.empty(B, N, C, dtype=dt, device=device)

    def proj_bwd():
        native.call("kdf_bev_project_bwd", p(gg), p(feats), p(grid), None, None, p(cel), p(order2), p(offs),
                    native.dtype_code(feats), B, N, C, H, W, 0, p(gf), st)
    add("bev_bwd_max_kernel", time_kernel(proj_bwd), B * (2 * C * s * H * W + 2 * C * s * v * N + C * s * (1 - v) * N + 4 * N + 4 * v * N),
        "grad/max rows per cell (2*C*s*HW) + rows of valid points read once for the exact tie split (C*s*v*N; the second "
        "sweep hits L1/L2) + a gradient row per point (C*s*N, zeros for points outside) + cell ids and order", in_step=fp32)

    def index_only():
        native.call("kdf_bev_index", p(pts), B, N, 4, *geom, H, W, p(cel), None, p(cnt), st)
    add("bev_index_kernel", time_kernel(index_only), B * (16 * N + 4 * N + 4 * H * W), "16N + 4N + 4*HW per frame")
    del feats, gf, gg

    # ---- fusion (weighted) forward / backward on pre-BN rows
    M = B * H * W
    cam_pre = torch.randn(M, C, device=device, dtype=dt)
    lid_pre = torch.randn(M, C, device=device, dtype=dt)
    sc = [torch.rand(C, **f32) + 0.5 for _ in range(2)]
    sh = [torch.randn(C, **f32) * 0.1 for _ in range(2)]
    w1, b1 = torch.randn(C, 2 * C, **f32) * 0.05, torch.randn(C, **f32) * 0.1
    w2, b2 = torch.randn(2, C, **f32) * 0.1, torch.randn(2, **f32) * 0.1
    fo = torch.empty(M, C, dtype=dt, device=device)
    attn = torch.empty(M, 2, **f32)

    def fus_fwd():
        native.call("kdf_fusion_weighted_fwd", p(cam_pre), p(lid_pre), native.dtype_code(cam_pre), M, C, p(sc[0]), p(sh[0]),
                    p(sc[1]), p(sh[1]), p(w1), p(b1), p(w2), p(b2), p(fo), p(attn), st)
    add("fusion_weighted_fwd_kernel", time_kernel(with_flush(fus_fwd)) - t_flush, M * 3 * C * s,
        "3*C*s per pixel (whole block fused, single pass); L2 flushed between launches")
    go = torch.randn(M, C, device=device, dtype=dt)
    g1, g2 = torch.empty_like(cam_pre), torch.empty_like(lid_pre)
    gaff, gw1, gb1 = torch.empty(4, C, **f32), torch.empty(C, 2 * C, **f32), torch.empty(C, **f32)
    gw2, gb2 = torch.empty(2, C, **f32), torch.empty(2, **f32)

    def fus_bwd():
        native.call("kdf_fusion_weighted_bwd", p(go), p(cam_pre), p(lid_pre), native.dtype_code(cam_pre), M, C, p(sc[0]),
                    p(sh[0]), p(sc[1]), p(sh[1]), p(w1), p(b1), p(w2), p(b2), p(attn), p(g1), p(g2), p(gaff), p(gw1),
                    p(gb1), p(gw2), p(gb2), st)
    add("fusion_weighted_bwd_kernel", time_kernel(with_flush(fus_bwd)) - t_flush, M * 5 * C * s,
        "5*C*s per pixel (grad_out + 2 inputs read, 2 input grads written)")

    # ---- KD loss forward+backward
    zs = torch.randn(B, 2, H, W, device=device, dtype=dt)
    zt = torch.randn(B, 2, H, W, device=device, dtype=dt)
    lab = (torch.rand(B, H, W, device=device) < 0.13).long()
    taps_s = [torch.randn(B, H, W, C, device=device, dtype=dt).permute(0, 3, 1, 2) for _ in range(2)]
    taps_t = [torch.randn(B, H, W, C, device=device, dtype=dt).permute(0, 3, 1, 2) for _ in range(2)]
    cw = torch.tensor(CLASS_WEIGHTS, device=device)
    d_l = [torch.empty_like(t) for t in taps_s]
    dz, scal = torch.empty_like(zs), torch.empty(8, **f32)
    kws = torch.empty(native.lib.kdf_kd_loss_workspace_bytes(), dtype=torch.uint8, device=device)
    K = 2

    def kd():
        native.call("kdf_kd_loss_fwd_bwd", p(zs), p(zt), p(lab), p(cw), B, K, H * W, native.dtype_code(zs), 4.0, 0.5, 1.0, -1,
                    p(taps_s[0]), p(taps_t[0]), p(d_l[0]), taps_s[0].numel(), p(taps_s[1]), p(taps_t[1]), p(d_l[1]),
                    taps_s[1].numel(), native.dtype_code(taps_s[0]), 1.0, p(dz), p(scal), p(kws), st)
    add("kd_loss_kernel (label histogram + loss fwd+bwd)", time_kernel(with_flush(kd)) - t_flush, M * (2 * K * s + 8 + K * s + 2 * 3 * C * s),
        "per pixel 2K*s + 8 + K*s logits/labels + 3*C*s per mimic tap (2 taps); the stand-alone call (histogram inside)", in_step=False)

    def kd_count():
        native.call("kdf_kd_label_count", p(lab), B, K, H * W, -1, p(kws), st)

    def kd_counted():
        native.call("kdf_kd_loss_fwd_bwd_counted", p(zs), p(zt), p(lab), p(cw), B, K, H * W, native.dtype_code(zs), 4.0, 0.5, 1.0, -1,
                    p(taps_s[0]), p(taps_t[0]), p(d_l[0]), taps_s[0].numel(), p(taps_s[1]), p(taps_t[1]), p(d_l[1]),
                    taps_s[1].numel(), native.dtype_code(taps_s[0]), 1.0, p(dz), p(scal), p(kws), st)
    kd_count()

    t_cnt = time_kernel(with_flush(kd_count)) - t_flush

    def kd_pair():
        kd_count()
        kd_counted()
    t_pair = time_kernel(with_flush(kd_pair)) - t_flush
    add("kd_loss_kernel (loss fwd+bwd; label histogram taken on the side stream when the batch arrives)", t_pair - t_cnt,
        M * (2 * K * s + K * s + 2 * 3 * C * s),
        "per pixel 2K*s + K*s logits + 3*C*s per mimic tap (2 taps); the step's form: kdf_kd_label_count runs next to the "
        "teacher's forward, kdf_kd_loss_fwd_bwd_counted is what sits between the logits and the backward "
        "(timed as [count + loss] - [count])")
    add("kd_label_count (memset + histogram, side stream)", t_cnt, M * 8, "8 B label per pixel", in_step=True)
    # ---- fused 1x1-convolution layers of the camera branch (tcgen05 + TMA): three representative layers and the sum over all
    #      seventeen layers of the student's forward (tools/pw_conv_bench.py lists each)
    if not fp32:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import pw_conv_bench as pwb
        scale_m = B / 32.0
        tot_ms, tot_bytes = 0.0, 0
        named = {"stage2 expand": "pw_conv_fwd stage2 expand 32->192 @128x128 (rows + statistics)",
                 "stage3 project": "pw_conv_fwd stage3 project 384->64 @64x64 (BN+ReLU6 prologue, rows + statistics)",
                 "fpn post / fusion proj": "pw_conv_fwd fpn post 128->128 @64x64 (BN+ReLU prologue, rows + statistics)"}
        for lname, Mp, Kp, Np, pro in pwb.LAYERS[:14]:
            Mp = int(Mp * scale_m)
            x = torch.randn(Mp, Kp, device=device).to(torch.bfloat16)
            wq = torch.randn(Np, Kp, device=device) * (2.0 / Kp) ** 0.5
            pack = ops._pw_pack_factor(Kp, Np)
            wb = ops.pw_conv_weight(wq, pack)
            psc, psh = torch.rand(Kp, **f32) + 0.5, torch.randn(Kp, **f32) * 0.1
            ms_l = time_kernel(with_flush(lambda: ops.pw_conv_fwd(x, wb, pack, pro=(psc, psh, 2) if pro else None, want_stats=True))) - t_flush
            reps = {"fpn lateral4": 2, "fpn post / fusion proj": 3}.get(lname, 1)      # shapes that occur more than once in the network
            tot_ms += ms_l * reps
            tot_bytes += Mp * (Kp + Np) * 2 * reps
            if lname in named:
                add(named[lname], ms_l, Mp * (Kp + Np) * 2, "rows in + rows out, once (K*s + N*s per pixel row); L2 flushed between launches")
            del x
        add("pw_conv_fwd, the 17 1x1 convolutions of the student's forward (sum)", tot_ms, tot_bytes,
            "sum over the layers of rows in + rows out; batch statistics come out of the same kernels")
        # ---- depthwise 3x3 (TMA-staged tiles, packed fp32 FMAs): the eight layers of the student, forward (+ statistics), data
        #      gradient and weight gradient; two representative layers by name, all of them summed
        dw_layers = (("stage1", 32, 128, 1), ("stage2", 192, 128, 2), ("stage3", 384, 64, 1), ("stage4", 384, 64, 2),
                     ("stage5", 768, 32, 1), ("fpn smooth", 128, 64, 1), ("head dw0", 128, 64, 1), ("head dw1", 64, 64, 1))
        tot = {"fwd": [0.0, 0], "dgrad": [0.0, 0], "wgrad": [0.0, 0]}
        for lname, Cc, Hh, st_ in dw_layers:
            OHh = (Hh - 1) // st_ + 1
            xin = torch.randn(B, Hh, Hh, Cc, device=device).to(dt)
            yo = torch.empty(B, OHh, OHh, Cc, device=device, dtype=dt)
            gr = torch.randn(B, OHh, OHh, Cc, device=device).to(dt)
            gx = torch.empty_like(xin)
            w9 = torch.randn(Cc, 9, **f32) * 0.3
            gw = torch.empty(Cc, 9, **f32)
            sts = torch.empty(2, Cc, dtype=torch.float64, device=device)
            nb = (xin.numel() + yo.numel()) * s
            t_f = time_kernel(with_flush(lambda: native.call("kdf_dwconv3x3_fwd", p(xin), p(w9), 1, B, Hh, Hh, Cc, st_, 0, p(yo), p(sts), st))) - t_flush
            t_d = time_kernel(with_flush(lambda: native.call("kdf_dwconv3x3_bwd_data", p(gr), p(w9), 1, B, Hh, Hh, Cc, st_, p(gx), st))) - t_flush
            t_w = time_kernel(with_flush(lambda: native.call("kdf_dwconv3x3_bwd_weight", p(xin), p(gr), 1, B, Hh, Hh, Cc, st_, p(gw), st))) - t_flush
            for key, t in (("fwd", t_f), ("dgrad", t_d), ("wgrad", t_w)):
                tot[key][0] += t
                tot[key][1] += nb
            if lname in ("stage2", "stage3"):
                shape = f"{Cc} ch @{Hh}x{Hh} stride {st_}"
                add(f"dwconv3x3 forward + statistics, {lname} ({shape})", t_f, nb, "map in + map out, once; L2 flushed between launches")
                add(f"dwconv3x3 data gradient, {lname} ({shape})", t_d, nb, "gradient map in + gradient map out, once")
                add(f"dwconv3x3 weight gradient, {lname} ({shape})", t_w, nb, "input map + gradient map in, once")
            del xin, yo, gr, gx
        add("dwconv3x3 forward + statistics, the 8 depthwise layers of the student (sum)", tot["fwd"][0], tot["fwd"][1], "sum over the layers")
        add("dwconv3x3 data gradients, the 8 depthwise layers (sum)", tot["dgrad"][0], tot["dgrad"][1], "sum over the layers")
        add("dwconv3x3 weight gradients, the 8 depthwise layers (sum)", tot["wgrad"][0], tot["wgrad"][1], "sum over the layers")
        # ---- camera stem (fp32 NCHW image -> bf16 rows) and the head's classifier
        img = torch.rand(B, 3, 256, 256, **f32)
        wst = torch.randn(32, 3, 3, 3, **f32) * 0.3
        so = torch.empty(B, 128, 128, 32, device=device, dtype=dt)
        sst = torch.empty(2, 32, dtype=torch.float64, device=device)
        sgw = torch.empty(32, 27, **f32)
        sgr = torch.randn(B, 128, 128, 32, device=device).to(dt)
        nb = img.numel() * 4 + so.numel() * s
        add("stem_conv_fwd_kernel (3x3 s2, 3->32, image fp32 NCHW -> bf16 rows + statistics)",
            time_kernel(with_flush(lambda: native.call("kdf_stem_conv_fwd", p(img), p(wst), B, 256, 256, None, None, 0, p(so), p(sst), st))) - t_flush,
            nb, "image in (4 B/value) + rows out; 864 multiply-adds per output pixel on the CUDA cores: shared-memory bound, not HBM bound")
        add("stem_conv_wgrad_kernel", time_kernel(with_flush(lambda: native.call("kdf_stem_conv_bwd_weight", p(img), p(sgr), B, 256, 256, p(sgw), st))) - t_flush,
            nb, "image + gradient rows in")
        del img, so, sgr
        xr = torch.randn(M, 32, device=device).to(dt)
        wc, bc = torch.randn(2, 32, **f32) * 0.3, torch.randn(2, **f32)
        lo = torch.empty(B, 2, H, W, device=device, dtype=dt)
        dl = torch.randn(B, 2, H, W, device=device).to(dt)
        dxr, gwc, gbc = torch.empty_like(xr), torch.empty(2, 32, **f32), torch.empty(2, **f32)
        add("cls_conv_fwd_kernel (32 -> 2 classes, planar logits)",
            time_kernel(with_flush(lambda: native.call("kdf_cls_conv_fwd", p(xr), p(wc), p(bc), M, 32, 2, H * W, p(lo), st))) - t_flush,
            M * (32 + 2) * s, "64 B row in, 2 logits out per pixel (8.9 MB: launch-latency bound)")
        add("cls_conv_bwd_kernel", time_kernel(with_flush(lambda: native.call("kdf_cls_conv_bwd", p(xr), p(dl), p(wc), M, 32, 2, H * W, p(dxr), p(gwc), p(gbc), st))) - t_flush,
            M * (64 + 2) * s, "row + 2 logit gradients in, gradient row out per pixel")
        # ---- the cell-sorted form of the LiDAR projection: built and tested, not in the step (slower in total)
        cell_s, count_s, offsets_s, spts, cs, _ = point_mlp.bev_build_sorted(pts, geom, (H, W))
        add("bev_build_sorted (index+scan+permute: the points in cell order)", time_kernel(lambda: point_mlp.bev_build_sorted(pts, geom, (H, W)), 10),
            B * (16 * N + 4 * N + 4 * N + 4 * N + 16 * N + 4 * N + 8 * H * W), "as bev_build_order + 16N sorted points and 4N cell ids out", in_step=False)
        z3 = torch.randn(Mpts, 128, device=device).to(dt)
        z2 = torch.randn(Mpts, 128, device=device).to(dt)
        sc3, sh3 = torch.rand(128, **f32) + 0.5, torch.randn(128, **f32) * 0.1
        add("bev_reduce_affine_kernel over cell-sorted rows (contiguous segments)",
            time_kernel(lambda: point_mlp.bev_reduce_affine(z3, sc3, sh3, None, offsets_s, B, N, (H, W), True), 10),
            B * (C * s * v * N + 2 * C * s * H * W), "rows of valid points once; grid and extreme out", in_step=False)
        grid_s, gz_s = point_mlp.bev_reduce_affine(z3, sc3, sh3, None, offsets_s, B, N, (H, W), True)
        gg = torch.randn(B, H, W, C, device=device, dtype=dt)
        add("bev_bwd_share_kernel (per-cell shares + per-row tie bits, no gradient rows)",
            time_kernel(lambda: point_mlp.bev_bwd_share(gg, z3, grid_s, gz_s, offsets_s, B, N, (H, W)), 10),
            B * (C * s * v * N + (C // 8) * v * N + 4 * C * s * H * W), "rows of valid points once, C/8 bytes of tie bits per valid row out, 4 rows per cell", in_step=False)
        share, bits, _s3 = point_mlp.bev_bwd_share(gg, z3, grid_s, gz_s, offsets_s, B, N, (H, W))
        gs, ga, gb = torch.rand(128, **f32) + 0.5, torch.randn(128, **f32) * 0.01, torch.randn(128, **f32) * 0.01
        add("mlp_layer_bwd_kernel<1, share> (layer 3 backward forming dy from shares + tie bits)",
            time_kernel(lambda: point_mlp.mlp_layer_bwd_share(cs.view(-1), share, bits, z3, gs, ga, gb, z2, sc3, sh3, W3), 10),
            Mpts * (3 * C * s + C // 8 + 4), "z, z_prev rows in, dy_prev row out, 16 B of tie bits + 4 B cell id per point", in_step=False)
        del z3, z2, gg, share, bits
    return out, v


def run_native(args):
    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl native needs a CUDA device (the product has no CPU fallback)")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # NCCL's own banner must not land on stdout (one JSON line)
        dist.init_process_group("nccl", device_id=device)
    n_gpus = world
    from src import native
    from src.data_loading.synthetic_frames import make_frames
    from src.training.parallel import frame_seed, reduce_max

    global OVERLAP_ALLREDUCE
    OVERLAP_ALLREDUCE = not args.no_overlap_allreduce
    trainer = build_models(device, args.fp32, use_graph=not args.no_graph, student_fusion=args.student)
    B, N = args.batch, args.points
    n_data = 3
    batches = [make_frames(B, N, seed=frame_seed(rank, i), device=device) for i in range(n_data)]

    def step(i, b=None):
        b = b or batches[i % n_data]
        return trainer.training_step(b["image"], b["points"], b["segmentation"])

    def fence():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- value: inputs resident in HBM
    with ClockSampler(device) as clk:     # started early: nvidia-smi needs a few hundred ms before its first sample
        if not args.no_graph:             # untimed: eager steps + one-off capture of the step graph
            for i in range(trainer.graph_warmup_steps + 1):
                step(i)
        for i in range(args.warmup):
            step(i)
        fence()
        k0 = native.launch_stats["kernels"]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        clk.mark_begin()
        e0.record()
        for i in range(args.steps):
            terms, _ = step(args.warmup + i)
        e1.record()
        fence()
        clk.mark_end()
        launches = native.launch_stats["kernels"] - k0
    ms = reduce_max(e0.elapsed_time(e1) / args.steps, device)
    value = B * n_gpus / (ms * 1e-3)
    loss_terms = [float(x) for x in terms[:4].float().cpu()]

    # ---------------- e2e: pinned host batches, H2D inside the timed region, D2H of the loss terms
    e2e = None
    if not args.no_e2e:
        host = [{k: v.cpu().pin_memory() for k, v in b.items() if torch.is_tensor(v)} for b in batches[:2]]
        h2d = sum(v.numel() * v.element_size() for v in host[0].values())
        dev_bufs = [{k: torch.empty_like(v, device=device) for k, v in host[0].items()} for _ in range(2)]
        host_terms = [torch.empty(8, dtype=torch.float32).pin_memory() for _ in range(2)]
        copy_stream = torch.cuda.Stream(device)
        ready = [torch.cuda.Event() for _ in range(2)]
        done = [torch.cuda.Event() for _ in range(2)]
        total = args.warmup + args.steps

        def prefetch(i):
            slot = i % 2
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(done[slot])              # the step that last used this slot has finished
                for k, v in host[i % 2].items():
                    dev_bufs[slot][k].copy_(v, non_blocking=True)
                ready[slot].record(copy_stream)

        fence()
        for ev in done:
            ev.record()
        prefetch(0)
        t0 = t1 = None
        for i in range(total):
            if i == args.warmup:
                fence()
                t0 = torch.cuda.Event(enable_timing=True); t0.record()
            slot = i % 2
            torch.cuda.current_stream().wait_event(ready[slot])
            if i + 1 < total:
                prefetch(i + 1)                                  # overlaps with this step's compute
            tt, _ = step(i, dev_bufs[slot])
            done[slot].record()
            host_terms[slot].copy_(tt, non_blocking=True)        # D2H read of the step's result
        t1 = torch.cuda.Event(enable_timing=True); t1.record()
        fence()
        ms_e2e = reduce_max(t0.elapsed_time(t1) / args.steps, device)
        e2e = {"value": B * n_gpus / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 32,
               "ms_per_step": ms_e2e, "api": "src.training.trainer.Trainer.training_step on pinned host batches "
               "(double-buffered H2D on a copy stream), loss terms copied back every step"}

    # ---------------- per-kernel rooflines (rank 0 only does the work; others wait at the barrier)
    kernels, roof = [], None
    if not args.no_kernels and rank == 0:
        del batches
        torch.cuda.empty_cache()
        kernels, valid_frac = kernel_rooflines(args, device, args.fp32)
        dom = max((k for k in kernels if k["in_step"]), key=lambda k: k["ms"])     # the longest hand-written kernel of the step
        roof = {"bound": "hbm", "achieved": dom["achieved"], "peak": dom["peak"], "unit": "GB/s", "frac": dom["frac"],
                "traffic": dom["traffic"], "kernel": dom["kernel"], "launch_ms": dom["ms"],
                "algorithmic_bytes_per_launch": dom["algorithmic_bytes"], "peak_source": peak_gbs()[1],
                "valid_point_fraction": valid_frac}
    if world > 1:
        dist.barrier()

    # ---------------- the other BASELINE configurations (tools/bench_extras.py), each with its own clock record
    extras = None
    if not args.no_extras and not args.fp32:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_extras as bx
        me = sys.modules[__name__]
        extras = {}
        trainer.release_graphs()
        torch.cuda.empty_cache()
        if n_gpus == 1:
            extras["ablation_b64"] = [bx.trainer_leg(me, device, rank, world, 64, args.points, st, False)
                                      for st in ("weighted", "concat", "minimal")]
            extras["fp32_step"] = bx.trainer_leg(me, device, rank, world, args.batch, args.points, args.student, True, steps=3, warmup=2)
            extras["projection"] = bx.projection(device)
            extras["reference_on_cuda"] = bx.reference_on_cuda(me, device, 4, args.points)
            if not args.no_cpu_baseline:
                extras["cpu_ce_only"] = bx.cpu_ce_only(me, args.cpu_batch, args.points)
        elif 256 % n_gpus == 0:
            leg = bx.trainer_leg(me, device, rank, world, 256 // n_gpus, args.points, args.student, False)
            leg["scaling"] = "strong"
            extras["strong256"] = leg
    if world > 1:
        dist.barrier()

    # ---------------- CPU baseline (rank 0, N=1 only)
    cpu = None
    if rank == 0 and n_gpus == 1 and not args.no_cpu_baseline:
        cb = time_cpu(args.cpu_batch, args.points, args.cpu_steps, 1, args.student)
        cpu = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32" if args.fp32 else "bf16", "data": "synthetic",
                "config": workload_config(args, n_gpus), "clocks": clk.result, "e2e": e2e,
                "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu, "kernels": kernels, "extras": extras,
                "loss_terms_last_step": dict(zip(("loss", "ce", "kl", "mse"), loss_terms))}
        emit(line)
    if world > 1:
        # the captured step graph holds the communicator's all-reduce: drop the graphs first, then tear down
        dist.barrier()
        torch.cuda.synchronize()
        trainer.release_graphs()
        dist.destroy_process_group()


_JSON_FD = None


def emit(line):
    """The ONE JSON line, on the process's real stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    global _JSON_FD
    args = parse_args()
    # Libraries print on stdout too (NCCL's version banner when the box sets NCCL_DEBUG): file descriptor 1 is pointed at
    # stderr for the duration of the run and the JSON line goes to a duplicate of the original stdout.
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
