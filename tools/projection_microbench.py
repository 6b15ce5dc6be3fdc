#!/usr/bin/env python
"""LiDAR projection microbench (BASELINE.json configs[4]): 100k-2M points/frame x 64 frames, sharded by frame over
the ranks (no collective), against the reference's PyTorch scatter run on the same GPU.

    python tools/projection_microbench.py                       # 1 GPU, 64 frames
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        tools/projection_microbench.py                          # 8 frames per GPU

Per (points/frame, feature dtype) one JSON line: our projection through the C ABI -- cell ids + occupancy only
(`kdf_bev_index`), the whole forward (`kdf_bev_project_fwd`: cell ids, occupancy, counting sort, per-cell maximum) and
the backward (`kdf_bev_project_bwd`) -- timed with CUDA events (max over ranks), as frames/s, points/s and algorithmic
GB/s against the measured HBM peak; and `torch_scatter`: the statements of
SpatialLiDAREncoder.forward_vectorized after the point MLP (reference lidar_encoder.py:42-55, 69-99: normalise, mask,
truncate, clamp, flat index, boolean-mask gather, scatter_reduce_ amax) in eager PyTorch on the same device, forward
and autograd backward, on a few frames (it materialises int64 / gathered temporaries), scaled per frame.
The grids of the two paths are compared bit for bit on the frames both processed.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "lightweight-multi-modal-scene-understanding-via-knowledge-distillation_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def torch_scatter_projection(points, feats, H, W):
    """What the reference does after its point MLP (lidar_encoder.py:42-55, 69-99), feats given point-major [B,N,C]."""
    B, N, C = feats.shape
    x_range = torch.tensor([-50, 50], device=points.device)            # int64 buffers, as registered at :38-39
    y_range = torch.tensor([-50, 50], device=points.device)
    x, y = points[..., 0], points[..., 1]
    xn = (x - x_range[0]) / (x_range[1] - x_range[0])
    yn = (y - y_range[0]) / (y_range[1] - y_range[0])
    valid = (xn >= 0) & (xn <= 1) & (yn >= 0) & (yn <= 1)
    coords = torch.stack([xn, yn], dim=-1)
    gc = (coords * torch.tensor([W - 1, H - 1], device=points.device, dtype=torch.float32)).long()
    gc[..., 0] = gc[..., 0].clamp(0, W - 1)
    gc[..., 1] = gc[..., 1].clamp(0, H - 1)
    flat = gc[..., 1] * W + gc[..., 0]
    batch_idx = torch.arange(B, device=points.device).view(B, 1).expand(B, N)
    gflat = batch_idx * (H * W) + flat
    vf = feats[valid]                                                   # [Nv, C] boolean-mask gather
    vi = gflat[valid]
    out = torch.zeros(B * H * W, C, device=points.device, dtype=feats.dtype)
    out.scatter_reduce_(0, vi.unsqueeze(1).expand(-1, C), vf, reduce="amax", include_self=False)
    return out.view(B, H, W, C)


def timed(fn, iters, warm=2):
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


def measure(points_list, frames, dtypes, channels=128, ref_frames=4, iters=5, rank=0, world=1, dev=None, emit=None):
    """The measurement loop: one dict per (points/frame, dtype).  ``frames`` = total frames over all ranks."""
    from src import native, ops
    from src.data_loading.synthetic_frames import make_frames

    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        peak = 6650.0
    H = W = 64
    C = channels
    geom = ops.bev_range_constants([-50, -50, -5, 50, 50, 3])
    F = frames // world                                              # frames of this rank
    p, st = native.ptr, native.stream_ptr(dev)
    lines = []
    for N in points_list:
        pts = make_frames(F, N, image_size=(8, 8), seed=1000 * rank + 7, device=dev)["points"]
        for dname in dtypes:
            dt, s = (torch.float32, 4) if dname == "fp32" else (torch.bfloat16, 2)
            feats = torch.empty(F, N, C, dtype=dt, device=dev)
            for f0 in range(0, F, 8):                                 # filled in slices: no fp32 temporary of the whole tensor
                feats[f0:f0 + 8] = torch.rand(feats[f0:f0 + 8].shape, device=dev, dtype=torch.float32).to(dt)
            grid = torch.empty(F, H, W, C, dtype=dt, device=dev)
            cnt = torch.empty(F, H * W, dtype=torch.int32, device=dev)
            cel = torch.empty(F, N, dtype=torch.int32, device=dev)
            order = torch.empty(F, N, dtype=torch.int32, device=dev)
            offs = torch.empty(F, H * W + 1, dtype=torch.int32, device=dev)
            wsb = native.lib.kdf_bev_workspace_bytes(F, N, H, W)
            ws = torch.empty(wsb, dtype=torch.uint8, device=dev)

            def index_only():
                native.call("kdf_bev_index", p(pts), F, N, 4, *geom, H, W, p(cel), None, p(cnt), st)

            def fwd():
                native.call("kdf_bev_project_fwd", p(pts), 4, p(feats), native.dtype_code(feats), F, N, C, *geom, H, W, 0,
                            p(grid), p(cnt), p(cel), None, p(order), p(offs), p(ws), wsb, st)
            t_idx = timed(index_only, iters)
            t_fwd = timed(fwd, iters)
            v = (cel >= 0).float().mean().item()
            # backward on as many frames as fit next to feats (grad rows are as large as feats)
            free, _ = torch.cuda.mem_get_info(dev)
            Fb = max(1, min(F, int((free - (8 << 30)) // (N * C * s))))
            gg = torch.randn(Fb, H * W, C, device=dev, dtype=torch.float32).to(dt)
            gf = torch.empty(Fb, N, C, dtype=dt, device=dev)

            def bwd():
                native.call("kdf_bev_project_bwd", p(gg), p(feats), p(grid), None, None, p(cel), p(order), p(offs),
                            native.dtype_code(feats), Fb, N, C, H, W, 0, p(gf), st)
            t_bwd = timed(bwd, iters) * F / Fb
            del gf, gg

            # the reference's eager scatter on a few frames, checked bit for bit against ours
            R = max(1, min(ref_frames, F))
            fr = feats[:R].clone().requires_grad_(True)
            ref = torch_scatter_projection(pts[:R], fr, H, W)
            same = bool(torch.equal(ref.detach(), grid[:R]))
            t_ref_f = timed(lambda: torch_scatter_projection(pts[:R], fr.detach(), H, W), 3, 1) / R
            gref = torch.randn_like(ref)

            def ref_fb():
                fr.grad = None
                torch_scatter_projection(pts[:R], fr, H, W).backward(gref)
            t_ref_fb = timed(ref_fb, 3, 1) / R
            del fr, ref, gref

            ts = torch.tensor([t_idx, t_fwd, t_bwd], device=dev, dtype=torch.float64)
            vv = torch.tensor([v], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(ts, op=dist.ReduceOp.MAX)
                dist.all_reduce(vv, op=dist.ReduceOp.SUM)
            t_idx, t_fwd, t_bwd = ts.tolist()
            v = vv.item() / world
            Ft = F * world
            b_idx = 16 * N + 4 * N + 4 * H * W
            b_fwd = 16 * N + C * s * v * N + C * s * H * W + 4 * H * W
            b_bwd = 2 * C * s * H * W + 2 * C * s * v * N + C * s * (1 - v) * N + 4 * N + 4 * v * N

            def leg(ms, bytes_per_frame):
                gbs = bytes_per_frame * F / (ms * 1e-3) / 1e9          # per GPU
                return {"ms": ms, "frames_per_s": Ft / (ms * 1e-3), "points_per_s": Ft * N / (ms * 1e-3),
                        "gbs_per_gpu": gbs, "frac_of_peak": gbs / peak}
            line = {"bench": "bev_projection", "points_per_frame": N, "frames": Ft, "n_gpus": world, "dtype": dname, "C": C,
                    "grid": [H, W], "valid_fraction": v, "peak_gbs": peak,
                    "index_only": leg(t_idx, b_idx), "forward": leg(t_fwd, b_fwd), "backward": leg(t_bwd, b_bwd),
                    "torch_scatter": {"frames_timed": R, "fwd_ms_per_frame": t_ref_f, "fwd_bwd_ms_per_frame": t_ref_fb,
                                      "fwd_frames_per_s_per_gpu": 1e3 / t_ref_f, "grid_bit_identical": same},
                    "speedup_fwd_vs_torch_scatter": t_ref_f / (t_fwd / F),
                    "speedup_fwd_bwd_vs_torch_scatter": t_ref_fb / ((t_fwd + t_bwd) / F)}
            if rank == 0:
                lines.append(line)
                if emit:
                    emit(line)
            del feats, grid, cnt, cel, order, offs, ws
            torch.cuda.empty_cache()
        del pts
    return lines


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=64, help="total frames (sharded over the ranks)")
    ap.add_argument("--points", type=int, nargs="*", default=[100_000, 250_000, 500_000, 1_000_000, 2_000_000])
    ap.add_argument("--dtypes", nargs="*", default=["fp32", "bf16"])
    ap.add_argument("--channels", type=int, default=128)
    ap.add_argument("--ref-frames", type=int, default=4, help="frames the eager PyTorch scatter is timed on")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--out", default=None, help="also append the JSON lines to this file (rank 0)")
    args = ap.parse_args()

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # NCCL's own banner must not land on stdout (one JSON line)
        dist.init_process_group("nccl", device_id=dev)
    lines = measure(args.points, args.frames, args.dtypes, args.channels, args.ref_frames, args.iters, rank, world, dev,
                    emit=lambda ln: print(json.dumps(ln), flush=True))
    if rank == 0 and args.out:
        with open(args.out, "a") as f:
            for ln in lines:
                f.write(json.dumps(ln) + "\n")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
