"""The BASELINE.json configurations beyond the headline one, measured inside ``bench.py``'s default run so that the
driver's BENCH / SCALE records carry them (``extras`` on the one JSON line):

  ablation_b64        configs[2]: KD step with the weighted / concat / minimal student at 64 frames per GPU
  fp32_step           the fp32 companion of the headline step (the 1e-5 parity path)
  projection          configs[4]: LiDAR projection microbench, 100k / 500k / 2M points x 64 frames, with the
                      reference's eager PyTorch scatter timed on the same GPU ("reference on CUDA")
  reference_on_cuda   the oracle port of the reference's eager fp32 training step (KD step and CE-only step) run on the
                      GPU -- what stock PyTorch makes of this workload on a B200
  cpu_ce_only         the reference's own (CE-only, no teacher) step on the host cores
  strong256           configs[3]: global batch 256 over the N ranks of the run (256/N frames per GPU), N > 1

Everything is bounded (a few steps each); clocks are sampled per leg."""
from __future__ import annotations

import os
import statistics
import time

import torch


def _time_steps(step, steps, warmup, fence):
    for i in range(warmup):
        step(i)
    fence()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        step(warmup + i)
    e1.record()
    fence()
    return e0.elapsed_time(e1) / steps


def trainer_leg(bench, device, rank, world, batch, points, student, fp32, steps=5, warmup=3, use_graph=True):
    """One KD-step configuration through Trainer.training_step: ms/step (max over ranks) and whole-job frames/s."""
    import torch.distributed as dist
    from src.data_loading.synthetic_frames import make_frames
    from src.training.parallel import frame_seed, reduce_max
    trainer = bench.build_models(device, fp32, use_graph=use_graph, student_fusion=student)
    batches = [make_frames(batch, points, seed=frame_seed(rank, 50 + i), device=device) for i in range(2)]

    def step(i):
        b = batches[i % 2]
        return trainer.training_step(b["image"], b["points"], b["segmentation"])

    def fence():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    with bench.ClockSampler(device) as clk:
        if use_graph:
            for i in range(trainer.graph_warmup_steps + 1):
                step(i)
        clk.mark_begin()
        ms = _time_steps(step, steps, warmup, fence)
        clk.mark_end()
    ms = reduce_max(ms, device)
    trainer.release_graphs()
    del trainer, batches
    torch.cuda.empty_cache()
    return {"student": student, "dtype": "f32" if fp32 else "bf16", "frames_per_gpu": batch, "global_batch": batch * world,
            "points_per_frame": points, "ms_per_step": ms, "value": batch * world / (ms * 1e-3), "unit": "frames/s",
            "steps": steps, "warmup": warmup, "launch": "CUDA graph replay" if use_graph else "eager", "clocks": clk.result}


def reference_on_cuda(bench, device, batch, points, steps=3, warmup=1):
    """The oracle port of the reference's eager fp32 path with its tensors on the GPU: KD step and CE-only step."""
    from oracle import kd_oracle, model_oracle
    from oracle.weights import make_state_dict, synthetic_frames
    out = {}
    img, pts, lab = (t.to(device) for t in synthetic_frames(0, batch, points))
    w = torch.tensor(bench.CLASS_WEIGHTS, device=device)
    for kind in ("kd", "ce_only"):
        sd_s = {k: v.to(device) for k, v in model_oracle.clone_state(make_state_dict(5, fusion_type="weighted")).items()}
        for v in sd_s.values():
            if v.is_floating_point() and v.dim() > 0:
                v.requires_grad_(True)
        for k in sd_s:
            if "running_" in k or k.endswith("grid_tensor"):
                sd_s[k].requires_grad_(False)
        sd_t = {k: v.to(device) for k, v in model_oracle.clone_state(make_state_dict(6, fusion_type="concat", random_running_stats=True)).items()}
        opt = torch.optim.AdamW([v for v in sd_s.values() if v.requires_grad], lr=1e-3, weight_decay=1e-3)

        def step(_i):
            opt.zero_grad()
            sl, sm = model_oracle.model_forward(img, pts, sd_s, fusion_type="weighted", train=True)
            if kind == "kd":
                with torch.no_grad():
                    tl, tm = model_oracle.model_forward(img, pts, sd_t, fusion_type="concat", train=False)
                loss = kd_oracle.kd_loss(sl, tl, lab, w, [sm[k] for k in kd_oracle.MIMIC_TAPS], [tm[k] for k in kd_oracle.MIMIC_TAPS])["loss"]
            else:
                loss = torch.nn.functional.cross_entropy(sl, lab, weight=w, ignore_index=-1)
            loss.backward()
            opt.step()
        ms = _time_steps(step, steps, warmup, torch.cuda.synchronize)
        out[kind] = {"ms_per_step": ms, "value": batch / (ms * 1e-3), "unit": "frames/s", "frames_per_step": batch,
                     "dtype": "f32", "what": "oracle port of the reference's eager PyTorch step, tensors on the GPU (stock ATen / cuDNN kernels)"}
        del sd_s, sd_t, opt
        torch.cuda.empty_cache()
    return out


def cpu_ce_only(bench, batch, points, steps=3, warmup=1):
    """The reference's own training step (weighted CE, no teacher) on the host: trainer.py:81-93 via the oracle port."""
    from oracle import model_oracle
    from oracle.weights import make_state_dict, synthetic_frames
    torch.set_num_threads(os.cpu_count() or 1)
    sd = model_oracle.clone_state(make_state_dict(5, fusion_type="weighted"), requires_grad=True)
    opt = torch.optim.AdamW([v for v in sd.values() if v.requires_grad], lr=1e-3, weight_decay=1e-3)
    w = torch.tensor(bench.CLASS_WEIGHTS)
    img, pts, lab = synthetic_frames(0, batch, points)
    ts = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad()
        sl, _ = model_oracle.model_forward(img, pts, sd, fusion_type="weighted", train=True)
        torch.nn.functional.cross_entropy(sl, lab, weight=w, ignore_index=-1).backward()
        opt.step()
        if i >= warmup:
            ts.append(time.perf_counter() - t0)
    sec = statistics.median(ts)
    return {"ms_per_step": sec * 1e3, "value": batch / sec, "unit": "frames/s", "frames_per_step": batch, "cores": os.cpu_count() or 1,
            "kind": "port", "what": "the reference's CE-only step (no teacher, BASELINE.md 3.1 / configs[0] at full sweeps) on the host cores"}


def projection(device, points=(100_000, 500_000, 2_000_000), frames=64):
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("_pmb", os.path.join(root, "tools", "projection_microbench.py"))
    pmb = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(pmb)
    lines = pmb.measure(list(points), frames, ["bf16", "fp32"], ref_frames=2, iters=3, dev=device)
    keep = []
    for ln in lines:
        keep.append({"points_per_frame": ln["points_per_frame"], "frames": ln["frames"], "dtype": ln["dtype"],
                     "fwd_ms": ln["forward"]["ms"], "fwd_frac_of_peak": ln["forward"]["frac_of_peak"],
                     "bwd_ms": ln["backward"]["ms"], "bwd_frac_of_peak": ln["backward"]["frac_of_peak"],
                     "index_only_ms": ln["index_only"]["ms"], "index_frac_of_peak": ln["index_only"]["frac_of_peak"],
                     "fwd_points_per_s": ln["forward"]["points_per_s"],
                     "torch_scatter_fwd_ms_per_frame": ln["torch_scatter"]["fwd_ms_per_frame"],
                     "torch_scatter_fwd_bwd_ms_per_frame": ln["torch_scatter"]["fwd_bwd_ms_per_frame"],
                     "grid_bit_identical": ln["torch_scatter"]["grid_bit_identical"],
                     "speedup_fwd": ln["speedup_fwd_vs_torch_scatter"], "speedup_fwd_bwd": ln["speedup_fwd_bwd_vs_torch_scatter"]})
    return keep
