"""Shared pieces of the two entry points: argument parsing, data loaders (real
PandaSet when the root exists, seeded synthetic frames otherwise), model
construction and optional torch.distributed set-up."""
from __future__ import annotations

import argparse
import os

import torch
import torch.distributed as dist

from src.data_loading.pandaset_dataset import create_pandaset_dataloaders
from src.data_loading.synthetic_frames import create_synthetic_dataloaders
from src.models.camera_encoder import TwinLiteEncoder
from src.models.fusion_module import CompleteSegmentationModel
from src.models.lidar_encoder import LiDAREncoder

DEFAULT_ROOT = r"D:\kelvin\Dataset\data"          # the reference's hard-coded root (train_pandaset.py:81)


def base_parser(description: str) -> argparse.ArgumentParser:
    ap = argparse.ArgumentParser(description=description)
    ap.add_argument("--root", default=DEFAULT_ROOT, help="PandaSet root; synthetic frames are used when it does not exist")
    ap.add_argument("--synthetic", action="store_true", help="force seeded synthetic PandaSet-shaped frames")
    ap.add_argument("--synthetic-samples", type=int, nargs=2, default=(64, 16), metavar=("TRAIN", "VAL"))
    ap.add_argument("--points", type=int, default=5000, help="points per synthetic frame (reference subsamples to 5000)")
    ap.add_argument("--batch-size", type=int, default=4)
    ap.add_argument("--workers", type=int, default=2)
    ap.add_argument("--epochs", type=int, default=None)
    ap.add_argument("--bf16", action="store_true", help="bf16 activations (points and index math stay fp32)")
    ap.add_argument("--kd", action="store_true", help="distil from a frozen concat/256 teacher")
    ap.add_argument("--teacher-ckpt", default=None, help="checkpoint (reference format) to initialise the teacher")
    ap.add_argument("--kd-temperature", type=float, default=4.0)
    ap.add_argument("--kd-alpha", type=float, default=0.5)
    ap.add_argument("--kd-beta", type=float, default=1.0)
    ap.add_argument("--resume", choices=("ask", "yes", "no"), default="ask")
    return ap


def init_distributed() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("a CUDA device is required (the hot path has no CPU fallback)")
    if "RANK" in os.environ and "WORLD_SIZE" in os.environ and int(os.environ["WORLD_SIZE"]) > 1:
        local = int(os.environ.get("LOCAL_RANK", 0))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        return torch.device("cuda", local)
    return torch.device("cuda")


def make_loaders(args, verbose: bool):
    distributed = dist.is_available() and dist.is_initialized()
    if not args.synthetic and os.path.isdir(args.root):
        scenes = sorted(d for d in os.listdir(args.root) if d.isdigit())
        n_train = int(0.8 * len(scenes))                      # 80/20 scene split (train_pandaset.py:84-86)
        if verbose:
            print(f"Found {len(scenes)} scenes\nTrain: {n_train} scenes | Val: {len(scenes) - n_train} scenes")
        return create_pandaset_dataloaders(args.root, scenes[:n_train], scenes[n_train:], batch_size=args.batch_size,
                                           num_workers=args.workers, verbose=verbose, distributed=distributed)
    if verbose:
        print(f"Dataset root {args.root!r} not found -> seeded synthetic PandaSet-shaped frames")
    n_tr, n_va = args.synthetic_samples
    return create_synthetic_dataloaders(n_tr, n_va, batch_size=args.batch_size, num_workers=0,
                                        num_points=args.points, distributed=distributed)


def build_model(fusion_type: str, fusion_out_channels: int, num_classes: int, device) -> CompleteSegmentationModel:
    """The wiring both reference scripts use (train_with_fusion_ablation.py:27-39)."""
    return CompleteSegmentationModel(
        camera_encoder=TwinLiteEncoder(return_multiscale=True),
        lidar_encoder=LiDAREncoder(encoder_type="spatial", grid_size=(64, 64), use_vectorized=True),
        num_classes=num_classes, fusion_type=fusion_type, fusion_out_channels=fusion_out_channels,
        camera_fpn_stages=["stage3", "stage4", "stage5"], camera_fpn_channels=128,
        output_mode="same").to(device)


def build_teacher(args, num_classes: int, device):
    if not args.kd:
        return None
    teacher = build_model("concat", 256, num_classes, device)       # the reference's best variant
    if args.teacher_ckpt:
        state = torch.load(args.teacher_ckpt, map_location=device)
        teacher.load_state_dict(state.get("model_state", state))
    elif int(os.environ.get("RANK", 0)) == 0:
        print("\n" + "!" * 80 + "\nWARNING: --kd without --teacher-ckpt distils from a RANDOMLY INITIALISED concat teacher "
              "(eval mode, running\nstatistics at their initial values): the KL and feature-mimic terms then pull the student "
              "towards noise.\nUse this only for throughput runs; pass --teacher-ckpt <reference-format .pth> for real "
              "training.\n" + "!" * 80 + "\n")
    return teacher.eval()


def decide_resume(args, ckpt_path: str) -> bool:
    """ONE decision for all ranks: rank 0 looks at the checkpoint (and asks when ``--resume ask``), every rank gets
    its answer -- otherwise only rank 0 would load and the ranks would disagree on the start epoch."""
    rank0 = int(os.environ.get("RANK", 0)) == 0
    resume = False
    if rank0 and os.path.exists(ckpt_path) and args.resume != "no":
        resume = args.resume == "yes" or input(f"\nFound checkpoint at {ckpt_path}. Resume training? (y/n): ").lower() == "y"
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        box = [resume]
        dist.broadcast_object_list(box, src=0)
        resume = bool(box[0])
    return resume
