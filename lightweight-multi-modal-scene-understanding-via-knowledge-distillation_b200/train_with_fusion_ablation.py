"""Fusion ablation driver with the reference's protocol
(train_with_fusion_ablation.py:69-135): concat/256, minimal/128 and weighted/128
students, 2 classes, class weights [0.4, 3.5], 20 epochs each, results written to
fusion_ablation_results.json -- optionally distilled from a frozen teacher."""
import json
import os

import torch

from src.training.trainer import Trainer
from train_common import base_parser, build_model, build_teacher, init_distributed, make_loaders

VARIANTS = (("concat", 256), ("minimal", 128), ("weighted", 128))


def train_fusion_variant(fusion_type, fusion_out_channels, args, device, rank0=True):
    """Train one variant; returns (best mIoU, total params, fusion params) like the reference (:10-66)."""
    if rank0:
        print(f"\n{'=' * 80}\nTRAINING: {fusion_type.upper()} FUSION\n{'=' * 80}")
    train_loader, val_loader = make_loaders(args, verbose=False)
    model = build_model(fusion_type, fusion_out_channels, 2, device)
    summary = model.get_architecture_summary()
    if rank0:
        print(f"\nModel: {fusion_type}\n  Total params: {summary['total_params']}\n  Fusion params: {summary['fusion_params']}")
    trainer = Trainer(model=model, train_loader=train_loader, val_loader=val_loader, device=device,
                      lr=1e-3, weight_decay=1e-3, save_dir=f"checkpoints/fusion_ablation_{fusion_type}",
                      class_weights=[0.4, 3.5], num_epochs=args.epochs or 20,
                      teacher=build_teacher(args, 2, device), kd_temperature=args.kd_temperature,
                      kd_alpha=args.kd_alpha, kd_beta=args.kd_beta,
                      amp_dtype=torch.bfloat16 if args.bf16 else None)
    return trainer.train(), summary["total_params"], summary["fusion_params"]


def main(argv=None):
    ap = base_parser(__doc__)
    ap.add_argument("--variants", nargs="+", default=[v for v, _ in VARIANTS], choices=[v for v, _ in VARIANTS])
    ap.add_argument("--results", default="fusion_ablation_results.json")
    args = ap.parse_args(argv)
    device = init_distributed()
    rank0 = int(os.environ.get("RANK", 0)) == 0
    if rank0:
        print(f"\n{'=' * 80}\nFUSION ABLATION STUDY - 2-CLASS DRIVABLE AREA SEGMENTATION\n{'=' * 80}\nDevice: {device}")
    results = {}
    for ft, ch in VARIANTS:
        if ft not in args.variants:
            continue
        miou, total, fusion = train_fusion_variant(ft, ch, args, device, rank0)
        results[ft] = {"miou": miou, "total_params": total, "fusion_params": fusion}
    if rank0:
        print(f"\n{'=' * 80}\nFUSION ABLATION RESULTS\n{'=' * 80}")
        print(f"{'Fusion':<12} {'mIoU':>8} {'Total Params':>15} {'Fusion Params':>15}\n" + "-" * 80)
        for ft, d in results.items():
            print(f"{ft:<12} {d['miou']:>8.4f} {d['total_params']:>15} {d['fusion_params']:>15}")
        best = max(results.items(), key=lambda kv: kv[1]["miou"])
        print(f"\nBEST FUSION: {best[0].upper()}  (mIoU {best[1]['miou']:.4f}, {best[1]['total_params']} params)")
        with open(args.results, "w") as f:
            json.dump(results, f, indent=2)
        print(f"Results saved to {args.results}")
    return results


if __name__ == "__main__":
    main()
