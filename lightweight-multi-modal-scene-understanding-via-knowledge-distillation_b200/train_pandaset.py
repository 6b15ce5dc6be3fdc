"""Entry point with the reference's defaults (train_pandaset.py:71-166): concat
fusion, 3-class head, class weights [0.39, 2.61, 33.09], 30 epochs, batch 4,
checkpoints under checkpoints/pandaset_weighted -- now taking the fusion type,
class count and the distillation options as arguments.

    python train_pandaset.py --synthetic --fusion weighted --num-classes 2 --kd --bf16
    torchrun --nproc-per-node 8 train_pandaset.py --synthetic --kd --bf16
"""
import os

import torch

from src.training.trainer import Trainer
from train_common import base_parser, build_model, build_teacher, decide_resume, init_distributed, make_loaders

CLASS_WEIGHTS = {3: [0.39, 2.61, 33.09], 2: [0.4, 3.5]}       # train_pandaset.py:136, train_with_fusion_ablation.py:47


def main(argv=None):
    ap = base_parser(__doc__)
    ap.add_argument("--fusion", default="concat", choices=("concat", "minimal", "weighted"))
    ap.add_argument("--fusion-out-channels", type=int, default=None)
    ap.add_argument("--num-classes", type=int, default=3)
    ap.add_argument("--save-dir", default="checkpoints/pandaset_weighted")
    args = ap.parse_args(argv)

    device = init_distributed()
    rank0 = int(os.environ.get("RANK", 0)) == 0
    train_loader, val_loader = make_loaders(args, verbose=rank0)
    if rank0:
        print(f"\nUsing device: {device}\n\nBuilding model...")
    out_ch = args.fusion_out_channels or (256 if args.fusion == "concat" else 128)
    model = build_model(args.fusion, out_ch, args.num_classes, device)
    if rank0:
        s = model.get_architecture_summary()
        print("\nModel Architecture:")
        for label, key in (("Camera params: ", "camera_params"), ("LiDAR params:  ", "lidar_params"),
                           ("Fusion params: ", "fusion_params"), ("Head params:   ", "head_params"),
                           ("Total params:  ", "total_params")):
            print(f"  {label} {s[key]}")

    trainer = Trainer(model=model, train_loader=train_loader, val_loader=val_loader, device=device,
                      lr=1e-3, weight_decay=1e-3, save_dir=args.save_dir,
                      class_weights=CLASS_WEIGHTS.get(args.num_classes), num_epochs=args.epochs or 30,
                      teacher=build_teacher(args, args.num_classes, device), kd_temperature=args.kd_temperature,
                      kd_alpha=args.kd_alpha, kd_beta=args.kd_beta,
                      amp_dtype=torch.bfloat16 if args.bf16 else None)

    start_epoch = 0
    ckpt = os.path.join(args.save_dir, "latest.pth")
    if decide_resume(args, ckpt):                      # train_pandaset.py:155-160, decided once for all ranks
        start_epoch = trainer.load_checkpoint(ckpt)
    return trainer.train(start_epoch=start_epoch)


if __name__ == "__main__":
    main()
