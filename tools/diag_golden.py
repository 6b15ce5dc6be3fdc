"""Developer tool: per-tensor error of the CUDA model against the golden fixtures."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "lightweight-multi-modal-scene-understanding-via-knowledge-distillation_b200"))
import numpy as np, torch
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
from test_gpu_model import build, rel_err, _sample
from oracle.weights import make_state_dict, synthetic_frames
from oracle import model_oracle, kd_oracle
for ft in ("weighted",):
    z = np.load(os.path.join(ROOT, "tests/golden", f"model_{ft}_train.npz"))
    wseed, fseed, B, N, ih, iw = (int(v) for v in z["meta"])
    sd = make_state_dict(wseed, fusion_type=ft)
    model = build(ft); model.load_state_dict(sd); model.cuda().train()
    img, pts, lab = synthetic_frames(fseed, B, N, image_hw=(ih, iw), edge_cases=True, nonfinite=False)
    logits, mid = model(img.cuda(), pts.cuda(), return_intermediates=True)
    print("logits", rel_err(logits.detach().cpu(), z["logits"]))
    for k in ("camera_feat", "lidar_feat", "pre_fusion", "post_fusion"):
        print(k, rel_err(_sample(mid[k]), z["sample_" + k]))
    from src.training.trainer import _Criterion
    loss = _Criterion(torch.tensor([0.4, 3.5]).cuda())(logits, lab.cuda()); loss.backward()
    print("loss", loss.item(), float(z["loss"]))
    # oracle grads for all params (CPU autograd) for a complete picture
    so = model_oracle.clone_state(sd, requires_grad=True)
    sl, sm = model_oracle.model_forward(img, pts, so, fusion_type=ft, train=True)
    kd_oracle.ce_loss(sl, lab, torch.tensor([0.4, 3.5])).backward()
    rows = []
    for name, p in model.named_parameters():
        r = so[name].grad
        if r.abs().max() < 1e-6: continue
        rows.append((rel_err(p.grad.cpu(), r), name))
    rows.sort(reverse=True)
    for e, n in rows[:25]: print(f"{e:.3e} {n}")
