"""Three representative layers of tools/pw_conv_bench.py, launched a few times each (for ncu captures)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lightweight-multi-modal-scene-understanding-via-knowledge-distillation_b200"))
from src import ops  # noqa: E402

for name, M, K, N, pro in [("stage2 expand", 524288, 32, 192, False), ("stage3 project", 131072, 384, 64, True),
                           ("fpn post", 131072, 128, 128, True)]:
    x = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    w = torch.randn(N, K, device="cuda") * (2.0 / K) ** 0.5
    pack = ops._pw_pack_factor(K, N)
    wb = ops.pw_conv_weight(w, pack)
    sc, sh = torch.rand(K, device="cuda") + 0.5, torch.randn(K, device="cuda") * 0.1
    for _ in range(3):
        ops.pw_conv_fwd(x, wb, pack, pro=(sc, sh, 2) if pro else None, want_stats=True)
torch.cuda.synchronize()
