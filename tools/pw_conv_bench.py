"""Every 1x1 convolution of the student's forward at the bench shapes (B=32): kdf_pw_conv_fwd (rows + statistics, with
the BatchNorm+ReLU6 prologue where the network has one) against the library GEMM (F.linear, bf16) alone.
One JSON line per layer: ms, algorithmic GB/s (rows in + rows out), fraction of the measured HBM peak."""
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lightweight-multi-modal-scene-understanding-via-knowledge-distillation_b200"))
from src import ops  # noqa: E402

LAYERS = [("stage1 project", 524288, 32, 32, True), ("stage2 expand", 524288, 32, 192, False),
          ("stage2 project", 131072, 192, 64, True), ("stage3 expand", 131072, 64, 384, False),
          ("stage3 project", 131072, 384, 64, True), ("stage4 expand", 131072, 64, 384, False),
          ("stage4 project", 32768, 384, 128, True), ("stage5 expand", 32768, 128, 768, False),
          ("stage5 project", 32768, 768, 128, True), ("fpn lateral3", 131072, 64, 128, False),
          ("fpn lateral4", 32768, 128, 128, False), ("fpn post / fusion proj", 131072, 128, 128, True),
          ("head pw0", 131072, 128, 64, True), ("head pw1", 131072, 64, 32, True),
          ("teacher fuse pw", 131072, 256, 256, False), ("teacher head pw0", 131072, 256, 64, False)]


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    t_flush = timeit(lambda: flush.fill_(1))
    tot = {"ours": 0.0, "lib": 0.0, "bytes": 0}
    for name, M, K, N, pro in LAYERS:
        x = torch.randn(M, K, device="cuda").to(torch.bfloat16)
        w = torch.randn(N, K, device="cuda") * (2.0 / K) ** 0.5
        pack = ops._pw_pack_factor(K, N)
        wb, w16 = ops.pw_conv_weight(w, pack), w.to(torch.bfloat16)
        sc, sh = torch.rand(K, device="cuda") + 0.5, torch.randn(K, device="cuda") * 0.1
        pr = (sc, sh, 2) if pro else None

        def ours():
            flush.fill_(1)
            ops.pw_conv_fwd(x, wb, pack, pro=pr, want_stats=True)

        def lib():
            flush.fill_(1)
            F.linear(x, w16)
        ms, ms_lib = timeit(ours) - t_flush, timeit(lib) - t_flush
        nbytes = M * (K + N) * 2
        tot["ours"] += ms; tot["lib"] += ms_lib; tot["bytes"] += nbytes
        print(json.dumps({"layer": name, "M": M, "K": K, "N": N, "prologue": pro, "ms": round(ms, 4), "GBps": round(nbytes / ms / 1e6, 1),
                          "frac_of_peak": round(nbytes / ms / 1e6 / peak, 3), "library_gemm_only_ms": round(ms_lib, 4)}), flush=True)
    print(json.dumps({"layer": "ALL", "ms": round(tot["ours"], 3), "library_gemm_only_ms": round(tot["lib"], 3),
                      "GBps": round(tot["bytes"] / tot["ours"] / 1e6, 1), "peak_GBps": peak}))


if __name__ == "__main__":
    main()
