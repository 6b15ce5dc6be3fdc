"""Developer tool: the row-BatchNorm kernels alone at the camera branch's shapes (B=32), GB/s against the measured peak."""
import os, sys, json, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "lightweight-multi-modal-scene-understanding-via-knowledge-distillation_b200"))
import bench
from src import native
from src.native import call, ptr, lib
dev = torch.device("cuda", 0)
st = native.stream_ptr(dev)
peak = bench.peak_gbs()[0]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
shapes = [(524288, 32), (524288, 192), (131072, 192), (131072, 64), (131072, 384), (32768, 384), (32768, 128), (32768, 768),
          (131072, 128), (131072, 256), (131072, 32)]
t_flush = bench.time_kernel(lambda: flush.fill_(1))
for M, C in shapes:
    x = torch.randn(M, C, device=dev).to(torch.bfloat16); g = torch.randn(M, C, device=dev).to(torch.bfloat16)
    y = torch.empty_like(x)
    f32 = dict(dtype=torch.float32, device=dev)
    sc, sh, mean, invstd = torch.rand(C, **f32) + 0.5, torch.randn(C, **f32) * 0.1, torch.randn(C, **f32) * 0.1, torch.rand(C, **f32) + 0.5
    gamma, beta = torch.rand(C, **f32) + 0.5, torch.randn(C, **f32)
    dg, db = torch.empty(C, **f32), torch.empty(C, **f32)
    ws = torch.empty(lib.kdf_rowbn_bwd_workspace_bytes(C), dtype=torch.uint8, device=dev)
    o = [torch.empty(C, **f32) for _ in range(4)]
    def stats(): call("kdf_rowbn_stats", ptr(x), 1, M, C, ptr(gamma), ptr(beta), None, 1e-5, 0.0, None, None, ptr(o[0]), ptr(o[1]), ptr(o[2]), ptr(o[3]), ptr(ws), st)
    def fwd(): call("kdf_rowbn_apply_fwd", ptr(x), None, 1, M, C, ptr(sc), ptr(sh), 2, ptr(y), st)
    def bwd(): call("kdf_rowbn_bwd", ptr(g), ptr(x), 1, M, C, ptr(sc), ptr(sh), ptr(mean), ptr(invstd), 2, 1, ptr(y), ptr(dg), ptr(db), ptr(ws), st)
    res = {}
    for name, fn, units in (("stats", stats, 1), ("apply_fwd", fwd, 2), ("bwd(reduce+apply)", bwd, 5)):
        def g2():
            flush.fill_(1); fn()
        ms = bench.time_kernel(g2) - t_flush
        res[name] = (ms * 1e3, units * M * C * 2 / ms / 1e6)
    print(f"M={M:7d} C={C:4d} {M*C*2/1e6:6.1f} MB  " + "  ".join(f"{k} {v[0]:6.1f} us {v[1]:5.0f} GB/s ({v[1]/peak:.2f})" for k, v in res.items()), flush=True)
