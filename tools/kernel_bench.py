"""Developer tool: launch each hand-written kernel once or a few times at the bench shapes (for ncu captures)
and print CUDA-event timings.   python tools/kernel_bench.py [--iters 3] [--only rowbn,bev,kd,fusion]"""
import os, sys, argparse
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "lightweight-multi-modal-scene-understanding-via-knowledge-distillation_b200"))
import torch
from src import native, ops
from src.data_loading.synthetic_frames import make_frames
ap = argparse.ArgumentParser(); ap.add_argument("--iters", type=int, default=3); ap.add_argument("--only", default="rowbn,bev,kd,fusion")
ap.add_argument("--dw-shapes", default="32:128:1,192:128:2,384:64:1,384:64:2,768:32:1,128:64:1,64:64:1,256:64:1")
ap.add_argument("--batch", type=int, default=32); ap.add_argument("--points", type=int, default=170000)
a = ap.parse_args()
dev = torch.device("cuda", 0); p = native.ptr; st = native.stream_ptr(dev)
B, N, C, H, W = a.batch, a.points, 128, 64, 64
dt = torch.bfloat16
def timeit(name, fn, nbytes):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    print(f"{name:34s} {ms:8.3f} ms  {nbytes/ms/1e6:8.0f} GB/s")
only = a.only.split(",")
M = B * N
if "rowbn" in only:
    for Cc in (128, 64):
        x = torch.randn(M, Cc, device=dev, dtype=dt); y = torch.empty_like(x); g = torch.randn_like(x)
        f32 = dict(device=dev, dtype=torch.float32)
        gamma, beta = torch.rand(Cc, **f32) + 0.5, torch.randn(Cc, **f32)
        mean, invstd, scale, shift = (torch.empty(Cc, **f32) for _ in range(4))
        ws = torch.empty(native.lib.kdf_rowbn_bwd_workspace_bytes(Cc), dtype=torch.uint8, device=dev)
        dg, db = torch.empty(Cc, **f32), torch.empty(Cc, **f32)
        nb = x.numel() * 2
        timeit(f"rowbn_stats C={Cc}", lambda: native.call("kdf_rowbn_stats", p(x), 1, M, Cc, p(gamma), p(beta), None, 1e-5, 0.1, None, None, p(mean), p(invstd), p(scale), p(shift), p(ws), st), nb)
        timeit(f"rowbn_apply_fwd C={Cc}", lambda: native.call("kdf_rowbn_apply_fwd", p(x), None, 1, M, Cc, p(scale), p(shift), 1, p(y), st), 2 * nb)
        timeit(f"rowbn_bwd(reduce+apply) C={Cc}", lambda: native.call("kdf_rowbn_bwd", p(g), p(x), 1, M, Cc, p(scale), p(shift), p(mean), p(invstd), 1, 1, p(y), p(dg), p(db), p(ws), st), 5 * nb)
        timeit(f"torch copy C={Cc}", lambda: y.copy_(x), 2 * nb)
        del x, y, g
if "bev" in only:
    pts = make_frames(B, N, seed=1, device=dev)["points"]
    geom = ops.bev_range_constants([-50, -50, -5, 50, 50, 3])
    feats = torch.rand(B, N, C, device=dev, dtype=dt)
    grid = torch.empty(B, H, W, C, dtype=dt, device=dev); cnt = torch.empty(B, H * W, dtype=torch.int32, device=dev)
    cel = torch.empty(B, N, dtype=torch.int32, device=dev); ties = torch.empty(B, H * W, C, dtype=torch.int32, device=dev)
    order = torch.empty(B, N, dtype=torch.int32, device=dev); offs = torch.empty(B, H * W + 1, dtype=torch.int32, device=dev)
    wsb = native.lib.kdf_bev_workspace_bytes(B, N, H, W); ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    v = 0.622
    timeit("bev_project_fwd (tie counts)", lambda: native.call("kdf_bev_project_fwd", p(pts), 4, p(feats), 1, B, N, C, *geom, H, W, 0, p(grid), p(cnt), p(cel), p(ties), p(order), p(offs), p(ws), wsb, st), B * (16 * N + C * 2 * v * N + C * 2 * H * W))
    timeit("bev_bwd (tie counts)", lambda: native.call("kdf_bev_project_bwd", p(torch.rand(B, H * W, C, device=dev, dtype=dt)), p(feats), p(grid), p(ties), None, p(cel), p(order), p(offs), 1, B, N, C, H, W, 0, p(torch.empty(B, N, C, dtype=dt, device=dev)), st), B * (C * 2 * v * N + C * 2 * N))
    timeit("bev_project_fwd", lambda: native.call("kdf_bev_project_fwd", p(pts), 4, p(feats), 1, B, N, C, *geom, H, W, 0, p(grid), p(cnt), p(cel), None, p(order), p(offs), p(ws), wsb, st), B * (16 * N + C * 2 * v * N + C * 2 * H * W))
    timeit("bev_reduce (max)", lambda: native.call("kdf_bev_reduce", p(feats), 1, p(order), p(offs), B, N, C, H, W, 0, p(grid), None, st), B * (C * 2 * v * N + C * 2 * H * W))
    gg = torch.rand(B, H * W, C, device=dev, dtype=dt); gf = torch.empty(B, N, C, dtype=dt, device=dev)
    timeit("bev_bwd (max)", lambda: native.call("kdf_bev_project_bwd", p(gg), p(feats), p(grid), None, None, p(cel), p(order), p(offs), 1, B, N, C, H, W, 0, p(gf), st), B * (2 * C * 2 * v * N + C * 2 * (1 - v) * N + 2 * C * 2 * H * W + 8 * N))
    timeit("bev_index", lambda: native.call("kdf_bev_index", p(pts), B, N, 4, *geom, H, W, p(cel), None, p(cnt), st), B * 20 * N)
    del feats, gf
if "kd" in only:
    zs = torch.randn(B, 2, H, W, device=dev, dtype=dt); zt = torch.randn_like(zs)
    lab = (torch.rand(B, H, W, device=dev) < 0.13).long(); cw = torch.tensor([0.4, 3.5], device=dev)
    ts = [torch.randn(B, H, W, C, device=dev, dtype=dt).permute(0, 3, 1, 2) for _ in range(2)]
    tt = [torch.randn(B, H, W, C, device=dev, dtype=dt).permute(0, 3, 1, 2) for _ in range(2)]
    timeit("kd_loss", lambda: ops.kd_loss_fwd_bwd(zs, zt, lab, cw, ts, tt), B * H * W * (2 * 3 * C * 2))
    timeit("kd_loss (hoisted label count)", lambda: ops.kd_loss_fwd_bwd(zs, zt, lab, cw, ts, tt, counted_ws=ops.kd_label_count(lab, 2)), B * H * W * (2 * 3 * C * 2))
if "fusion" in only:
    Mp = B * H * W
    f32 = dict(device=dev, dtype=torch.float32)
    cam, lid = torch.randn(Mp, C, device=dev, dtype=dt), torch.randn(Mp, C, device=dev, dtype=dt)
    sc = [torch.rand(C, **f32) + 0.5 for _ in range(2)]; sh = [torch.randn(C, **f32) * 0.1 for _ in range(2)]
    w1, b1 = torch.randn(C, 2 * C, **f32) * 0.05, torch.randn(C, **f32) * 0.1
    w2, b2 = torch.randn(2, C, **f32) * 0.1, torch.randn(2, **f32) * 0.1
    fo = torch.empty(Mp, C, dtype=dt, device=dev); attn = torch.empty(Mp, 2, **f32)
    timeit("fusion_weighted_fwd", lambda: native.call("kdf_fusion_weighted_fwd", p(cam), p(lid), 1, Mp, C, p(sc[0]), p(sh[0]), p(sc[1]), p(sh[1]), p(w1), p(b1), p(w2), p(b2), p(fo), p(attn), st), Mp * 3 * C * 2)
    go = torch.randn(Mp, C, device=dev, dtype=dt); g1, g2 = torch.empty_like(cam), torch.empty_like(lid)
    gaff, gw1, gb1 = torch.empty(4, C, **f32), torch.empty(C, 2 * C, **f32), torch.empty(C, **f32)
    gw2, gb2 = torch.empty(2, C, **f32), torch.empty(2, **f32)
    timeit("fusion_weighted_bwd", lambda: native.call("kdf_fusion_weighted_bwd", p(go), p(cam), p(lid), 1, Mp, C, p(sc[0]), p(sh[0]), p(sc[1]), p(sh[1]), p(w1), p(b1), p(w2), p(b2), p(attn), p(g1), p(g2), p(gaff), p(gw1), p(gb1), p(gw2), p(gb2), st), Mp * 5 * C * 2)
if "pw" in only:
    # three representative 1x1-convolution layers of the student's forward (tools/pw_conv_bench.py has all sixteen)
    for name, Mp, K, Nn, pro in (("pw stage2 expand 32->192", 524288, 32, 192, False), ("pw stage3 project 384->64 +BN/ReLU6", 131072, 384, 64, True),
                                 ("pw fpn post 128->128 +BN/ReLU", 131072, 128, 128, True)):
        x = torch.randn(Mp, K, device=dev).to(dt)
        w = torch.randn(Nn, K, device=dev) * (2.0 / K) ** 0.5
        pack = ops._pw_pack_factor(K, Nn)
        wb = ops.pw_conv_weight(w, pack)
        sc, sh = torch.rand(K, device=dev) + 0.5, torch.randn(K, device=dev) * 0.1
        timeit(name, lambda: ops.pw_conv_fwd(x, wb, pack, pro=(sc, sh, 2) if pro else None, want_stats=True), Mp * (K + Nn) * 2)
        del x
if "mlp" in only:
    zprev = torch.randn(M, 128, device=dev, dtype=dt)
    sc, sh = torch.rand(128, device=dev) + 0.5, torch.randn(128, device=dev) * 0.1
    W3 = (torch.randn(128, 128, device=dev) / 11).to(dt)
    timeit("mlp_layer_fwd mode1 (L3)", lambda: ops.mlp_layer_fwd(1, zprev, sc, sh, W3), M * 512)
    pts2 = torch.randn(M, 4, device=dev)
    q, r = torch.randn(64, 4, device=dev) * 0.02, torch.randn(64, device=dev) * 0.1
    W2 = (torch.randn(128, 64, device=dev) / 8).to(dt)
    timeit("mlp_layer_fwd mode0 (L1+L2)", lambda: ops.mlp_layer_fwd(0, pts2, q, r, W2), M * 272)
    from src import point_mlp as _pm
    timeit("mlp_eval3 (L1+L2+L3, eval)", lambda: _pm.mlp_eval3_fwd(pts2, q, r, W2, sc, sh, W3), M * 272)
    dy = (torch.randn(M, 128, device=dev) * (torch.rand(M, 128, device=dev) < 0.05)).to(dt)
    z3 = torch.randn(M, 128, device=dev, dtype=dt)
    gs, ga, gb = torch.rand(128, device=dev) + 0.5, torch.randn(128, device=dev) * 0.01, torch.randn(128, device=dev) * 0.01
    timeit("mlp_layer_bwd mode1 (L3)", lambda: ops.mlp_layer_bwd(1, dy, z3, gs, ga, gb, zprev, sc, sh, W3), M * 1024)
    timeit("mlp_layer_bwd mode0 (L2+L1)", lambda: ops.mlp_layer_bwd(0, dy, z3, gs, ga, gb, pts2, q, r, W2), M * 528)
    del dy
if "affine" in only:
    from src import point_mlp
    pts = make_frames(B, N, seed=1, device=dev)["points"]
    geom = ops.bev_range_constants([-50, -50, -5, 50, 50, 3])
    z3 = torch.randn(M, 128, device=dev, dtype=dt)
    sc, sh = torch.rand(128, device=dev) + 0.5, torch.randn(128, device=dev) * 0.1
    v = 0.622
    timeit("bev_build_order", lambda: point_mlp.bev_build_order(pts, geom, (H, W)), B * N * (16 + 4 + 4 + 4 + 4))
    _c, _k, _o, _f = point_mlp.bev_build_order(pts, geom, (H, W))
    _nb = native.lib.kdf_bev_workspace_bytes(B, N, H, W); _ws = torch.empty(_nb, dtype=torch.uint8, device=dev)
    timeit("bev_build_order (preallocated)", lambda: native.call("kdf_bev_build_order", p(pts), 4, B, N, *geom, H, W, p(_k), p(_c), p(_o), p(_f), p(_ws), _nb, st), B * N * (16 + 4 + 4 + 4 + 4))
    cell, count, order, offsets = point_mlp.bev_build_order(pts, geom, (H, W))
    timeit("bev_reduce_affine (+extreme)", lambda: point_mlp.bev_reduce_affine(z3, sc, sh, order, offsets, B, N, (H, W), True), B * (C * 2 * v * N + C * 2 * H * W))
    timeit("bev_reduce_affine", lambda: point_mlp.bev_reduce_affine(z3, sc, sh, order, offsets, B, N, (H, W), False), B * (C * 2 * v * N + C * 2 * H * W))
    grid, grid_z = point_mlp.bev_reduce_affine(z3, sc, sh, order, offsets, B, N, (H, W), True)
    gg = torch.randn(B, H, W, C, device=dev, dtype=dt)
    timeit("bev_bwd_affine", lambda: point_mlp.bev_bwd_affine(gg, z3, grid, grid_z, order, offsets, cell, B, N, (H, W), zero_outside=False), B * (2 * C * 2 * v * N + 3 * C * 2 * H * W))
    timeit("point_moments", lambda: point_mlp.point_moments(pts), B * N * 16)
if "sorted" in only:
    from src import point_mlp
    pts = make_frames(B, N, seed=1, device=dev)["points"]
    geom = ops.bev_range_constants([-50, -50, -5, 50, 50, 3])
    v = 0.622
    timeit("bev_build_order", lambda: point_mlp.bev_build_order(pts, geom, (H, W)), B * N * (16 + 4 + 4 + 4 + 4))
    timeit("bev_build_sorted", lambda: point_mlp.bev_build_sorted(pts, geom, (H, W)), B * N * (16 + 4 + 4 + 4 + 16 + 16 + 4))
    cell0, count0, order0, offsets0 = point_mlp.bev_build_order(pts, geom, (H, W))
    cell, count, offsets, spts, cs, _ = point_mlp.bev_build_sorted(pts, geom, (H, W))
    z3 = torch.randn(M, 128, device=dev, dtype=dt)
    sc, sh = torch.rand(128, device=dev) + 0.5, torch.randn(128, device=dev) * 0.1
    timeit("bev_reduce_affine indexed", lambda: point_mlp.bev_reduce_affine(z3, sc, sh, order0, offsets0, B, N, (H, W), True), B * (C * 2 * v * N + C * 2 * H * W))
    timeit("bev_reduce_affine contiguous", lambda: point_mlp.bev_reduce_affine(z3, sc, sh, None, offsets, B, N, (H, W), True), B * (C * 2 * v * N + C * 2 * H * W))
    grid, grid_z = point_mlp.bev_reduce_affine(z3, sc, sh, None, offsets, B, N, (H, W), True)
    gg = torch.randn(B, H, W, C, device=dev, dtype=dt)
    grid0, grid_z0 = point_mlp.bev_reduce_affine(z3, sc, sh, order0, offsets0, B, N, (H, W), True)
    timeit("bev_bwd_affine (dy rows)", lambda: point_mlp.bev_bwd_affine(gg, z3, grid0, grid_z0, order0, offsets0, cell0, B, N, (H, W), zero_outside=False), B * (2 * C * 2 * v * N + 3 * C * 2 * H * W))
    timeit("bev_bwd_share (share + bits)", lambda: point_mlp.bev_bwd_share(gg, z3, grid, grid_z, offsets, B, N, (H, W)), B * (C * 2 * v * N + 16 * v * N + 4 * C * 2 * H * W))
    share, bits, _s = point_mlp.bev_bwd_share(gg, z3, grid, grid_z, offsets, B, N, (H, W))
    dy0, _s0 = point_mlp.bev_bwd_affine(gg, z3, grid0, grid_z0, order0, offsets0, cell0, B, N, (H, W), zero_outside=False)
    zprev = torch.randn(M, 128, device=dev, dtype=dt)
    W3 = (torch.randn(128, 128, device=dev) / 11.3).to(dt)
    gs, ga, gb = torch.rand(128, device=dev) + 0.5, torch.randn(128, device=dev) * 0.01, torch.randn(128, device=dev) * 0.01
    timeit("mlp_layer_bwd mode1 (dy rows)", lambda: ops.mlp_layer_bwd(1, dy0, z3, gs, ga, gb, zprev, sc, sh, W3, row_cell=cell0.view(-1)), M * 1024)
    timeit("mlp_layer_bwd_share", lambda: point_mlp.mlp_layer_bwd_share(cs.view(-1), share, bits, z3, gs, ga, gb, zprev, sc, sh, W3), M * (768 + 16 + 4))
    del z3, zprev, dy0
if "stem" in only:
    img = torch.rand(B, 3, 256, 256, device=dev); w = torch.randn(32, 3, 3, 3, device=dev) * 0.3
    out = torch.empty(B, 128, 128, 32, device=dev, dtype=dt); sts = torch.empty(2, 32, dtype=torch.float64, device=dev)
    gw = torch.empty(32, 27, device=dev); gr = torch.randn(B, 128, 128, 32, device=dev, dtype=dt)
    scl, shf = torch.rand(32, device=dev) + 0.5, torch.randn(32, device=dev)
    nb = img.numel() * 4 + out.numel() * 2
    timeit("stem fwd + stats", lambda: native.call("kdf_stem_conv_fwd", p(img), p(w), B, 256, 256, None, None, 0, p(out), p(sts), st), nb)
    timeit("stem fwd + folded BN + ReLU6", lambda: native.call("kdf_stem_conv_fwd", p(img), p(w), B, 256, 256, p(scl), p(shf), 2, p(out), None, st), nb)
    timeit("stem wgrad", lambda: native.call("kdf_stem_conv_bwd_weight", p(img), p(gr), B, 256, 256, p(gw), st), nb)
    wb = w.to(dt)
    with torch.autocast("cuda", dtype=dt):
        timeit("library: channels_last + cast + cudnn fwd", lambda: torch.nn.functional.conv2d(img.contiguous(memory_format=torch.channels_last), w, None, 2, 1), nb)
if "cls" in only:
    Mc = B * H * W
    xr = torch.randn(Mc, 32, device=dev).to(dt); wc = torch.randn(2, 32, device=dev) * 0.3; bc = torch.randn(2, device=dev)
    lo = torch.empty(B, 2, H, W, device=dev, dtype=dt); dl = torch.randn(B, 2, H, W, device=dev).to(dt)
    dxr = torch.empty_like(xr); gwc = torch.empty(2, 32, device=dev); gbc = torch.empty(2, device=dev)
    timeit("cls_conv fwd", lambda: native.call("kdf_cls_conv_fwd", p(xr), p(wc), p(bc), Mc, 32, 2, H * W, p(lo), st), Mc * 68)
    timeit("cls_conv bwd", lambda: native.call("kdf_cls_conv_bwd", p(xr), p(dl), p(wc), Mc, 32, 2, H * W, p(dxr), p(gwc), p(gbc), st), Mc * 132)
if "dw" in only:
    import torch.nn as nn
    for (Cc, Hh, st_) in (tuple(int(v) for v in t.split(":")) for t in a.dw_shapes.split(",")):
        conv = nn.Conv2d(Cc, Cc, 3, stride=st_, padding=1, groups=Cc, bias=False).to(dev)
        x = torch.randn(B, Cc, Hh, Hh, device=dev, dtype=dt).contiguous(memory_format=torch.channels_last).requires_grad_(True)
        y = ops.dwconv3x3(conv, x)
        g = torch.randn_like(y)
        nb_in, nb_out = x.numel() * 2, y.numel() * 2
        with torch.no_grad():
            timeit(f"dwconv fwd C={Cc} {Hh}x{Hh} s{st_}", lambda: ops.dwconv3x3(conv, x), nb_in + nb_out)
        w9 = conv.weight.detach().reshape(Cc, 9).float().contiguous()
        xr_ = x.detach().permute(0, 2, 3, 1); yo = torch.empty_like(y.permute(0, 2, 3, 1)); sts = torch.empty(2, Cc, dtype=torch.float64, device=dev)
        timeit(f"dwconv fwd+stats C={Cc} s{st_} (C ABI)", lambda: native.call("kdf_dwconv3x3_fwd", p(xr_), p(w9), 1, B, Hh, Hh, Cc, st_, 0, p(yo), p(sts), st), nb_in + nb_out)
        gx = torch.empty(B, Hh, Hh, Cc, device=dev, dtype=dt); gw = torch.empty(Cc, 9, device=dev)
        xr, gr = x.detach().permute(0, 2, 3, 1), g.permute(0, 2, 3, 1)
        timeit(f"dwconv dgrad C={Cc} s{st_}", lambda: native.call("kdf_dwconv3x3_bwd_data", p(gr), p(w9), 1, B, Hh, Hh, Cc, st_, p(gx), st), nb_in + nb_out)
        timeit(f"dwconv wgrad C={Cc} s{st_}", lambda: native.call("kdf_dwconv3x3_bwd_weight", p(xr), p(gr), 1, B, Hh, Hh, Cc, st_, p(gw), st), nb_in + nb_out)
        wb = conv.weight.detach().to(dt)
        timeit(f"cudnn fwd C={Cc} s{st_}", lambda: torch.nn.functional.conv2d(x.detach(), wb, None, st_, 1, 1, Cc), nb_in + nb_out)
        del x, y, g, gx
