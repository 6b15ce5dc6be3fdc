"""Developer tool: profiles/*_ncu_traffic.json (DRAM bytes per launch of each hand-written kernel, keyed by the names
bench.py uses) from an `ncu --page raw --csv` export of tools/kernel_bench.py.
    python tools/ncu_traffic.py gpurun_out/ncu_full_final2_raw.csv profiles/r01_ncu_traffic.json profiles/r01_ncu_full_final2.csv"""
import csv, json, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
ik, ir, iw = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
def to_bytes(v, u):
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
last, seq = {}, []
for r in rows[2:]:
    last[r[ik]] = to_bytes(r[ir], units[ir]) + to_bytes(r[iw], units[iw])       # the last (warm) launch of each kernel
    seq.append((r[ik], last[r[ik]]))
# pw_conv_fwd_kernel runs three layer shapes in tools/kernel_bench.py (in this order, `iters`+1 launches each): last launch of each
pw = [v for k, v in seq if "pw_conv_fwd_kernel" in k]
per = len(pw) // 3 if pw else 0
def find(sub, *more):
    out = [v for k, v in last.items() if sub in k and all(m in k for m in more)]
    assert len(out) == 1, (sub, more, [k[:60] for k in last if sub in k])
    return out[0]
order3 = find("bev_chunk_count_kernel") + find("bev_chunk_scan_kernel") + find("bev_chunk_fill_kernel")
m = {
    "mlp_layer_fwd_kernel<0>": find("mlp_layer_fwd_kernel<0"),
    "mlp_layer_fwd_kernel<1>": find("mlp_layer_fwd_kernel<1"),
    "mlp_eval3_kernel": find("mlp_eval3_kernel"),
    "mlp_layer_bwd_kernel<1>": find("mlp_layer_bwd_kernel<1"),
    "mlp_layer_bwd_kernel<0>": find("mlp_layer_bwd0_tma_kernel"),
    "bev_build_order (index+scan+fill)": order3,
    "bev_reduce_affine_kernel": find("bev_reduce_affine_kernel"),
    "bev_bwd_affine_kernel": find("bev_bwd_affine_kernel"),
    "bev_project_fwd (index+scan+fill+max)": order3 + find("bev_reduce_max_kernel"),
    "bev_bwd_max_kernel": find("bev_bwd_max_kernel"),
    "bev_index_kernel": find("bev_index_hist_kernel"),
    "fusion_weighted_fwd_kernel": find("fusion_weighted_fwd_tma_kernel"),
    "fusion_weighted_bwd_kernel": find("fusion_weighted_bwd_tc_kernel"),
    "kd_loss_kernel (label histogram + loss fwd+bwd)": find("kd_loss_kernel") + find("kd_label_count_kernel"),
    "kd_loss_kernel (loss fwd+bwd; label histogram taken on the side stream when the batch arrives)": find("kd_loss_kernel"),
    "kd_label_count (memset + histogram, side stream)": find("kd_label_count_kernel"),
}
if per:
    m["pw_conv_fwd stage2 expand 32->192 @128x128 (rows + statistics)"] = pw[per - 1]
    m["pw_conv_fwd stage3 project 384->64 @64x64 (BN+ReLU6 prologue, rows + statistics)"] = pw[2 * per - 1]
    m["pw_conv_fwd fpn post 128->128 @64x64 (BN+ReLU prologue, rows + statistics)"] = pw[3 * per - 1]
json.dump({"what": "dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full --clock-control none, "
                   "tools/kernel_bench.py --iters 1 --only bev,mlp,fusion,kd,affine,pw at the bench shapes (B=32, N=170000, C=128, 64x64, bf16)",
           "captures": [sys.argv[3]], "bytes_per_launch": {k: int(v) for k, v in m.items()}}, open(sys.argv[2], "w"), indent=1)
print(json.dumps(m, indent=1))
