"""Developer tool: one line per launch from `ncu --page raw --csv` with the metrics the roofline notes use."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
want = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "ms"), ("dram__bytes_read.sum", "rd"), ("dram__bytes_write.sum", "wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"), ("launch__registers_per_thread", "regs"),
        ("launch__grid_size", "grid"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bankconf"), ("lts__t_sector_hit_rate.pct", "l2hit%")]
idx = [(n, hdr.index(k)) for k, n in want if k in hdr]
for r in rows[2:]:
    out = []
    for n, i in idx:
        v = r[i]
        if n == "kernel":
            v = v[:48]
        else:
            try:
                v = f"{float(v):.4g}"
            except ValueError:
                pass
            v = f"{n}={v}{units[i] if n in ('rd', 'wr') else ''}"
        out.append(v)
    print("  ".join(out))
