"""Developer tool: cut ONE full training step (between the last two adamw_flat_kernel launches) out of an
`ncu --metrics gpu__time_duration.sum --csv` launch list and write the per-kernel share summary.
    python tools/launch_list_step.py gpurun_out/launches.csv profiles/r01_ncu_launch_list.csv profiles/r01_ncu_launch_list_summary.csv "<command line>" """
import collections, csv, re, sys
src, out_list, out_sum, cmd = sys.argv[1:5]
rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
ad = [i for i, r in enumerate(rows) if "adamw_flat_kernel" in r[ki]]
step = rows[ad[-2] + 1:ad[-1] + 1]
def short(n):
    n = re.sub(r"^void ", "", n).replace("kdf::", "")
    return re.sub(r"\(.*$", "", n)[:90]
tot = sum(float(r[vi]) for r in step) / 1000
agg = collections.defaultdict(lambda: [0, 0.0])
for r in step:
    a = agg[short(r[ki])]
    a[0] += 1
    a[1] += float(r[vi]) / 1000
out = [f"# {cmd}",
       f"# final round-1 build: the LAST FULL eager training step (between the last two adamw_flat_kernel launches): {len(step)} kernel launches, {tot:.0f} us in total",
       "# cold-cache, serialised per-launch times: compare SHARES of the step, not absolutes",
       "kernel,launches,total_us,share"]
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f'"{k}",{c},{t:.1f},{t / tot:.4f}')
open(out_sum, "w").write("\n".join(out) + "\n")
with open(out_list, "w") as f:
    w = csv.writer(f, quoting=csv.QUOTE_ALL)
    w.writerow(hdr)
    w.writerows(step)
print("\n".join(out[:16]))
