"""Developer tool: the DESIGN.md section 4d table from a bench.py JSON line."""
import json, sys
d = json.load(open(sys.argv[1]))
print("| kernel | ms | GB/s (algorithmic) | of measured peak | ncu DRAM traffic / algorithmic |")
print("|---|---|---|---|---|")
for k in d["kernels"]:
    t = k.get("traffic")
    tr = f"{t / 1e9:.2f} GB / {k['algorithmic_bytes'] / 1e9:.2f} GB" if t else "—"
    star = "" if k.get("in_step", True) else " (not in the bf16 step)"
    print(f"| {k['kernel']}{star} | {k['ms']:.3f} | {k['achieved']:,.0f} | {k['frac']:.2f} | {tr} |")
print()
print(f"step {d['ms_per_step']:.2f} ms, {d['value']:.0f} frames/s; e2e {d['e2e']['value']:.0f}; launches/step {d['gpu_launches'] / d['steps']:.0f}; cpu {d['cpu_baseline']['value']:.2f} frames/s on {d['cpu_baseline']['cores']} threads")
