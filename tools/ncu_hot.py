"""Developer tool: top stall / shared-memory-conflict SASS lines from `ncu --page source --csv` output."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr = rows[1]
ia, isamp, iex = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
iexc = hdr.index("L1 Wavefronts Shared Excessive")
def num(x):
    try: return int(float(x))
    except ValueError: return 0
data = [(num(r[isamp]), num(r[iex]), num(r[iexc]), r[ia].strip(), i) for i, r in enumerate(rows[2:]) if len(r) > max(isamp, iexc)]
print(rows[0][1], "| total samples", sum(d[0] for d in data), "| warp instructions", sum(d[1] for d in data))
print("--- top stall lines: idx samples executed excess_smem sass")
for d in sorted(data, reverse=True)[:top]:
    print(d[4], d[0], d[1], d[2], d[3][:100])
print("--- top excessive shared wavefronts")
for d in sorted(data, key=lambda d: -d[2])[:10]:
    print(d[4], d[0], d[1], d[2], d[3][:100])
