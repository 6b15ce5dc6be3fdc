"""Developer tool: per-kernel time per step from a tools/profile_step.py table (3 profiled steps)."""
import re, sys
def load(path):
    rows = {}
    for l in open(path):
        parts = re.split(r'\s{2,}', l.strip())
        if len(parts) < 11: continue
        name = parts[0]
        try: calls = int(parts[-1])
        except ValueError: continue
        s = parts[6]
        t = float(s[:-2]) if s.endswith('ms') else float(s[:-2]) / 1000 if s.endswith('us') else 0.0
        if t > 0 and not name.startswith(('autograd', '_', 'aten', 'Optimizer')): rows[name[:90]] = (t / 3, calls / 3)
    return rows
a = load(sys.argv[1]); b = load(sys.argv[2]) if len(sys.argv) > 2 else None
keys = sorted(set(a) | set(b or {}), key=lambda k: -max(a.get(k, (0, 0))[0], (b or {}).get(k, (0, 0))[0]))
ta = tb = 0
for k in keys[: int(sys.argv[3]) if len(sys.argv) > 3 else 45]:
    x = a.get(k, (0, 0)); y = (b or {}).get(k, (0, 0))
    print(f"{x[0]:7.3f} ms x{x[1]:5.1f}" + (f"   {y[0]:7.3f} ms x{y[1]:5.1f}  {y[0]-x[0]:+.3f}" if b else "") + f"  {k}")
print("total", sum(v[0] for v in a.values()), sum(v[0] for v in (b or {}).values()))
