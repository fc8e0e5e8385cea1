"""Top stall sites of an `ncu --set full --import-source on` capture, from `ncu -i X --page source --csv`.
usage: python tools/ncu_hot.py gpurun_out/X.ncu-rep [N]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
def f(r, k):
    try: return float(r[col[k]])
    except Exception: return 0.0
tot = sum(f(r, "# Samples") for r in data)
tot_inst = sum(f(r, "Instructions Executed") for r in data)
print(f"total samples {tot:.0f}, warp instructions {tot_inst:.0f}, SASS lines {len(data)}")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(f(r, s) for r in data) for s in stalls}
print("stall reasons (all samples): " + ", ".join(f"{k[6:]} {v / tot * 100:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]))
idx = sorted(range(len(data)), key=lambda i: -f(data[i], "# Samples"))[:N]
for i in sorted(idx):
    r = data[i]
    top = max(stalls, key=lambda s: f(r, s))
    print(f"{i:5d} {f(r, '# Samples') / tot * 100:5.2f}%  exec {f(r, 'Instructions Executed'):12.0f}  thr {f(r, 'Avg. Threads Executed'):4.1f}  {top[6:]:12s} {r[col['Source']][:110]}")
