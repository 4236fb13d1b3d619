"""Summarise `ncu -i X.ncu-rep --page source --csv`: stall totals, the hottest SASS instructions, instruction mix."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
for i, r in enumerate(rows[:30]):
    if "Address" in r and "Source" in r:
        hdr, start = r, i
        break
ix = {h: k for k, h in enumerate(hdr)}
body = [r for r in rows[start + 1:] if len(r) == len(hdr)]
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
tot_s = sum(f(r, "# Samples") for r in body)
tot_i = sum(f(r, "Instructions Executed") for r in body)
print("samples", tot_s, "warp instructions", tot_i, "SASS lines", len(body))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("stall totals:", ", ".join("%s %.1f%%" % (h[6:], 100 * sum(f(r, h) for r in body) / max(tot_s, 1)) for h in sorted(stalls, key=lambda h: -sum(f(r, h) for r in body))[:8]))
mix = collections.Counter()
for r in body:
    op = r[ix["Source"]].split()
    op = [o for o in op if not o.startswith("@")]
    mix[op[0].split(".")[0] if op else "?"] += f(r, "Instructions Executed")
print("mix:", ", ".join("%s %.1f%%" % (k, 100 * v / tot_i) for k, v in mix.most_common(18)))
print("excess global sectors:", sum(f(r, "L2 Theoretical Sectors Global Excessive") for r in body), "of", sum(f(r, "L2 Theoretical Sectors Global") for r in body))
print("excess shared wavefronts:", sum(f(r, "L1 Wavefronts Shared Excessive") for r in body), "of", sum(f(r, "L1 Wavefronts Shared") for r in body))
for r in sorted(body, key=lambda r: -f(r, "# Samples"))[:top]:
    dom = max(stalls, key=lambda h: f(r, h))
    print("%6.2f%%  exec %10d  %-14s exg %9d  %s" % (100 * f(r, "# Samples") / tot_s, f(r, "Instructions Executed"), dom[6:], f(r, "L2 Theoretical Sectors Global Excessive"), r[ix["Source"]][:110]))
