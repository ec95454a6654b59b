"""Dev aid: per-kernel stall summary + hottest SASS lines from an ncu report's source page."""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}", "--launch-skip", skip, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
data = []
for r in rows[2:]:
    if len(r) != len(hdr) or not r[hdr.index("# Samples")].isdigit():
        break   # a second kernel's table follows: keep the first
    data.append(r)
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = {s: sum(int(r[ix[s]]) for r in data) for s in stalls}
ns = sum(int(r[ix["# Samples"]]) for r in data)
print(rows[0][1][:100]); print("samples", ns, "instructions", len(data))
print("  ".join(f"{s[6:]}={100*v/max(ns,1):.0f}%" for s, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]))
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:22]:
    top = max(stalls, key=lambda s: int(r[ix[s]]))
    print(f'{r[ix["# Samples"]]:>6} {top[6:]:<14} {r[ix["Source"]].strip()[:90]}')
