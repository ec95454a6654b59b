"""Dev aid: per CUDA source line stall samples / executed instructions from an ncu report captured with --import-source on."""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kern}",
                      "--launch-skip", sys.argv[3] if len(sys.argv) > 3 else "0", "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[hi]; S = hdr.index("# Samples"); I = hdr.index("Instructions Executed")
lines = [(int(r[S]), int(r[I]), r[0], r[1]) for r in rows[hi + 1:] if len(r) == len(hdr) and r[0].isdigit() and r[S].isdigit()]
tot_s, tot_i = sum(l[0] for l in lines), sum(l[1] for l in lines)
print(f"{kern}: {tot_s} samples, {tot_i} warp instructions over {len(lines)} source lines")
for s, i, no, src in sorted(lines, key=lambda l: -l[0])[:int(sys.argv[4]) if len(sys.argv) > 4 else 22]:
    print(f"{100*s/max(tot_s,1):5.1f}% smp {100*i/max(tot_i,1):5.1f}% inst  L{no:>4}  {src.strip()[:105]}")
