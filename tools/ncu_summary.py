"""Summarise `ncu --set full` reports into profiles/*.json (one record per captured launch, key metrics only).

    python tools/ncu_summary.py gpurun_out/full_c2.ncu-rep:c2 gpurun_out/full_c3.ncu-rep:c3 --out profiles/r01_ncu_full_summary.json

Also refreshes profiles/traffic.json: DRAM bytes (read + write) per launch of the attention kernel per workload,
which bench.py copies into roofline.traffic."""
import argparse, csv, io, json, os, subprocess, sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
           "launch__shared_mem_per_block_static", "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
           "lts__t_bytes.sum", "l1tex__t_bytes.sum"]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def load(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    recs = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                d[m] = f"{r[i]} {units[i]}".strip()
        recs.append(d)
    return recs


def to_bytes(s):
    v, u = s.split()
    return float(v.replace(",", "")) * UNIT.get(u, 1.0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("reports", nargs="+", help="path.ncu-rep:label")
    ap.add_argument("--out", required=True)
    ap.add_argument("--traffic", default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json"))
    a = ap.parse_args()
    allr, traffic = [], {}
    try:
        traffic = json.load(open(a.traffic))
    except (OSError, ValueError):
        pass
    for spec in a.reports:
        rep, label = spec.rsplit(":", 1)
        recs = load(rep)
        for r in recs:
            r["capture"] = label
        allr += recs
        att = [r for r in recs if "memory_read_umma" in r["kernel"]]
        if att:
            b = [to_bytes(r["dram__bytes_read.sum"]) + to_bytes(r["dram__bytes_write.sum"]) for r in att]
            traffic[label] = {"kernel": att[0]["kernel"], "dram_bytes_per_launch": sum(b) / len(b), "launches": len(b), "report": os.path.basename(rep)}
    json.dump(allr, open(a.out, "w"), indent=1)
    json.dump(traffic, open(a.traffic, "w"), indent=1)
    for r in allr:
        print(r["capture"], r["kernel"][:60], r.get("gpu__time_duration.sum"), r.get("dram__bytes_read.sum"), r.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"))


if __name__ == "__main__":
    main()
