"""Condense an `ncu --page raw --csv` export into the per-kernel figures quoted in DESIGN.md/profiles.
usage: python tools/ncu_summary.py raw.csv [out.json]"""
import csv
import json
import sys

WANT = {
    "gpu__time_duration.sum": "time",
    "launch__registers_per_thread": "regs",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "pipe_fp64_pct",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "pipe_xu_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "pipe_alu_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "pipe_fma_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "pipe_lsu_pct",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed": "l1_lsu_wavefronts_pct",
    "lts__t_sectors.avg.pct_of_peak_sustained_elapsed": "l2_sectors_pct",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "smsp__inst_executed.sum": "inst",
}


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return None


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in data:
        d = {"kernel": r[ix["Kernel Name"]].split("(")[0].replace("void ", ""), "grid": r[ix["Grid Size"]]}
        for k, name in WANT.items():
            if k in ix:
                d[name] = num(r[ix[k]])
                if name in ("time", "dram_read", "dram_write"):
                    d[name + "_unit"] = units[ix[k]]
        stalls = []
        for h in hdr:
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                v = num(r[ix[h]])
                if v:
                    stalls.append((v, h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
        d["top_stalls_per_issue"] = {n: round(v, 2) for v, n in sorted(stalls, reverse=True)[:5]}
        out.append(d)
    seen = set()
    for d in out:
        key = (d["kernel"], d["grid"])
        if key in seen:
            continue
        seen.add(key)
        print(json.dumps(d))
    if len(sys.argv) > 2:
        json.dump(out, open(sys.argv[2], "w"), indent=1)


if __name__ == "__main__":
    main()
