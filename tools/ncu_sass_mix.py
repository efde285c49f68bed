"""Per-opcode instruction mix and stall samples from an `ncu --page source --csv` export (SASS view).
usage: python tools/ncu_sass_mix.py source.csv [n_top]"""
import collections
import csv
import re
import sys


def main():
    r = csv.reader(open(sys.argv[1]))
    ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    hdr, ix = None, None
    ops, samp, rows, tot, ts = collections.Counter(), collections.Counter(), [], 0, 0
    for row in r:
        if row and row[0] == "Address":
            hdr, ix = row, {h: i for i, h in enumerate(row)}
            continue
        if hdr is None or len(row) < len(hdr) - 2:
            continue
        src = row[ix["Source"]]
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
        op = m.group(2).split(".")[0] if m else "?"
        try:
            n, s = int(row[ix["Instructions Executed"]] or 0), int(row[ix["# Samples"]] or 0)
        except ValueError:
            continue
        ops[op] += n
        samp[op] += s
        tot += n
        ts += s
        rows.append((s, n, src, row))
    print("total warp-instructions", tot, "samples", ts)
    for op, n in ops.most_common(25):
        print(f"{op:10s} inst {n / tot * 100:5.1f}%  samples {samp[op] / max(ts, 1) * 100:5.1f}%")
    print("--- top sampled instructions")
    for s, n, src, row in sorted(rows, key=lambda t: -t[0])[:ntop]:
        st = {k: int(row[ix[k]] or 0) for k in hdr if k.startswith("stall_") and "Not" not in k}
        top = sorted(st.items(), key=lambda x: -x[1])[:2]
        print(s, n, src[:80], top)


if __name__ == "__main__":
    main()
