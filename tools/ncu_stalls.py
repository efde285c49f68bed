"""Per-instruction stall samples of one kernel from `ncu --page source --csv` (SASS view).
usage: python tools/ncu_stalls.py src.csv [top_n]"""
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    data = [r for r in rows[hi + 1:] if len(r) >= len(hdr) - 2 and r[0].startswith("0x")]
    ix = {h: i for i, h in enumerate(hdr)}
    S = ix["# Samples"]
    tot = sum(int(r[S] or 0) for r in data)
    cats = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = {c: sum(int(r[ix[c]] or 0) for r in data) for c in cats}
    print("samples", tot, "instructions", len(data))
    for c, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]:
        print(f"  {c:24s} {v:8d} {100 * v / tot:5.1f} %")
    for k, r in enumerate(data):
        r.append(k)
    for r in sorted(data, key=lambda r: -int(r[S] or 0))[:n]:
        st = {c: int(r[ix[c]] or 0) for c in cats}
        big = sorted(st.items(), key=lambda kv: -kv[1])[:2]
        print(f"{r[-1]:5d} {r[S]:>6s} {r[ix['Source']].strip()[:64]:64s} {big}")


if __name__ == "__main__":
    main()
