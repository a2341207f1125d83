"""Condenses the per-instruction page of an ncu report into regions: samples, executed instructions and the opcode mix
of every `chunk` consecutive SASS instructions (run here, no GPU needed).

    ncu -i rep.ncu-rep --page source --csv > src.csv;  python tools/ncu_sass_regions.py src.csv [chunk]
"""
import csv
import sys
from collections import Counter


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    hdr = rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    body = rows[2:]
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    total = sum(int(r[col["# Samples"]] or 0) for r in body)
    print(f"{len(body)} instructions, {total} samples")
    for lo in range(0, len(body), chunk):
        part = body[lo:lo + chunk]
        smp = sum(int(r[col["# Samples"]] or 0) for r in part)
        exe = sum(int(r[col["Instructions Executed"]] or 0) for r in part)
        ops = Counter()
        for r in part:
            op = r[col["Source"]].split()
            op = [o for o in op if not o.startswith("@")][0].split(".")[0] if op else "?"
            ops[op] += int(r[col["Instructions Executed"]] or 0)
        st = Counter({s: sum(int(r[col[s]] or 0) for r in part) for s in stalls})
        top = ", ".join(f"{k}:{v}" for k, v in ops.most_common(6))
        tst = ", ".join(f"{k[6:]}:{100 * v // max(smp, 1)}%" for k, v in st.most_common(4))
        print(f"[{lo:5d}] samples {100 * smp / total:5.1f}%  exec {exe / 1e6:7.2f}M  | {top} | {tst}")


if __name__ == "__main__":
    main()
