"""Opcode histogram of the SASS of one kernel in the built library (run where cuobjdump is installed; no GPU needed).

    python tools/sass_histogram.py <object or .so> <substring of the mangled kernel name> [out.csv]
"""
import csv
import re
import subprocess
import sys
from collections import Counter


def main():
    obj, pattern = sys.argv[1], sys.argv[2]
    text = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    counts, current, kernels = Counter(), None, []
    for line in text.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            current = m.group(1) if pattern in m.group(1) else None
            if current:
                kernels.append(current)
            continue
        if current:
            m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_x]+)*)", line)
            if m:
                counts[m.group(1).split(".")[0] + ("." + m.group(1).split(".")[1] if m.group(1).startswith(("DMMA", "LDG", "LDS", "STS")) and "." in m.group(1) else "")] += 1
    assert len(kernels) == 1, f"pattern must select exactly one kernel, got {kernels}"
    total = sum(counts.values())
    print(kernels[0], total, "instructions")
    rows = [(op, n, f"{100 * n / total:.1f}") for op, n in counts.most_common()]
    for op, n, pct in rows[:30]:
        print(f"{op:16s} {n:6d} {pct:>5s} %")
    if len(sys.argv) > 3:
        with open(sys.argv[3], "w") as fh:
            w = csv.writer(fh)
            w.writerow(["kernel", kernels[0]])
            w.writerow(["opcode", "static_count", "percent"])
            w.writerows(rows)


if __name__ == "__main__":
    main()
