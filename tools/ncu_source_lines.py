"""Per-source-line warp-stall samples from an `ncu --page source --csv --print-source cuda,sass` export.
usage: python tools/ncu_source_lines.py export.csv [top_n]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 28
secs = [{"name": r[1], "start": i} for i, r in enumerate(rows) if r and r[0] == "Function Name"]
for k, s in enumerate(secs):
    end = secs[k + 1]["start"] if k + 1 < len(secs) else len(rows)
    hdr = rows[s["start"] + 1]
    li = hdr.index("Line No")
    src = [i for i, h in enumerate(hdr) if h == "Source"][0]
    samp = hdr.index("# Samples")
    stall_cols = [(h, i) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    per_line, text, per_stall, tot = collections.Counter(), {}, collections.defaultdict(collections.Counter), 0
    for r in rows[s["start"] + 2:end]:
        if len(r) <= samp:
            continue
        try:
            n = int(r[samp])
        except ValueError:
            continue
        per_line[r[li]] += n
        tot += n
        text.setdefault(r[li], r[src][:100])
        for h, i in stall_cols:
            try:
                per_stall[r[li]][h] += int(r[i])
            except ValueError:
                pass
    print("=====", s["name"][:100], "samples", tot)
    for ln, n in per_line.most_common(top):
        print(f"{ln:>5} {100 * n / max(tot, 1):5.1f}%  {text[ln]:100s} {per_stall[ln].most_common(3)}")
