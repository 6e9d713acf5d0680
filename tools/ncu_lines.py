"""Per-source-line hot spots from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
sections, cur, hdr = [], None, None
for i, r in enumerate(rows):
    if not r:
        continue
    if r[0] == "Line No":
        label = " | ".join(" ".join(x[:60] for x in rows[j][:2]) for j in range(max(0, i - 2), i) if rows[j])
        cur = {"label": label, "lines": []}; sections.append(cur); hdr = r
        continue
    if cur is not None and r[0] != "" and len(r) > 8:
        try:
            cur["lines"].append((int(r[0]), r[1][:100], int(r[hdr.index("Instructions Executed")]), int(r[hdr.index("# Samples")])))
        except ValueError:
            pass
tot_i = sum(l[2] for s in sections for l in s["lines"]); tot_s = sum(l[3] for s in sections for l in s["lines"])
for s in sections:
    si = sum(l[2] for l in s["lines"]); ss = sum(l[3] for l in s["lines"])
    if ss < 0.01 * tot_s:
        continue
    print(f"===== {s['label']}  inst {si} samples {ss}")
    for l in sorted(s["lines"], key=lambda l: -l[3])[:top]:
        print(f"  L{l[0]:4d} inst {100*l[2]/max(si,1):5.1f}% samp {100*l[3]/max(ss,1):5.1f}%  {l[1]}")
