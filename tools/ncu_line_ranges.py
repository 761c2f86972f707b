"""Aggregate an .ncu-rep's per-source-line stall samples / instructions over line ranges of one file.
usage: python tools/ncu_line_ranges.py rep file.cuh name:lo-hi ..."""
import csv, io, subprocess, sys
rep, fname = sys.argv[1], sys.argv[2]
ranges = [(a.split(":")[0], *map(int, a.split(":")[1].split("-"))) for a in sys.argv[3:]]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur, hdr, rows = "?", None, []
for r in csv.reader(io.StringIO(out)):
    if len(r) >= 2 and r[0] in ("File Path", "File Name"):
        cur = r[1].split("/")[-1]
    elif len(r) > 8 and r[0] == "Line No":
        hdr = r
    elif hdr is not None and len(r) == len(hdr) and r[2] == "-":
        rows.append((cur, int(r[0]), int(r[hdr.index("Instructions Executed")] or 0), int(r[hdr.index("# Samples")] or 0)))
ti = sum(r[2] for r in rows) or 1
ts = sum(r[3] for r in rows) or 1
acc = {n: [0, 0] for n, _, _ in ranges}
other = [0, 0]
for f, ln, i, s in rows:
    hit = False
    if f == fname:
        for n, lo, hi in ranges:
            if lo <= ln <= hi:
                acc[n][0] += i; acc[n][1] += s; hit = True; break
    if not hit:
        other[0] += i; other[1] += s
for n, (i, s) in acc.items():
    print(f"{n:14s} instr {100*i/ti:5.1f}%  samples {100*s/ts:5.1f}%")
print(f"{'other':14s} instr {100*other[0]/ti:5.1f}%  samples {100*other[1]/ts:5.1f}%")
