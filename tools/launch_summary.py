"""Summarise an ncu launch list (gpu__time_duration.sum CSV): time per kernel, and per slot of the phased engine."""
import csv, sys, collections, re
path = sys.argv[1]
rows = []
hdr = None
for r in csv.reader(open(path)):
    if len(r) > 5 and r[0] == "ID":
        hdr = r; continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        try:
            t = float(d["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        unit = d.get("Metric Unit", "ns")
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(unit, 1e-3)
        name = re.sub(r"\(.*", "", d["Kernel Name"]).replace("void ", "")
        name = re.sub(r"altro_b200::", "", name)
        rows.append((int(d["ID"]), name, d["Grid Size"], t * scale))
tot = sum(r[3] for r in rows)
agg = collections.OrderedDict()
for _, name, grid, t in rows:
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += t
print(f"total {tot/1e3:.2f} ms over {len(rows)} launches")
for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {t/1e3:9.3f} ms {100*t/tot:5.1f}%  x{c:5d}  avg {t/c:8.1f} us  {name}")
if len(sys.argv) > 2:
    # per-slot table for the phased engine: a slot starts at each k_update_expansions launch
    slots = []
    for _, name, grid, t in rows:
        if name.startswith("k_update_expansions"):
            slots.append(collections.OrderedDict())
        if slots:
            key = name.split("<")[0]
            slots[-1][key] = slots[-1].get(key, 0.0) + t
            slots[-1].setdefault("_grid_" + key, grid)
    step = int(sys.argv[2])
    for i in range(0, len(slots), step):
        s = slots[i]
        print(i, " ".join(f"{k}={v:.0f}" for k, v in s.items() if not k.startswith("_")), "tiles:", s.get("_grid_k_ls_wide"))
