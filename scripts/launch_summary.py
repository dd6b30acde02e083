"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1], errors="ignore")) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
agg = collections.OrderedDict()
seq = []
for r in rows[rows.index(hdr) + 1:]:
    try:
        v = float(r[iv].replace(",", ""))
    except ValueError:
        continue
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}.get(r[iu], 1.0)
    seq.append((r[ik].split("(")[0], v))
for k, v in seq[skip:]:
    a = agg.setdefault(k, [0, 0.0, []]); a[0] += 1; a[1] += v; a[2].append(v)
tot = sum(a[1] for a in agg.values())
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:60]:60s} n={a[0]:4d} total={a[1]/1e3:9.3f} ms  avg={a[1]/a[0]:9.1f} us  share={100*a[1]/tot:5.1f}%  first={[round(x) for x in a[2][:8]]}")
print(f"total {tot/1e3:.3f} ms over {len(seq)-skip} launches")
