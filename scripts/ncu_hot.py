"""Top stalled SASS lines of every kernel in an ncu report (source page)."""
import csv, subprocess, io, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 20
only = set(int(x) for x in sys.argv[3].split(",")) if len(sys.argv) > 3 else None
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
heads = [i for i, r in enumerate(rows) if r and r[0] == "Address"] + [len(rows)]
for kid, (a, b) in enumerate(zip(heads[:-1], heads[1:])):
    if only is not None and kid not in only: continue
    hdr = rows[a]; data = [x for x in rows[a + 1:b] if len(x) == len(hdr)]
    isamp = hdr.index('# Samples'); isrc = hdr.index('Source'); iex = hdr.index('Instructions Executed')
    tot = sum(int(x[isamp]) for x in data)
    print("kernel", kid, "total samples", tot, "ninstr", len(data))
    stall = [i for i, c in enumerate(hdr) if c.startswith('stall_') and 'Not Issued' not in c]
    top = sorted(range(len(data)), key=lambda i: -int(data[i][isamp]))[:topn]
    for i in sorted(top):
        x = data[i]; agg = {hdr[c][6:]: int(x[c] or 0) for c in stall}; t = max(1, sum(agg.values()))
        print(f"{i:5d} {100*int(x[isamp])/tot:5.1f}% ex={x[iex]:>8s} {x[isrc][:64]:64s}", {k: round(100 * v / t) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:3]})
