"""Per-region sample shares of an ncu report, regions delimited by marker SASS instructions."""
import csv, subprocess, io, sys
rep=sys.argv[1]; n_items=float(sys.argv[2]) if len(sys.argv)>2 else 27308.0
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h=[i for i,r in enumerate(rows) if r and r[0]=="Address"][0]
hdr=rows[h]; data=[x for x in rows[h+1:] if len(x)==len(hdr)]
isamp=hdr.index('# Samples'); isrc=hdr.index('Source'); iex=hdr.index('Instructions Executed')
tot=sum(int(x[isamp]) for x in data)
print("total samples",tot,"ninstr",len(data))
marks=[]
for i,x in enumerate(data):
    s=x[isrc]
    if any(k in s for k in ("LDTM","STTM","SYNCS.PHASECHK","BAR.SYNC","UBLKCP","SYNCS.ARRIVE","CALL","RET")):
        print(i, x[isamp], x[iex], s[:90]); marks.append(i)
stall=[i for i,c in enumerate(hdr) if c.startswith('stall_') and 'Not Issued' not in c]
marks=[0]+marks+[len(data)]
for a,b in zip(marks[:-1],marks[1:]):
    s=sum(int(x[isamp]) for x in data[a:b]); ex=sum(int(x[iex]) for x in data[a:b])
    if s/tot>0.008:
        agg={hdr[c][6:]:sum(int(x[c] or 0) for x in data[a:b]) for c in stall}; t=max(1,sum(agg.values()))
        print(f"[{a},{b}) {100*s/tot:5.1f}% winstr/item={ex/n_items:8.0f}", {k:round(100*v/t) for k,v in sorted(agg.items(), key=lambda kv:-kv[1])[:4]})
