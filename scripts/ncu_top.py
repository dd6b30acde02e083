"""Summarise an ncu report: key metrics + top stalled SASS lines (with source lines)."""
import csv, subprocess, sys, io
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = r[0], r[1], r[2]
d = dict(zip(hdr, zip(units, vals)))
for k in ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
          'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'lts__t_sectors.sum',
          'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
          'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
          'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
          'sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
          'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
          'l1tex__t_sector_hit_rate.pct', 'sm__cycles_elapsed.avg']:
    if k in d: print(f"{k:75s} {d[k][0]:10s} {d[k][1]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = None
for i, row in enumerate(rows):
    if row and row[0] == "Address": h = i; break
if h is None:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    for i, row in enumerate(rows):
        if row and row[0] == "Address": h = i; break
hdr = rows[h]; data = [x for x in rows[h+1:] if len(x) == len(hdr)]
isrc = hdr.index('Source'); isamp = hdr.index('# Samples'); iex = hdr.index('Instructions Executed')
tot = sum(int(x[isamp]) for x in data)
print("total samples", tot, "instructions", len(data))
stall = [i for i, c in enumerate(hdr) if c.startswith('stall_') and 'Not Issued' not in c]
top = sorted(range(len(data)), key=lambda i: -int(data[i][isamp]))[:topn]
for i in sorted(top):
    x = data[i]
    st = sorted([(int(x[c]), hdr[c][6:]) for c in stall if x[c] not in ('', '0')], reverse=True)[:2]
    print(f"{i:5d} {100*int(x[isamp])/tot:5.1f}% ex={x[iex]:>10s} {x[isrc][:64]:64s} {st}")
