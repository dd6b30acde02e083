"""Turn the ncu artefacts of a round into the committed text summaries under profiles/."""
import collections, csv, io, json, subprocess, sys

def launches(path, out):
    lines = [l for l in open(path) if not l.startswith('==')]
    agg = collections.OrderedDict(); tot = 0.0
    for row in csv.DictReader(lines):
        name = row['Kernel Name'].split('(')[0]
        v = float(row['Metric Value'].replace(',', ''))
        u = row['Metric Unit']
        v = v / 1e6 if u == 'ns' else v / 1e3 if u == 'us' else v * 1e3 if u == 's' else v
        d = agg.setdefault(name, [0, 0.0]); d[0] += 1; d[1] += v; tot += v
    with open(out, 'w') as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none, `python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph --no-roofline --no-e2e`\n")
        f.write("# (warm-up timestep + one timed timestep; per-launch times are cold-cache and serialised: compare SHARES)\n")
        f.write(f"{'kernel':58s} {'launches':>8s} {'total ms':>10s} {'avg ms':>9s} {'share':>7s}\n")
        for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k[:58]:58s} {c:8d} {t:10.3f} {t/c:9.4f} {100*t/tot:6.1f}%\n")
        f.write(f"{'total':58s} {'':8s} {tot:10.3f}\n")

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum', 'lts__t_sectors.sum', 'lts__t_sector_hit_rate.pct',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__block_size',
        'launch__grid_size', 'launch__shared_mem_per_block_dynamic', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active',
        'l1tex__t_sector_hit_rate.pct', 'sm__cycles_elapsed.avg', 'sm__throughput.avg.pct_of_peak_sustained_elapsed']

def full(rep, out, title, index=0):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = r[0], r[1], r[2 + index]
    d = dict(zip(hdr, zip(units, vals)))
    with open(out, 'w') as f:
        f.write(f"# {title}\n# ncu --set full --clock-control none --import-source on (launch {index} of the capture)\n")
        for k in ['Kernel Name'] + KEYS:
            if k in d:
                f.write(f"{k:72s} {d[k][0]:16s} {d[k][1]}\n")
        stalls = [(float(d[h][1] or 0), h) for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio')]
        f.write("# warp stall reasons (warps per issue-active cycle)\n")
        for v, h in sorted(stalls, reverse=True)[:8]:
            f.write(f"{h:72s} {v:.3f}\n")

if __name__ == "__main__":
    # round 2, final state: `python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph --no-roofline --no-e2e`
    launches("gpurun_out/r2f_launches.csv", "profiles/r2_launches.txt")
    rep = "gpurun_out/r2f_full.ncu-rep"    # -k regex:k_tc_(edge2|node2) -s 10 -c 5: FULL, edge, TAIL, HEAD, edge
    full(rep, "profiles/r2_edge_full.txt", "k_tc_edge2 (CTA pair, cta_group::2), C3 (B=4096, n=20, cond+null), one CSPLayer", 1)
    full(rep, "profiles/r2_node2_full.txt", "k_tc_node2 mode FULL (node MLP of layer l, FiLM block + LayerNorm + hoist GEMM of layer l+1), C3: 163 840 rows", 0)
    full(rep, "profiles/r2_node2_tail.txt", "k_tc_node2 mode TAIL (node MLP of the last layer), C3: 163 840 rows", 2)
    full(rep, "profiles/r2_node2_head.txt", "k_tc_node2 mode HEAD (FiLM block + LayerNorm + hoist GEMM of the first layer), C3: 163 840 rows", 3)
