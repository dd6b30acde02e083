"""Per-kernel count of the Blackwell-specific SASS instructions in libchemeleon_b200.so
(`cuobjdump -sass`): UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG = tensor-map TMA,
UBLKCP = bulk copy, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, .2CTA = cta_group::2.
    python scripts/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections, os, re, subprocess, sys
so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "chemeleon_b200", "libchemeleon_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
pats = collections.OrderedDict([("UTCHMMA", r"\bUTCHMMA"), ("UTCHMMA.2CTA", r"\bUTCHMMA\.2CTA"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"),
                                ("UTMALDG", r"\bUTMALDG"), ("UBLKCP", r"\bUBLKCP"), ("UTCBAR", r"\bUTCBAR"),
                                ("UTCBAR.2CTA.MULTICAST", r"\bUTCBAR[.\w]*2CTA"), ("SYNCS", r"\bSYNCS"), ("STAS (st.async)", r"\bSTAS"),
                                ("MUFU.TANH", r"\bMUFU\.TANH"), ("HMMA (legacy)", r"\bHMMA")])
cur, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur:
        for k, p in pats.items():
            if re.search(p, line):
                counts[cur][k] += 1
print("# cuobjdump -sass chemeleon_b200/libchemeleon_b200.so | per-kernel instruction counts (static)")
print("kernel".ljust(44) + "".join(k.rjust(max(9, len(k) + 2)) for k in pats))
for fn, c in counts.items():
    if not any(c.values()):
        continue
    name = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip().split("(")[0]
    print(name[:43].ljust(44) + "".join(str(c[k]).rjust(max(9, len(k) + 2)) for k in pats))
