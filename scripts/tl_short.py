"""Condense node_prof timeline lines (stdin) into phase durations."""
import re, sys
for line in sys.stdin:
    if not line.startswith("panel"):
        if "forward" in line or "variant" in line or "issuer" in line: print(line.strip())
        continue
    ev = {m.group(1): int(m.group(2)) for m in re.finditer(r"((?:mma|epi):[^@|]+)@(\d+)", line)}
    g = lambda k: ev.get(k, 0)
    print(line.split()[1], "G1", g("mma:G1 issued"), "E1", g("epi:E1 done") - g("epi:acc(G1) ok"), "G2", g("mma:G2 issued") - g("mma:x(E1) ok"),
          "E2", g("epi:E2 done") - g("epi:acc(G2) ok"), "G3", g("mma:G3 issued") - g("mma:x(E2) ok"),
          "p1", g("epi:pass1 done") - g("epi:acc(G3) ok"), "p2", g("epi:pass2 done") - g("epi:pass1 done"), "p3", g("epi:E3 done") - g("epi:pass2 done"),
          "G4", g("mma:u3 issued") - g("mma:x(E3) ok"), "total", g("epi:u3 done"))
