"""Run the tensor-core trunk of a 2-layer decoder at benchmark size: HEAD, FULL and TAIL launch of the
node-chain kernel, timed with CUDA events around the whole forward; with CB2_TIMELINE=1 (library built with
-DCB2_NODE_TIMELINE) print the clock64 timeline of the FULL launch (cluster 0, panels 1..3)."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from chemeleon_b200 import _lib
if os.environ.get("CB2_LIB"):
    _lib.LIB_PATH = os.environ["CB2_LIB"]      # development: a variant build of the library
from chemeleon_b200.config import SamplerConfig
from chemeleon_b200.engine import DecoderEngine
from chemeleon_b200.weights import random_init_state_dict

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
cfg = SamplerConfig(num_layers=2)
eng = DecoderEngine(random_init_state_dict(cfg, seed=4), cfg, precision="tc")
topo = eng.topology([n] * B, 2)
a = torch.randint(1, 90, (topo.N,), device="cuda")
x = torch.rand(topo.N, 3, device="cuda")
l = (torch.eye(3, device="cuda") * 4).reshape(1, 9).repeat(topo.B, 1)
cond = torch.nn.functional.silu(torch.randn(2 * topo.B, 1024, device="cuda"))
for flags in (_lib.MODEL_NODE_UNFUSED, 0):
    eng.model.flags = flags
    eng.forward(topo, a, x, l, cond, want_features=False); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): eng.forward(topo, a, x, l, cond, want_features=False)
    e1.record(); torch.cuda.synchronize()
    print(f"2-layer forward B={B} n={n} flags={flags}: {e0.elapsed_time(e1) / reps:.3f} ms")
if os.environ.get("CB2_TIMELINE"):
    import numpy as np
    buf = (C.c_longlong * 192)()
    fn = eng.lib.cb2_debug_node2_timeline
    fn.argtypes = [C.c_void_p]
    fn(buf)
    t = np.array(buf[:]).reshape(3, 64)
    names = {0: "mma:top", 1: "mma:G1 issued", 2: "mma:x(E1) ok", 3: "mma:G2 issued", 4: "mma:x(E2) ok", 5: "mma:G3 issued",
             6: "mma:x(E3) ok", 7: "mma:u0 issued", 8: "mma:u1 issued", 9: "mma:u2 issued", 10: "mma:u3 issued",
             16: "epi:acc(G1) ok", 17: "epi:E1 done", 18: "epi:acc(G2) ok", 19: "epi:E2 done", 20: "epi:acc(G3) ok",
             21: "epi:pass1 done", 22: "epi:pass2 done", 23: "epi:E3 done", 24: "epi:u0 ok", 25: "epi:u0 done", 26: "epi:u1 ok",
             27: "epi:u1 done", 28: "epi:u2 ok", 29: "epi:u2 done", 30: "epi:u3 ok", 31: "epi:u3 done"}
    for itx in range(3):
        base = t[itx, 0]
        ev = sorted((int(t[itx, k] - base), names[k]) for k in names if t[itx, k] != 0)
        print("panel", itx + 1, " | ".join(f"{nm}@{tt}" for tt, nm in ev))
        print("  G1 issuer: waiting for A", t[itx, 40], "for W", t[itx, 41], "issuing", t[itx, 42])
