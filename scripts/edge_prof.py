"""Run the tensor-core edge layer alone at benchmark size (for ncu / timing)."""
import ctypes as C, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from chemeleon_b200 import _lib
from chemeleon_b200.config import SamplerConfig
from chemeleon_b200.engine import DecoderEngine
from chemeleon_b200.topology import BatchTopology
from chemeleon_b200.weights import random_init_state_dict

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
cfg = SamplerConfig(num_layers=1)
eng = DecoderEngine(random_init_state_dict(cfg, seed=4), cfg, precision="tc")
topo = BatchTopology([n] * B, 2, "cuda", exact=False, tensor_core=True)
N = topo.N
x = torch.rand(N, 3, device="cuda")
P = torch.randn(2 * N, 1024, device="cuda").half()
agg = torch.empty(2 * N, 512, device="cuda", dtype=torch.float16)
cg = torch.randn(topo.B, 512, device="cuda")
if os.environ.get("CB2_SINGLE_CTA"):
    eng.model.flags |= _lib.MODEL_EDGE_SINGLE_CTA
def once():
    _lib.check(eng.lib.cb2_edge_layer(C.byref(eng.model), 0, topo.byref(), x.data_ptr(), P.data_ptr(), cg.data_ptr(),
                                      agg.data_ptr(), 512, 1, None, 0, torch.cuda.current_stream().cuda_stream), "edge")
once(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps): once()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
fl = 2 * topo.E * 1310720
print(f"edge layer B={B} n={n}: {ms:.3f} ms/launch  {fl/ms/1e9:.1f} TFLOP/s  tiles={topo.n_tiles}")
if os.environ.get("CB2_TIMELINE"):
    import numpy as np
    buf = (C.c_longlong * 288)()
    pair = not os.environ.get("CB2_SINGLE_CTA")
    fn = eng.lib.cb2_debug_edge2_timeline if pair else eng.lib.cb2_debug_edge_timeline
    fn.argtypes = [C.c_void_p]
    fn(buf)
    a = np.array(buf[:]).reshape(3, 96)
    if pair:
        names = {0: "mma:top", 1: "mma:x_free ok", 2: "mma:G1 issued", 3: "mma:a1_ready ok", 4: "mma:o0_ready ok"}
        for gq in range(4):
            for k, nm in enumerate(["top", "emb12 done", "E1 done", "a1 arrived", "emb0' done", "o_full ok", "E2 done", "refilled"]):
                names[8 + 8 * gq + k] = f"g{gq}:{nm}"
        for itx in range(3):
            base = a[itx, 0]
            ev = sorted((int(a[itx, k] - base), names[k]) for k in names if a[itx, k] != 0)
            print("tile", itx + 1, " | ".join(f"{n}@{t}" for t, n in ev))
            print("   G1 chunk issue times:", [int(a[itx, 48 + k] - base) for k in range(24)])
            print("   G2 stage issue times:", [int(a[itx, 72 + k] - base) for k in range(16)])
            print("   issuer chunk 12: before waits, a_full ok, (w_full ok = chunk time), MMAs issued, commits issued:",
                  [int(a[itx, k] - base) for k in (88, 89, 48 + 12, 90, 91)])
            print("   E1 (thread 0): entry (P loads issued), x_full ok, first 16 columns done, all columns done:",
                  [int(a[itx, 40 + k] - base) for k in range(4)])
            print("   loader stage 16 (for chunk 16): before w_empty wait, w_empty ok, TMA issued:",
                  [int(a[itx, k] - base) for k in (92, 93, 94)])
    else:
        names = {0: "mma:top", 1: "mma:acc_init ok", 2: "mma:G1 issued", 3: "mma:a1_ready ok", 4: "mma:G2(0) issued", 5: "G2(1)", 6: "G2(2)", 7: "G2(3)",
                 8: "w0:emb done", 9: "w0:acc1_full ok", 10: "w0:E1 done", 11: "u0:acc2_full ok", 12: "u0:E2 done", 13: "u0:init done",
                 15: "u1:acc2 ok", 16: "u1:E2 done", 17: "u1:init done", 19: "u2:acc2 ok", 20: "u2:E2 done", 21: "u2:init done",
                 23: "u3:acc2 ok", 24: "u3:E2 done", 25: "u3:init done"}
        for itx in range(3):
            base = a[itx, 0]
            ev = sorted((int(a[itx, k] - base), names[k]) for k in names if a[itx, k] != 0)
            print("item", itx + 1, " | ".join(f"{n}@{t}" for t, n in ev))
