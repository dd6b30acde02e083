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
P = torch.randn(2 * N, 1024, device="cuda")
agg = torch.empty(2 * N, 512, device="cuda", dtype=torch.float16)
def once():
    _lib.check(eng.lib.cb2_edge_layer(C.byref(eng.model), 0, topo.byref(), x.data_ptr(), P.data_ptr(), agg.data_ptr(),
                                      512, 1, None, 0, torch.cuda.current_stream().cuda_stream), "edge")
once(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps): once()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
fl = 2 * topo.E * 1310720
print(f"edge layer B={B} n={n}: {ms:.3f} ms/launch  {fl/ms/1e9:.1f} TFLOP/s  tiles={topo.n_tiles}")
