"""Small shapes through every tcgen05 kernel, for `compute-sanitizer --tool memcheck|racecheck|synccheck`:
the CTA-pair edge kernel (k_tc_edge2), the one-CTA edge kernel (k_tc_edge), and one full sampler timestep
(k_tc_film, k_tc_linear, heads, update kernels) on a ragged batch."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from chemeleon_b200 import _lib
from chemeleon_b200.config import SamplerConfig
from chemeleon_b200.engine import DecoderEngine
from chemeleon_b200.sampler import ChemeleonB200
from chemeleon_b200.topology import BatchTopology
from chemeleon_b200.weights import random_init_state_dict

cfg = SamplerConfig(num_layers=1)
eng = DecoderEngine(random_init_state_dict(cfg, seed=4), cfg, precision="tc")
natoms = [3, 5, 20, 20, 6, 40, 1]
topo = BatchTopology(natoms, 2, "cuda", exact=False, tensor_core=True)
g = torch.Generator().manual_seed(1)
x = torch.rand(topo.N, 3, generator=g).cuda()
P = torch.randn(2 * topo.N, 1024, generator=g).cuda().half()
cg = torch.randn(topo.B, 512, generator=g).cuda()
for flags in (0, _lib.MODEL_EDGE_SINGLE_CTA):
    eng.model.flags = flags
    agg = torch.zeros(2 * topo.N, 512, device="cuda", dtype=torch.float16)
    _lib.check(eng.lib.cb2_edge_layer(C.byref(eng.model), 0, topo.byref(), x.data_ptr(), P.data_ptr(), cg.data_ptr(),
                                      agg.data_ptr(), 512, 1, None, 0, torch.cuda.current_stream().cuda_stream), "edge")
    torch.cuda.synchronize()
    print("edge layer flags", flags, "checksum", float(agg.float().abs().sum()))
cfg2 = SamplerConfig(num_layers=2, timesteps=3)
model = ChemeleonB200(random_init_state_dict(cfg2, seed=0, head_scale=0.01, lattice_identity=True), cfg2, precision="tc",
                      use_cuda_graph=False)
text, null = torch.randn(len(natoms), 512, generator=g), torch.randn(1, 512, generator=g)
a, xx, l = model.sample_states(natoms, text, null, seed=3)
torch.cuda.synchronize()
print("sampler steps ok", bool(torch.isfinite(xx).all()), int(model.last_flags.sum()))
