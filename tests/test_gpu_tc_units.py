"""GPU unit tests of the tcgen05 building blocks, called through the C-ABI:
cb2_linear_tc against fp64 matmul of the same fp16 operands (bit-level layout check),
cb2_edge_layer tensor-core mode against the exact fp32 mode on the same inputs."""
import ctypes as C

import pytest
import torch

from helpers import rel_err

pytestmark = pytest.mark.gpu


def _stream():
    return torch.cuda.current_stream().cuda_stream


@pytest.mark.parametrize("M,N,K,silu", [(128, 256, 64, 0), (128, 512, 512, 0), (1, 512, 512, 1), (300, 1024, 512, 0),
                                        (1000, 512, 1024, 1)])
def test_linear_tc_matches_fp64(M, N, K, silu):
    from chemeleon_b200 import _lib
    from chemeleon_b200.weights import tile_k_major

    lib = _lib.load()
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g).half()
    W = (torch.randn(N, K, generator=g) / K ** 0.5).half()
    b = torch.randn(N, generator=g)
    ref = A.double() @ W.double().t() + b.double()
    if silu:
        ref = torch.nn.functional.silu(ref)
    Ad, Wt, bd = A.cuda(), tile_k_major(W.float()).cuda(), b.cuda()
    Cd = torch.full((M, N), float("nan"), device="cuda")
    _lib.check(lib.cb2_linear_tc(Ad.data_ptr(), K, Wt.data_ptr(), N, bd.data_ptr(), Cd.data_ptr(), N, M, K, silu,
                                 _stream()), "cb2_linear_tc")
    torch.cuda.synchronize()
    assert torch.isfinite(Cd).all()
    assert rel_err(Cd.cpu(), ref) < 2e-5


@pytest.mark.parametrize("natoms,V", [([20] * 7, 1), ([4, 7, 5, 1, 40, 33], 2), ([6, 6, 6], 2), ([20] * 300, 2)])
def test_edge_layer_tc_vs_fp32(natoms, V):
    from chemeleon_b200 import _lib
    from chemeleon_b200.config import SamplerConfig
    from chemeleon_b200.engine import DecoderEngine
    from chemeleon_b200.topology import BatchTopology
    from chemeleon_b200.weights import random_init_state_dict

    cfg = SamplerConfig(num_layers=1)
    sd = random_init_state_dict(cfg, seed=4)
    eng = DecoderEngine(sd, cfg, precision="tc")
    lib = eng.lib
    topo = BatchTopology(natoms, V, "cuda", exact=True, tensor_core=True)
    N = topo.N
    g = torch.Generator().manual_seed(1)
    x = (torch.rand(N, 3, generator=g) * 2 - 0.5).cuda()
    P = torch.randn(V * N, 1024, generator=g).cuda()
    ws = torch.empty(int(lib.cb2_workspace_bytes(topo.byref(), 0)), dtype=torch.uint8, device="cuda")
    agg32 = torch.zeros(V * N, 512, device="cuda")
    _lib.check(lib.cb2_edge_layer(C.byref(eng.model), 0, topo.byref(), x.data_ptr(), P.data_ptr(), agg32.data_ptr(),
                                  512, 0, ws.data_ptr(), ws.numel(), _stream()), "edge fp32")
    agg16 = torch.full((V * N, 512), float("nan"), device="cuda", dtype=torch.float16)
    P16 = P.half()                      # tensor-core mode gathers the hoisted terms as fp16
    _lib.check(lib.cb2_edge_layer(C.byref(eng.model), 0, topo.byref(), x.data_ptr(), P16.data_ptr(), agg16.data_ptr(),
                                  512, 1, None, 0, _stream()), "edge tc")
    torch.cuda.synchronize()
    assert torch.isfinite(agg16).all()
    err = rel_err(agg16.float().cpu(), agg32.cpu())
    print(f"edge layer tc vs fp32: rel err {err:.2e}")
    assert err < 2e-3


def test_edge_layer_pair_kernel_matches_default():
    """The CTA-pair variant of the edge kernel (both CFG variants per tile, GEMM1 shared, a1 halves
    exchanged through distributed shared memory; opt-in with CB2_EDGE_PAIR=1) gives the same
    aggregates as the default kernel (checksums to 1e-5).  Runs in a subprocess: the switch is read once per process."""
    import os
    import subprocess
    import sys

    code = r"""
import ctypes as C, sys, torch
sys.path.insert(0, %r)
from chemeleon_b200 import _lib
from chemeleon_b200.config import SamplerConfig
from chemeleon_b200.engine import DecoderEngine
from chemeleon_b200.topology import BatchTopology
from chemeleon_b200.weights import random_init_state_dict
cfg = SamplerConfig(num_layers=1)
eng = DecoderEngine(random_init_state_dict(cfg, seed=4), cfg, precision="tc")
for natoms in ([4, 7, 5, 1, 40, 33], [20] * 300):
    topo = BatchTopology(natoms, 2, "cuda", exact=False, tensor_core=True)
    g = torch.Generator().manual_seed(1)
    x = (torch.rand(topo.N, 3, generator=g) * 2 - 0.5).cuda()
    P = torch.randn(2 * topo.N, 1024, generator=g).cuda().half()
    agg = torch.full((2 * topo.N, 512), float("nan"), device="cuda", dtype=torch.float16)
    _lib.check(eng.lib.cb2_edge_layer(C.byref(eng.model), 0, topo.byref(), x.data_ptr(), P.data_ptr(), agg.data_ptr(),
                                      512, 1, None, 0, torch.cuda.current_stream().cuda_stream), "edge")
    torch.cuda.synchronize()
    assert torch.isfinite(agg).all()
    print("CHECKSUM", float(agg.double().sum()), float(agg.double().abs().sum()))
""" % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for pair in ("0", "1"):
        env = dict(os.environ, CB2_EDGE_PAIR=pair)
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append([l for l in r.stdout.splitlines() if l.startswith("CHECKSUM")])
    # not bit-identical: the K chunks are accumulated in a different order (fp32 rounding)
    assert len(outs[0]) == 2 and len(outs[1]) == 2, outs
    for a, b in zip(outs[0], outs[1]):
        va, vb = [float(v) for v in a.split()[1:]], [float(v) for v in b.split()[1:]]
        assert abs(va[0] - vb[0]) <= 1e-5 * va[1] and abs(va[1] - vb[1]) <= 1e-5 * va[1], outs
