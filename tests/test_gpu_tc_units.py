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
    ws = torch.empty(int(lib.cb2_linear_tc_workspace_bytes(M, K)), dtype=torch.uint8, device="cuda")
    assert lib.cb2_linear_tc(Ad.data_ptr(), K, Wt.data_ptr(), N, bd.data_ptr(), Cd.data_ptr(), N, M, K, silu,
                             None, 0, _stream()) == -5      # CB2_ERR_WORKSPACE: the library never allocates
    _lib.check(lib.cb2_linear_tc(Ad.data_ptr(), K, Wt.data_ptr(), N, bd.data_ptr(), Cd.data_ptr(), N, M, K, silu,
                                 ws.data_ptr(), ws.numel(), _stream()), "cb2_linear_tc")
    torch.cuda.synchronize()
    assert torch.isfinite(Cd).all()
    assert rel_err(Cd.cpu(), ref) < 2e-5


@pytest.mark.parametrize("cg_scale", [0.0, 10.0])
@pytest.mark.parametrize("natoms,V", [([20] * 7, 1), ([4, 7, 5, 1, 40, 33], 2), ([6, 6, 6], 2), ([20] * 300, 2),
                                      ([40] * 50 + [33] * 3, 2), ([1, 2, 3, 64, 41], 2)])
def test_edge_layer_tc_vs_fp32(natoms, V, cg_scale):
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
    # per-crystal lattice term at the size Angstrom-scale cells give it (|L L^T| ~ 1e2..1e3)
    cg = (torch.randn(topo.B, 512, generator=g) * cg_scale).cuda() if cg_scale else None
    cgp = cg.data_ptr() if cg is not None else None
    ws = torch.empty(int(lib.cb2_workspace_bytes(C.byref(eng.model), topo.byref(), 0)), dtype=torch.uint8, device="cuda")
    agg32 = torch.zeros(V * N, 512, device="cuda")
    _lib.check(lib.cb2_edge_layer(C.byref(eng.model), 0, topo.byref(), x.data_ptr(), P.data_ptr(), cgp,
                                  agg32.data_ptr(), 512, 0, ws.data_ptr(), ws.numel(), _stream()), "edge fp32")
    agg16 = torch.full((V * N, 512), float("nan"), device="cuda", dtype=torch.float16)
    P16 = P.half()                      # tensor-core mode gathers the hoisted terms as fp16
    _lib.check(lib.cb2_edge_layer(C.byref(eng.model), 0, topo.byref(), x.data_ptr(), P16.data_ptr(), cgp,
                                  agg16.data_ptr(), 512, 1, None, 0, _stream()), "edge tc")
    torch.cuda.synchronize()
    assert torch.isfinite(agg16).all()
    err = rel_err(agg16.float().cpu(), agg32.cpu())
    print(f"edge layer tc vs fp32: rel err {err:.2e}")
    assert err < 2e-3


@pytest.mark.parametrize("natoms", [[20] * 300, [4, 7, 5, 1, 40, 33], [6, 6, 6], [40] * 37 + [13] * 5 + [64, 3]])
def test_edge_pair_kernel_matches_single_cta_kernel(natoms):
    """V = 2: the CTA-pair kernel (cta_group::2, sinusoid GEMM shared by the CFG variants, a1 halves
    exchanged through distributed shared memory, tensor-map TMA) gives the aggregates of the one-CTA
    kernel run per variant (selected with cb2_model.flags); only the fp32 accumulation order differs."""
    from chemeleon_b200 import _lib
    from chemeleon_b200.config import SamplerConfig
    from chemeleon_b200.engine import DecoderEngine
    from chemeleon_b200.topology import BatchTopology
    from chemeleon_b200.weights import random_init_state_dict

    cfg = SamplerConfig(num_layers=1)
    eng = DecoderEngine(random_init_state_dict(cfg, seed=4), cfg, precision="tc")
    topo = BatchTopology(natoms, 2, "cuda", exact=False, tensor_core=True)
    g = torch.Generator().manual_seed(1)
    x = (torch.rand(topo.N, 3, generator=g) * 2 - 0.5).cuda()
    P = torch.randn(2 * topo.N, 1024, generator=g).cuda().half()
    cg = (torch.randn(topo.B, 512, generator=g) * 5).cuda()
    outs = []
    for flags in (_lib.MODEL_EDGE_SINGLE_CTA, 0):
        eng.model.flags = flags
        agg = torch.full((2 * topo.N, 512), float("nan"), device="cuda", dtype=torch.float16)
        for _ in range(2):          # twice: the second launch starts from whatever the first left in TMEM / smem
            _lib.check(eng.lib.cb2_edge_layer(C.byref(eng.model), 0, topo.byref(), x.data_ptr(), P.data_ptr(),
                                              cg.data_ptr(), agg.data_ptr(), 512, 1, None, 0, _stream()), "edge")
        torch.cuda.synchronize()
        assert torch.isfinite(agg).all()
        outs.append(agg.float().cpu())
    err = rel_err(outs[1], outs[0])
    print(f"pair vs single-CTA edge kernel: rel err {err:.2e}")
    assert err < 1e-3   # both round a1 / agg to fp16; a last-bit flip of an fp16 output is 5e-4 of its value


def test_edge_pair_kernel_is_repeatable():
    """compute-sanitizer is closed on the GPU pool, so races are hunted the blunt way: 12 launches of the
    CTA-pair kernel on the same input (tiles of several segment lengths, > 1 tile per cluster) must agree
    ELEMENT BY ELEMENT to the last fp16 bits -- a race on the aliased a1 / weight-stage region, the
    embedding ring, the st.async exchange or the TMEM hand-over corrupts whole 16-byte pieces or tiles and
    shows up as O(1) differences.  (Bit-identity is not required: four threads issue alternate K chunks, so
    the fp32 accumulation order of a tile varies from launch to launch and an fp16 output may round the other way.)"""
    from chemeleon_b200 import _lib
    from chemeleon_b200.config import SamplerConfig
    from chemeleon_b200.engine import DecoderEngine
    from chemeleon_b200.topology import BatchTopology
    from chemeleon_b200.weights import random_init_state_dict

    cfg = SamplerConfig(num_layers=1)
    eng = DecoderEngine(random_init_state_dict(cfg, seed=4), cfg, precision="tc")
    natoms = [20] * 900 + [6] * 400 + [33] * 30 + [40] * 60
    topo = BatchTopology(natoms, 2, "cuda", exact=False, tensor_core=True)
    g = torch.Generator().manual_seed(2)
    x = torch.rand(topo.N, 3, generator=g).cuda()
    P = torch.randn(2 * topo.N, 1024, generator=g).cuda().half()
    cg = (torch.randn(topo.B, 512, generator=g) * 3).cuda()
    ref = None
    n_diff = 0
    for _ in range(12):
        agg = torch.full((2 * topo.N, 512), float("nan"), device="cuda", dtype=torch.float16)
        _lib.check(eng.lib.cb2_edge_layer(C.byref(eng.model), 0, topo.byref(), x.data_ptr(), P.data_ptr(), cg.data_ptr(),
                                          agg.data_ptr(), 512, 1, None, 0, _stream()), "edge")
        torch.cuda.synchronize()
        assert torch.isfinite(agg).all()
        if ref is None:
            ref = agg.clone()
        else:
            n_diff += int((agg.view(torch.int16) != ref.view(torch.int16)).sum())
            a, r = agg.float(), ref.float()
            bad = (a - r).abs() > 2.5e-3 * r.abs() + 2e-4         # 2 fp16 ulps (+ an absolute floor near zero)
            assert int(bad.sum()) == 0, f"{int(bad.sum())} outputs differ by more than 2 fp16 ulps"
    print(f"pair kernel repeatability: {n_diff} of {11 * ref.numel()} fp16 outputs differ in their last bits, none by more than 2 ulps")


@pytest.mark.parametrize("natoms,layers,coords_only", [([6, 6, 6], 2, False), ([20] * 7, 1, False),
                                                      ([4, 7, 5, 1, 40, 33], 3, True), ([20] * 300 + [33] * 5, 6, False),
                                                      ([40] * 150 + [13] * 5 + [64, 3], 2, False)])
def test_node_chain_kernel_matches_unfused_kernels(natoms, layers, coords_only):
    """The CTA-pair node-chain kernel (k_tc_node2: node MLP -> FiLM block -> LayerNorm -> hoist GEMM per layer
    boundary, activations resident in shared memory, cta_group::2) against round 1's four kernels per layer
    (k_tc_film + 3 x k_tc_linear, selected with cb2_model.flags): same fp16 operands, same K order, same
    epilogue arithmetic -- decoder outputs agree to fp32 accumulation noise.  Covers 1 panel (phantom peer
    panel), odd panel counts, partial last panels, 1 / 2 / 3 / 6 layers (HEAD + TAIL only, with FULL links)
    and several work items per cluster; each forward is run twice (the second starts from the first's TMEM / smem)."""
    from chemeleon_b200 import _lib
    from chemeleon_b200.config import SamplerConfig
    from chemeleon_b200.engine import DecoderEngine
    from chemeleon_b200.weights import random_init_state_dict

    cfg = SamplerConfig(num_layers=layers)
    eng = DecoderEngine(random_init_state_dict(cfg, seed=11), cfg, precision="tc")
    topo = eng.topology(natoms, 2)
    g = torch.Generator().manual_seed(5)
    a = torch.randint(1, 90, (topo.N,), generator=g).cuda()
    x = torch.rand(topo.N, 3, generator=g).cuda()
    l = (torch.randn(topo.B, 3, 3, generator=g) * 0.3 + 4 * torch.eye(3)).reshape(topo.B, 9).cuda()
    cond = torch.nn.functional.silu(torch.randn(2 * topo.B, 1024, generator=g)).cuda()
    outs = []
    for flags in (_lib.MODEL_NODE_UNFUSED, 0):
        eng.model.flags = flags
        for _ in range(2):
            head, lat, feat = eng.forward(topo, a, x, l, cond, coords_only=coords_only)
        torch.cuda.synchronize()
        assert torch.isfinite(head[:, 104:107]).all() and torch.isfinite(feat).all()
        outs.append((head.cpu(), lat.cpu(), feat.cpu()))
    e_feat = rel_err(outs[1][2], outs[0][2])
    e_x = rel_err(outs[1][0][:, 104:107], outs[0][0][:, 104:107])
    print(f"node chain vs unfused: features {e_feat:.2e}, coords {e_x:.2e}")
    assert e_feat < 2e-4 and e_x < 2e-4
    if not coords_only:
        assert rel_err(outs[1][0][:, :104], outs[0][0][:, :104]) < 2e-4
        assert rel_err(outs[1][1], outs[0][1]) < 2e-4
