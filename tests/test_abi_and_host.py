"""CPU: the C-ABI library loads and exports every symbol include/chemeleon_b200.h
declares (no compute calls without a GPU); host-side logic (topology, schedules,
output boundary, partitioning)."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from chemeleon_b200 import _lib

    return _lib.load()


def test_header_symbols_all_exported(lib):
    from chemeleon_b200 import _lib

    hdr = open(os.path.join(ROOT, "include", "chemeleon_b200.h")).read()
    decl = set(re.findall(r"\b(cb2_[a-z0-9_]+)\s*\(", hdr))
    assert decl, "no declarations parsed"
    for name in decl:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert decl == set(_lib.EXPORTS), (decl ^ set(_lib.EXPORTS))
    assert lib.cb2_abi_version() == _lib.ABI_VERSION


def test_struct_layouts_match_header():
    """ctypes mirrors must have the C layout (checked against a tiny gcc-compiled probe)."""
    import subprocess
    import tempfile

    from chemeleon_b200 import _lib

    src = r'''
#include <stdio.h>
#include <stddef.h>
#include "chemeleon_b200.h"
int main(void){
  printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(cb2_layer_weights), sizeof(cb2_model), sizeof(cb2_batch),
         sizeof(cb2_forward_io), sizeof(cb2_state), sizeof(cb2_step_args), offsetof(cb2_model, final_g));
  printf("%zu %zu %zu %zu %zu\n", offsetof(cb2_batch, host_chunk_node_lo), offsetof(cb2_batch, tile_row_i),
         offsetof(cb2_step_args, rand_a), offsetof(cb2_step_args, graph_gid), offsetof(cb2_model, flags));
  return 0; }
'''
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "probe.c")
        open(p, "w").write(src)
        exe = os.path.join(d, "probe")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), p, "-o", exe])
        out = subprocess.check_output([exe]).decode().split()
    got = [int(v) for v in out]
    want = [C.sizeof(_lib.LayerWeights), C.sizeof(_lib.Model), C.sizeof(_lib.Batch), C.sizeof(_lib.ForwardIO),
            C.sizeof(_lib.State), C.sizeof(_lib.StepArgs), _lib.Model.final_g.offset,
            _lib.Batch.host_chunk_node_lo.offset, _lib.Batch.tile_row_i.offset, _lib.StepArgs.rand_a.offset,
            _lib.StepArgs.graph_gid.offset, _lib.Model.flags.offset]
    assert got == want


def test_bad_arguments_fail_loudly(lib):
    from chemeleon_b200 import _lib

    m = _lib.Model()  # abi_version 0
    b = _lib.Batch()
    io = _lib.ForwardIO()
    assert lib.cb2_decoder_forward(C.byref(m), C.byref(b), C.byref(io), None, 0, None) != 0
    assert b"abi_version" in lib.cb2_last_error()
    m.abi_version, m.hidden, m.n_atom_types, m.n_freqs, m.n_layers = _lib.ABI_VERSION, 256, 104, 128, 6
    assert lib.cb2_decoder_forward(C.byref(m), C.byref(b), C.byref(io), None, 0, None) == -2  # unsupported shape
    assert lib.cb2_workspace_bytes(None, None, 0) == 0


def test_no_gpu_means_no_fallback():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from chemeleon_b200 import _lib
    from chemeleon_b200.engine import DecoderEngine

    with pytest.raises(_lib.Cb2Error):
        DecoderEngine({}, device="cuda")


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "chemeleon_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            txt = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in txt.replace("# oracle", ""), f"{fn} mentions the oracle"


def _brute_edges(natoms):
    ei, ej, off = [], [], 0
    for n in natoms:
        for i in range(n):
            for j in range(n):
                ei.append(off + i)
                ej.append(off + j)
        off += n
    return np.array(ei), np.array(ej)


@pytest.mark.parametrize("natoms", [[4, 7, 5, 1], [20] * 9, [40, 3, 33, 40, 1, 128], [6, 6, 6]])
def test_topology_edges_and_tiles(natoms):
    from chemeleon_b200.topology import BatchTopology, build_tiles

    t = BatchTopology(natoms, 2, "cpu", chunk_edges=64)
    ei, ej = _brute_edges(natoms)
    assert np.array_equal(t.edge_i.numpy(), ei) and np.array_equal(t.edge_j.numpy(), ej)
    assert t.E == len(ei) and t.N == sum(natoms)
    # chunks are whole segments and cover everything
    nlo, elo = t.host_chunk_node_lo, t.host_chunk_edge_lo
    assert nlo[0] == 0 and nlo[-1] == t.N and elo[-1] == t.E
    assert (np.diff(elo) <= max(64, max(natoms) ** 1 * max(natoms))).all()
    assert np.array_equal(elo, t.node_eoff.numpy()[nlo])
    # tiles: every edge exactly once, segments whole and of the tile's length
    ri, rj, sn = t.tile_row_i.numpy(), t.tile_row_j.numpy(), t.tile_seg_n.numpy()
    ri, rj = ri.reshape(-1, 128), rj.reshape(-1, 128)
    seen = set()
    for k in range(ri.shape[0]):
        n = sn[k]
        S = 128 // n
        assert (ri[k, S * n:] == -1).all()
        for s in range(S):
            seg_i = ri[k, s * n:(s + 1) * n]
            if seg_i[0] < 0:
                assert (seg_i == -1).all()
                continue
            assert (seg_i == seg_i[0]).all()
            base = t.node_base.numpy()[seg_i[0]]
            assert t.node_n.numpy()[seg_i[0]] == n
            assert np.array_equal(rj[k, s * n:(s + 1) * n], base + np.arange(n))
            for j in rj[k, s * n:(s + 1) * n]:
                seen.add((int(seg_i[0]), int(j)))
    assert seen == set(zip(ei.tolist(), ej.tolist()))


def test_state_to_atoms_semantics():
    from chemeleon_b200.atoms import state_to_atoms, HAVE_ASE

    a = np.array([8, 3, 25, 8, 200, 1], dtype=np.int64)   # O Li Mn O (>103 -> X) H
    x = np.arange(18, dtype=np.float32).reshape(6, 3) / 20
    l = np.stack([np.eye(3) * 4, np.eye(3) * 5]).astype(np.float32)
    out = state_to_atoms(a, x, l, [4, 2])
    assert len(out) == 2 and len(out[0]) == 4 and len(out[1]) == 2
    z0 = list(out[0].get_atomic_numbers()) if HAVE_ASE else list(out[0].numbers)
    assert z0 == [3, 25, 8, 8]            # sorted by symbol: Li, Mn, O, O (stable)
    z1 = list(out[1].get_atomic_numbers()) if HAVE_ASE else list(out[1].numbers)
    assert z1 == [1, 0]                   # 'H' < 'X'
    sp = out[0].get_scaled_positions()
    assert np.allclose(sp[0], x[1]) and np.allclose(sp[2], x[0]) and np.allclose(sp[3], x[3])


def test_partition_is_balanced_and_complete():
    from chemeleon_b200.dist import partition_samples, sample_cost

    rng = np.random.RandomState(3)
    natoms = rng.randint(4, 41, size=1000).tolist()
    for w in (1, 2, 4, 8):
        parts = partition_samples(natoms, w)
        assert sorted(i for p in parts for i in p) == list(range(1000))
        loads = [sum(sample_cost(natoms[i]) for i in p) for p in parts]
        assert max(loads) / (sum(loads) / w) < 1.01
        for p in parts:
            ns = [natoms[i] for i in p]
            assert ns == sorted(ns)


def test_config_validation():
    from chemeleon_b200.config import SamplerConfig

    SamplerConfig().validate()
    with pytest.raises(ValueError):
        SamplerConfig(hidden_dim=256).validate()
    with pytest.raises(ValueError):
        SamplerConfig(edge_style="knn").validate()
    c = SamplerConfig.from_hparams({"hidden_dim": 512, "lr": 1e-3, "num_layers": 4})
    assert c.num_layers == 4


def test_weight_packing_layouts():
    from chemeleon_b200.config import SamplerConfig
    from chemeleon_b200.weights import fd_column_order, pack_weights, random_init_state_dict, tile_k_major

    cfg = SamplerConfig(num_layers=1)
    sd = random_init_state_dict(cfg, seed=1)
    pw = pack_weights(sd, cfg, device="cpu", tensor_core=True)
    w1 = sd["decoder.csp_layer_0.edge_mlp.0.weight"]
    L = pw.layers[0]
    assert torch.equal(L.w_hij[:512], w1[:, :512]) and torch.equal(L.w_hij[512:], w1[:, 512:1024])
    assert torch.equal(L.w_ip, w1[:, 1024:1033]) and torch.equal(L.w_fd, w1[:, 1033:])
    # tiled image: element (row r, col k) sits at [k//8, r, k%8]
    t = tile_k_major(w1[:, :512])
    assert t.shape == (64, 512, 8)
    assert float(t[5, 17, 3]) == float(w1[17, 43].half())
    perm = fd_column_order(128)
    assert sorted(perm.tolist()) == list(range(768))
    assert perm[2 * 5] == 5 and perm[2 * 5 + 1] == 384 + 5 and perm[256 + 2 * 7 + 1] == 384 + 128 + 7
    # FiLM time table row t = W_cond[:, :128] @ time_emb(t)
    from chemeleon_b200.schedules import time_embedding_table

    te = time_embedding_table(1000, 128)
    wc = sd["decoder.film_layer.mlp_cond.0.weight"]
    assert torch.allclose(pw.film_time_table[321], wc[:, :128] @ te[321], atol=1e-5)
    assert pw.w_head.shape == (128, 512) and torch.equal(pw.w_head[104:107], sd["decoder.coord_out.weight"])


def test_frame_wire_format_roundtrip():
    """streaming.Frame parses the layout documented in include/chemeleon_b200.h (cb2_pack_frame)."""
    from chemeleon_b200.streaming import Frame

    N, B = 7, 2
    npad = (N + 3) & ~3
    buf = np.zeros(16 + npad + 12 * N + 36 * B, dtype=np.uint8)
    buf[:16].view(np.int32)[:] = [41, N, B, 0]
    buf[16:16 + N] = [8, 22, 8, 3, 25, 8, 0]
    x = np.arange(3 * N, dtype=np.float32) / 32.0
    l = np.arange(9 * B, dtype=np.float32) + 1
    buf[16 + npad:16 + npad + 12 * N] = x.view(np.uint8)
    buf[16 + npad + 12 * N:] = l.view(np.uint8)
    f = Frame.from_bytes(buf.tobytes())
    assert (f.t, f.n_nodes, f.n_graphs) == (41, N, B)
    assert f.types.tolist() == [8, 22, 8, 3, 25, 8, 0]
    assert np.array_equal(f.frac_coords.reshape(-1), x) and np.array_equal(f.lattices.reshape(-1), l)
    atoms = f.to_atoms([4, 3])
    assert [len(a) for a in atoms] == [4, 3]
    assert sorted(atoms[0].get_atomic_numbers().tolist()) == [3, 8, 8, 22]


def test_pack_weights_c_matches_python_packer(lib):
    """cb2_pack_weights (host C code, for non-Python hosts) produces bit-for-bit the operand images
    weights.pack_weights builds: K-major tiles, W_fd column permutation, W2 row blocks, split-precision head."""
    from chemeleon_b200 import _lib
    from chemeleon_b200.weights import fd_column_order, tile_k_major

    g = torch.Generator().manual_seed(0)

    def c_pack(kind, w):
        rows, K = w.shape
        nb = int(lib.cb2_pack_bytes(kind, rows, K))
        out = torch.zeros(nb // 2, dtype=torch.float16)
        w = w.contiguous()
        _lib.check(lib.cb2_pack_weights(kind, w.data_ptr(), rows, K, out.data_ptr(), nb), "cb2_pack_weights")
        return out

    w = torch.randn(512, 1024, generator=g) * 0.05
    assert torch.equal(c_pack(_lib.PACK_KMAJOR, w).view(torch.int16), tile_k_major(w).reshape(-1).view(torch.int16))
    wfd = torch.randn(512, 768, generator=g) * 0.03
    assert torch.equal(c_pack(_lib.PACK_FD, wfd).view(torch.int16),
                       tile_k_major(wfd[:, fd_column_order(128)]).reshape(-1).view(torch.int16))
    w2 = torch.randn(512, 512, generator=g) * 0.04
    blocks = torch.cat([tile_k_major(0.5 * w2[i:i + 128]) for i in range(0, 512, 128)], dim=0)     # image of W2 / 2
    assert torch.equal(c_pack(_lib.PACK_ROW_BLOCKS, w2).view(torch.int16), blocks.reshape(-1).view(torch.int16))
    # head: same construction as weights.head_split_image
    import math
    wh = torch.zeros(128, 512)
    wh[:107] = torch.randn(107, 512, generator=g) * 0.04
    s = 2.0 ** math.floor(math.log2(1024.0 / float(wh.abs().max())))
    ws = wh.double() * s
    hi = ws.to(torch.float16)
    lo = (ws - hi.double()).to(torch.float16)
    full = torch.zeros(256, 1536, dtype=torch.float16)
    full[:128] = torch.cat([hi, hi, lo], dim=1)
    blob = torch.zeros(8 + full.numel(), dtype=torch.float16)
    blob[:2].view(torch.float32)[0] = 1.0 / s
    blob[8:] = tile_k_major(full).reshape(-1)
    assert torch.equal(c_pack(_lib.PACK_HEAD_SPLIT, wh).view(torch.int16), blob.view(torch.int16))
    assert lib.cb2_pack_weights(_lib.PACK_KMAJOR, w.data_ptr(), 512, 1024, blob.data_ptr(), 16) == -5
