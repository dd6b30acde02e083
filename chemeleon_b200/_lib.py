"""ctypes binding of libchemeleon_b200.so (the C-ABI in include/chemeleon_b200.h).

The library is the product: if it is missing or cannot be loaded this module
raises -- there is no PyTorch / CPU fallback for any of its entry points.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libchemeleon_b200.so")

ABI_VERSION = 2
MAX_LAYERS = 16
HIDDEN = 512
HEAD_COLS = 128
COEF_COLS = 16
TILE_ROWS = 128
PRECISION_FP32 = 0
PRECISION_TC_F16 = 1
MODEL_EDGE_SINGLE_CTA = 1
MODEL_NODE_UNFUSED = 2
PACK_KMAJOR, PACK_FD, PACK_ROW_BLOCKS, PACK_HEAD_SPLIT = 0, 1, 2, 3
FLAG_NONFINITE = 1
FLAG_TC_RANGE = 2
TC_RANGE_LIMIT = 96.0

vp = C.c_void_p


class LayerWeights(C.Structure):
    _fields_ = [(n, vp) for n in (
        "w_hij", "w_ip", "b1", "w_fd", "w2", "b2", "wn1", "bn1", "wn2", "bn2", "ln_g", "ln_b",
        "w_hij_t", "w_fd_t", "w2_t", "wn1_t", "wn2_t")]


class Model(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("hidden", C.c_int32), ("n_layers", C.c_int32),
        ("n_atom_types", C.c_int32), ("n_freqs", C.c_int32), ("timesteps", C.c_int32),
        ("emb", vp), ("film_wp", vp), ("film_bp", vp), ("film_g", vp), ("film_b", vp),
        ("film_wp_t", vp), ("film_time_table", vp),
        ("layers", LayerWeights * MAX_LAYERS),
        ("final_g", vp), ("final_b", vp), ("w_head", vp), ("b_head", vp), ("w_head_t", vp), ("w_lat", vp),
        ("flags", C.c_int32),
    ]


class Batch(C.Structure):
    _fields_ = [
        ("n_nodes", C.c_int32), ("n_graphs", C.c_int32), ("n_variants", C.c_int32), ("max_n", C.c_int32),
        ("n_edges", C.c_int64),
        ("node2graph", vp), ("node_base", vp), ("node_n", vp), ("graph_off", vp),
        ("edge_i", vp), ("edge_j", vp), ("node_eoff", vp),
        ("n_chunks", C.c_int32),
        ("host_chunk_node_lo", vp), ("host_chunk_edge_lo", vp),
        ("chunk_max_edges", C.c_int64),
        ("n_tiles", C.c_int32),
        ("tile_row_i", vp), ("tile_row_j", vp), ("tile_seg_n", vp),
    ]


class ForwardIO(C.Structure):
    _fields_ = [
        ("atom_types", vp), ("frac_coords", vp), ("lattices", vp), ("film_cond", vp),
        ("head_out", vp), ("lattice_out", vp), ("node_features", vp),
        ("coords_only", C.c_int32), ("precision", C.c_int32), ("flags", vp),
    ]


class State(C.Structure):
    _fields_ = [("atom_types", vp), ("frac_coords", vp), ("lattices", vp), ("t_dev", vp), ("flags", vp)]


class StepArgs(C.Structure):
    _fields_ = [
        ("coef", vp), ("text_part", vp), ("text_row", vp),
        ("cond_scale", C.c_float), ("timesteps", C.c_int32), ("precision", C.c_int32),
        ("noise_mode", C.c_int32), ("t_start", C.c_int32),
        ("rand_a", vp), ("rand_l", vp), ("rand_x", vp), ("rand_x2", vp),
        ("seed", C.c_uint64), ("seed_dev", vp), ("graph_gid", vp),
    ]


class TextTail(C.Structure):
    _fields_ = [("embed_dim", C.c_int32), ("text_dim", C.c_int32)] + [(n, vp) for n in (
        "w1", "b1", "ln_g", "ln_b", "w2", "b2", "null_embeds", "w_text", "b_cond")]


EXPORTS = {
    # name: (restype, argtypes)
    "cb2_abi_version": (C.c_int, []),
    "cb2_last_error": (C.c_char_p, []),
    "cb2_check_device": (C.c_int, [C.c_int]),
    "cb2_workspace_bytes": (C.c_size_t, [C.POINTER(Model), C.POINTER(Batch), C.c_int]),
    "cb2_embed_nodes": (C.c_int, [C.POINTER(Model), C.POINTER(Batch), vp, vp, vp]),
    "cb2_film_cond": (C.c_int, [C.POINTER(Model), C.POINTER(Batch), vp, vp, vp, vp]),
    "cb2_text_condition": (C.c_int, [C.POINTER(TextTail), vp, C.c_int32, vp, vp, C.c_size_t, vp]),
    "cb2_text_condition_workspace_bytes": (C.c_size_t, [C.POINTER(TextTail), C.c_int32]),
    "cb2_linear_f32": (C.c_int, [vp, C.c_int64, vp, vp, vp, C.c_int64, C.c_int64, C.c_int32, C.c_int32,
                                 C.c_int32, vp]),
    "cb2_linear_tc": (C.c_int, [vp, C.c_int64, vp, C.c_int32, vp, vp, C.c_int64, C.c_int64, C.c_int32, C.c_int32,
                                vp, C.c_size_t, vp]),
    "cb2_linear_tc_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int32]),
    "cb2_edge_layer": (C.c_int, [C.POINTER(Model), C.c_int32, C.POINTER(Batch), vp, vp, vp, vp, C.c_int64, C.c_int32,
                                 vp, C.c_size_t, vp]),
    "cb2_decoder_forward": (C.c_int, [C.POINTER(Model), C.POINTER(Batch), C.POINTER(ForwardIO), vp,
                                      C.c_size_t, vp]),
    "cb2_update_predictor": (C.c_int, [C.POINTER(Batch), C.POINTER(State), C.POINTER(StepArgs), vp, vp, vp]),
    "cb2_update_corrector": (C.c_int, [C.POINTER(Batch), C.POINTER(State), C.POINTER(StepArgs), vp, vp]),
    "cb2_sampler_step": (C.c_int, [C.POINTER(Model), C.POINTER(Batch), C.POINTER(State), C.POINTER(StepArgs),
                                   vp, C.c_size_t, vp]),
    "cb2_launch_count": (C.c_uint64, []),
    "cb2_pack_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32]),
    "cb2_pack_weights": (C.c_int, [C.c_int32, vp, C.c_int32, C.c_int32, vp, C.c_size_t]),
    "cb2_frame_bytes": (C.c_size_t, [C.c_int32, C.c_int32]),
    "cb2_pack_frame": (C.c_int, [C.POINTER(Batch), C.POINTER(State), vp, C.c_size_t, vp]),
    "cb2_validity_filter": (C.c_int, [vp, vp, vp, vp, C.c_int32, vp, C.c_float, C.c_float, vp, vp, vp, vp]),
}


class Cb2Error(RuntimeError):
    pass


_lib = None


def load(build_if_missing: bool = True):
    """Load the shared library (building it in-tree with nvcc if absent)."""
    global _lib
    if _lib is not None:
        return _lib
    from . import build as _build

    if not os.path.exists(LIB_PATH):
        if not build_if_missing:
            raise Cb2Error(f"{LIB_PATH} is missing and there is no fallback path; run "
                           "`python -m chemeleon_b200.build`")
        _build.build()
    elif build_if_missing and _build.have_nvcc() and _build.needs_build():
        _build.build()          # csrc/*.cu edited since the library was built: never run stale kernels
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in EXPORTS.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:  # pragma: no cover
            raise Cb2Error(f"{LIB_PATH} does not export {name}") from e
        fn.restype = res
        fn.argtypes = args
    if lib.cb2_abi_version() != ABI_VERSION:
        raise Cb2Error("libchemeleon_b200.so ABI version mismatch; rebuild with `python -m chemeleon_b200.build --force`")
    _lib = lib
    return lib


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = load().cb2_last_error()
        raise Cb2Error(f"{what} failed ({status}): {msg.decode() if msg else ''}")


def ptr(t) -> int:
    """Device/host pointer of a torch tensor (or None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr()
