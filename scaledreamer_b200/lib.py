"""ctypes binding of libsdb200.so (the C ABI declared in include/sdb200.h and include/sdb200_nn.h).

The product path has no CPU fallback: if the shared library is missing, or a call returns an error code,
a RuntimeError is raised. Tensors cross the boundary as raw device pointers.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsdb200.so")


class GridCfgC(C.Structure):
    _fields_ = [("n_levels", C.c_int), ("n_features_per_level", C.c_int), ("log2_hashmap_size", C.c_int),
                ("base_resolution", C.c_int), ("per_level_scale", C.c_float)]


class FieldC(C.Structure):
    _fields_ = [
        ("grid", GridCfgC), ("table", C.c_void_p), ("w1_density", C.c_void_p), ("w2_density", C.c_void_p),
        ("w1_feature", C.c_void_p), ("w2_feature", C.c_void_p), ("radius", C.c_float),
        ("density_bias_type", C.c_int), ("density_bias_const", C.c_float), ("density_blob_scale", C.c_float),
        ("density_blob_std", C.c_float), ("density_activation", C.c_int), ("fd_normal_eps", C.c_float),
        ("color_activation", C.c_int), ("bg_grid", GridCfgC), ("bg_table", C.c_void_p), ("bg_w1", C.c_void_p),
        ("bg_w2", C.c_void_p), ("bg_w3", C.c_void_p), ("bg_color_activation", C.c_int)]


class FieldGradsC(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("table", "w1_density", "w2_density", "w1_feature", "w2_feature",
                                          "bg_table", "bg_w1", "bg_w2", "bg_w3")]


class MarchCfgC(C.Structure):
    _fields_ = [("render_step_size", C.c_float), ("near_plane", C.c_float), ("far_plane", C.c_float),
                ("prune", C.c_int), ("alpha_thre", C.c_float), ("early_stop_eps", C.c_float),
                ("grid_resolution", C.c_int), ("output_normal", C.c_int)]


class PackedSamplesC(C.Structure):
    _fields_ = [("counter", C.c_void_p), ("capacity", C.c_int), ("ray_indices", C.c_void_p),
                ("t_starts", C.c_void_p), ("t_ends", C.c_void_p), ("weights", C.c_void_p), ("density", C.c_void_p),
                ("rgb", C.c_void_p), ("normal", C.c_void_p)]


class RenderTapeC(C.Structure):
    _fields_ = [("capacity", C.c_int), ("max_chunks", C.c_int), ("counter", C.c_void_p), ("enc", C.c_void_p),
                ("pos", C.c_void_p), ("sample", C.c_void_p), ("ray_chunks", C.c_void_p), ("ray_nchunks", C.c_void_p)]


class GemmArgsC(C.Structure):
    _fields_ = [("A", C.c_void_p), ("lda", C.c_longlong), ("B", C.c_void_p), ("ldb", C.c_longlong),
                ("out", C.c_void_p), ("ldc", C.c_longlong), ("M", C.c_int), ("N", C.c_int), ("K", C.c_int),
                ("bias", C.c_void_p), ("rowbias", C.c_void_p), ("rows_per_group", C.c_int), ("residual", C.c_void_p),
                ("ldr", C.c_longlong), ("alpha", C.c_float), ("act", C.c_int), ("out_fp32", C.c_int),
                ("batch", C.c_int), ("a_zs", C.c_longlong), ("b_zs", C.c_longlong), ("out_zs", C.c_longlong)]


class GemmTf32ArgsC(C.Structure):
    _fields_ = [("A", C.c_void_p), ("lda", C.c_longlong), ("a_zs_hi", C.c_longlong), ("a_zs_lo", C.c_longlong),
                ("B", C.c_void_p), ("ldb", C.c_longlong), ("b_zs_hi", C.c_longlong), ("b_zs_lo", C.c_longlong),
                ("M", C.c_int), ("N", C.c_int), ("K", C.c_int), ("batch", C.c_int), ("zdiv", C.c_int),
                ("out", C.c_void_p), ("ldc", C.c_longlong), ("out_zs_hi", C.c_longlong), ("out_zs_lo", C.c_longlong),
                ("bias", C.c_void_p), ("residual", C.c_void_p), ("ldr", C.c_longlong), ("res_zs_hi", C.c_longlong),
                ("res_zs_lo", C.c_longlong), ("alpha", C.c_float), ("act", C.c_int), ("round_out", C.c_int),
                ("a_mn_major", C.c_int)]


class UNetCfgC(C.Structure):
    _fields_ = [("in_channels", C.c_int), ("out_channels", C.c_int), ("model_channels", C.c_int),
                ("num_levels", C.c_int), ("channel_mult", C.c_int * 4), ("num_res_blocks", C.c_int),
                ("attn_levels", C.c_int), ("head_dim", C.c_int), ("context_dim", C.c_int), ("context_len", C.c_int),
                ("camera_dim", C.c_int), ("num_frames", C.c_int)]


class VaeCfgC(C.Structure):
    _fields_ = [("in_channels", C.c_int), ("ch", C.c_int), ("num_levels", C.c_int), ("ch_mult", C.c_int * 4),
                ("num_res_blocks", C.c_int), ("z_channels", C.c_int)]


class PromptCfgC(C.Structure):
    _fields_ = [("view_dependent", C.c_int), ("perp_neg", C.c_int), ("front_threshold", C.c_float),
                ("back_threshold", C.c_float), ("overhead_threshold", C.c_float), ("f_sb", C.c_float * 3),
                ("f_fsb", C.c_float * 3), ("f_fs", C.c_float * 3), ("f_sf", C.c_float * 3), ("neg_scale", C.c_float)]


_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Loads libsdb200.so (built in-tree by __graft_entry__.build()). Raises if absent: no fallback exists."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
            "scaledreamer_b200 has no CPU or PyTorch fallback for its CUDA path.")
    lib = C.CDLL(LIB_PATH)
    lib.sdb_last_error.restype = C.c_char_p
    lib.sdb_launch_count.restype = C.c_ulonglong
    lib.sdb_grid_num_entries.restype = C.c_longlong
    _declare(lib)
    _lib = lib
    return lib


class _TimedLib:
    """Measurement aid (bench.py's per-kernel breakdown): every C-ABI call that takes a stream is bracketed by CUDA
    events on the CURRENT stream, summed per entry point. Installed by call_timer_begin(), removed by call_timer_end();
    the product path never sees it otherwise."""

    def __init__(self, lib: C.CDLL):
        self._lib = lib
        self.records = {}

    def __getattr__(self, name):
        fn = getattr(self._lib, name)
        sig = SIGNATURES.get(name)
        if not name.startswith("sdb_") or not sig or sig[-1] is not _P or name.startswith("sdb_gemm_profile"):
            return fn

        def timed(*args):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n0 = self._lib.sdb_launch_count()
            a.record()
            rc = fn(*args)
            b.record()
            self.records.setdefault(name, []).append((a, b, self._lib.sdb_launch_count() - n0))
            return rc

        return timed


def call_timer_begin() -> None:
    global _lib
    lib = load()
    if not isinstance(lib, _TimedLib):
        _lib = _TimedLib(lib)


def call_timer_end() -> dict:
    """-> {entry point: {"calls": n, "ms": total device time, "launches": kernels launched}} since call_timer_begin()."""
    global _lib
    if not isinstance(_lib, _TimedLib):
        return {}
    timed, _lib = _lib, _lib._lib
    torch.cuda.synchronize()
    return {k: {"calls": len(v), "ms": sum(a.elapsed_time(b) for a, b, _ in v), "launches": int(sum(n for _, _, n in v))}
            for k, v in timed.records.items()}


_P, _I, _F, _LL = C.c_void_p, C.c_int, C.c_float, C.c_longlong

# name -> argtypes; every function returns int unless listed in load() above. Pointers are void* so that
# raw tensor addresses (python ints) are passed at full width.
SIGNATURES = {
    "sdb_grid_num_entries": [C.POINTER(GridCfgC)],
    "sdb_grid_describe": [C.POINTER(GridCfgC), _P, _P, _P, _P, _P],
    "sdb_hashgrid_forward": [C.POINTER(GridCfgC), _P, _P, _I, _P, _P],
    "sdb_hashgrid_backward": [C.POINTER(GridCfgC), _P, _P, _I, _P, _P],
    "sdb_field_forward": [C.POINTER(FieldC), _P, _I, _P, _P, _P, _P],
    "sdb_occgrid_update": [C.POINTER(FieldC), _P, _P, _I, _I, _F, _F, _F, _P, _P, _P, _P],
    "sdb_render_nerf_forward": [C.POINTER(FieldC), C.POINTER(MarchCfgC), _P, _P, _P, _P, _P, _P, _I, _I,
                                _P, _P, _P, _P, _P, _P, C.POINTER(PackedSamplesC), _P, _P],
    "sdb_render_nerf_backward": [C.POINTER(FieldC), C.POINTER(FieldGradsC), C.POINTER(MarchCfgC), _P, _P, _P, _P,
                                 _P, _P, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "sdb_render_tape_geometry": [C.POINTER(MarchCfgC), _F, _I, C.POINTER(C.c_longlong), C.POINTER(C.c_int)],
    "sdb_render_nerf_forward_v2": [C.POINTER(FieldC), C.POINTER(MarchCfgC), _P, _P, _P, _P, _P, _P, _I, _I,
                                   _P, _P, _P, _P, _P, _P, C.POINTER(RenderTapeC), _P, _P],
    "sdb_render_nerf_backward_tape": [C.POINTER(FieldC), C.POINTER(FieldGradsC), C.POINTER(MarchCfgC), _P, _P, _I, _I,
                                      _P, _P, _P, _P, _P, _P, _P, C.POINTER(RenderTapeC), _P],
    "sdb_render_nerf_backward_tape_zv": [C.POINTER(FieldC), C.POINTER(FieldGradsC), C.POINTER(MarchCfgC), _P, _P, _I,
                                         _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, C.POINTER(RenderTapeC), _P],
    "sdb_render_orient_forward": [C.POINTER(FieldC), _P, _I, C.POINTER(RenderTapeC), _P, _P, _P],
    "sdb_render_orient_backward": [C.POINTER(FieldC), C.POINTER(FieldGradsC), _I, C.POINTER(RenderTapeC), _P, _P, _P],
    "sdb_hypernet_forward": [_P, _I, _I, _P, _P, _P, _F, _P, _P, _I, _P, _P, _P],
    "sdb_hypernet_scratch_floats": [_I, _I],
    "sdb_hypernet_backward": [_P, _I, _I, _P, _P, _F, _P, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "sdb_hyper_field_tape_floats": [_I, _I],
    "sdb_hyper_field_forward": [C.POINTER(GridCfgC), _P, _P, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P],
    "sdb_hyper_field_backward": [C.POINTER(GridCfgC), _P, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "sdb_volsdf_coarse_points": [_P, _P, _P, _I, _I, _F, _F, _P, _P],
    "sdb_volsdf_resample": [_P, _P, _P, _I, _I, _I, _F, _F, _F, _P, _P],
    "sdb_volsdf_composite_forward": [_P, _P, _P, _P, _P, _I, _I, _F, _I, _P, _P, _P, _P, _P, _P, _P],
    "sdb_volsdf_composite_backward": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _F, _I, _P, _P, _P],
    "sdb_triplane_sample_forward": [_P, _P, _I, _I, _I, _I, _I, _P, _P],
    "sdb_triplane_sample_backward": [_P, _P, _I, _I, _I, _I, _I, _P, _P],
    "sdb_raygen": [_P, _P, _I, _I, _I, _P, _P, _P],
    "sdb_adamw_step": [_P, _P, _P, _P, _LL, _F, _F, _F, _F, _F, _I, _F, _P],
    "sdb_mlp3_forward": [_P, _LL, _I, _P, _P, _P, _I, _P, _P],
    "sdb_mlp3_backward": [_P, _LL, _I, _P, _P, _P, _I, _P, _P, _I, _P, _P, _P, _P],
    "sdb_march_count": [C.POINTER(MarchCfgC), _F, _P, _P, _P, _P, _I, _P, _P],
    "sdb_march_fill": [C.POINTER(MarchCfgC), _F, _P, _P, _P, _P, _I, _P, _P, _P, _P, _P, _P],
    "sdb_packed_visibility": [_P, _P, _P, _P, _I, _F, _P, _F, _P, _P, _P],
    "sdb_packed_composite_forward": [_P, _P, _P, _P, _P, _I, _P, _P, _P, _P, _P, _P, _P],
    "sdb_packed_composite_backward": [_P, _P, _P, _P, _I, _P, _P, _P, _P, _P, _P, _P, _P],
    "sdb_freq_encode": [_P, _LL, _I, _P, _I, _I, _P, _P],
    "sdb_occgrid_update_values": [_P, _P, _I, _I, _F, _F, _P, _P, _P, _P],
    "sdb_adan_step": [_P, _P, _P, _P, _P, _P, _LL, _F, _F, _F, _F, _F, _F, _I, _F, _I, _P],
    # ---- include/sdb200_nn.h
    "sdb_gemm_f16": [C.POINTER(GemmArgsC), _P],
    "sdb_gemm_profile_begin": [],
    "sdb_gemm_profile_dump": [C.c_char_p],
    "sdb_gemm_debug_timeline": [_P],
    "sdb_gemm_profile_end": [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int)],
    "sdb_conv3x3_f16": [_P, _I, _I, _I, _I, _P, _I, _P, _P, _P, _I, _P, _P],
    "sdb_conv3x3_small": [_P, _I, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "sdb_attention_f16": [_P, _LL, _P, _LL, _P, _LL, _I, _I, _I, _I, _P, _P, _LL, _P],
    "sdb_flash_attention_f16": [_P, _LL, _P, _LL, _P, _LL, _I, _I, _I, _I, _P, _LL, _P],
    "sdb_groupnorm_workspace_floats": [_I, _I, _I, _I],
    "sdb_groupnorm_f16": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _I, _P],
    "sdb_groupnorm_backward_f16": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _I, _P],
    "sdb_layernorm_f16": [_P, _P, _P, _P, _I, _I, _F, _P],
    "sdb_geglu_f16": [_P, _P, _LL, _I, _P],
    "sdb_upsample2x_f16": [_P, _P, _I, _I, _I, _I, _P],
    "sdb_im2col3x3s2_f16": [_P, _P, _I, _I, _I, _I, _I, _P],
    "sdb_col2im3x3s2_f16": [_P, _P, _I, _I, _I, _I, _I, _P],
    "sdb_gemm_tf32": [C.POINTER(GemmTf32ArgsC), _P],
    "sdb_round_tf32_f32": [_P, _P, _LL, _P],
    "sdb_transpose_f32": [_P, _LL, _LL, _P, _LL, _LL, _I, _I, _I, _I, _P],
    "sdb_layernorm_f32_forward": [_P, _P, _P, _P, _P, _P, _I, _I, _F, _I, _P],
    "sdb_layernorm_f32_backward_ws_floats": [_I, _I],
    "sdb_layernorm_f32_backward": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _P],
    "sdb_softmax_f32_forward": [_P, _LL, _I, _LL, _P, _I, _P],
    "sdb_softmax_f32_backward_rows": [_P, _P, _LL, _I, _LL, _P, _P, _I, _I, _P],
    "sdb_softmax_f32_backward_stats": [_P, _P, _I, _I, _I, _LL, _P, _P, _I, _I, _P],
    "sdb_attn_delta_f32": [_P, _P, _P, _I, _I, _I, _I, _P],
    "sdb_gelu_f32_forward": [_P, _P, _LL, _I, _P],
    "sdb_gelu_f32_backward": [_P, _P, _LL, _I, _P],
    "sdb_colsum_f32_ws_floats": [_LL, _I],
    "sdb_colsum_f32": [_P, _LL, _I, _LL, _P, _P, _P],
    "sdb_broadcast_f32": [_P, _LL, _P, _I, _P],
    "sdb_deconv_shuffle_f32": [_P, _P, _I, _I, _I, _I, _I, _P],
    "sdb_unet_create": [C.POINTER(UNetCfgC), _I, _I, _I, C.POINTER(C.c_void_p)],
    "sdb_unet_forward": [_P, _P, _P, _P, _P, _P, _P],
    "sdb_vae_encoder_create": [C.POINTER(VaeCfgC), _I, _I, _I, C.POINTER(C.c_void_p)],
    "sdb_vae_encoder_forward": [_P, _P, _P, _P],
    "sdb_vae_encoder_backward": [_P, _P, _P, _P],
    "sdb_net_destroy": [_P],
    "sdb_net_sizes": [_P, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)],
    "sdb_net_bind": [_P, _P, _P],
    "sdb_net_num_params": [_P],
    "sdb_net_param": [_P, _I, C.POINTER(C.c_char_p), C.POINTER(C.c_int), C.POINTER(C.c_int)],
    "sdb_net_load_param": [_P, C.c_char_p, _P, _LL, _P],
    "sdb_net_finalize": [_P, _P],
    "sdb_net_num_launches": [_P, _I],
    # ---- include/sdb200_asd.h
    "sdb_resize_bilinear_forward": [_P, _I, _I, _I, _I, _P, _I, _I, _F, _F, _P],
    "sdb_resize_bilinear_backward": [_P, _I, _I, _I, _I, _P, _I, _I, _F, _P],
    "sdb_asd_text_embeddings": [C.POINTER(PromptCfgC), _P, _P, _P, _P, _I, _I, _I, _P, _P, _P],
    "sdb_asd_text_embeddings_multi": [C.POINTER(PromptCfgC), _P, _P, _P, _I, _P, _P, _I, _I, _I, _P, _P, _P],
    "sdb_asd_prologue": [_P, _P, _P, _P, _P, _P, _P, _P, _F, _I, _I, _I, _P, _P, _P, _P],
    "sdb_asd_epilogue": [_P, _P, _P, _P, _P, _P, _P, _P, _F, _I, _F, _F, _F, _I, _I, _I, _P, _P, _P, _P, _P],
    "sdb_asd_t_plus": [_P, _P, _I, _F, _I, _I, _P, _P],
}


def _declare(lib: C.CDLL) -> None:
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch: fail loudly
        fn.argtypes = args
        if name == "sdb_net_destroy":
            fn.restype = None
        elif name in ("sdb_groupnorm_workspace_floats", "sdb_hyper_field_tape_floats", "sdb_hypernet_scratch_floats",
                      "sdb_layernorm_f32_backward_ws_floats", "sdb_colsum_f32_ws_floats"):
            fn.restype = C.c_longlong
        elif name != "sdb_grid_num_entries":
            fn.restype = C.c_int


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{what} failed (code {rc}): {load().sdb_last_error().decode()}")


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("scaledreamer_b200 kernels need CUDA tensors (there is no CPU path)")
    if not t.is_contiguous():
        raise RuntimeError("non-contiguous tensor passed to the C ABI")
    return t.data_ptr()


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def launch_count() -> int:
    return int(load().sdb_launch_count())


def grid_cfg_c(cfg) -> GridCfgC:
    """cfg: mapping/obj with the tcnn HashGrid keys."""
    get = (lambda k: cfg[k]) if isinstance(cfg, dict) else (lambda k: getattr(cfg, k))
    return GridCfgC(int(get("n_levels")), int(get("n_features_per_level")), int(get("log2_hashmap_size")),
                    int(get("base_resolution")), float(get("per_level_scale")))


def grid_num_entries(cfg) -> int:
    c = grid_cfg_c(cfg)
    n = load().sdb_grid_num_entries(C.byref(c))
    if n < 0:
        raise RuntimeError(f"bad grid config: {load().sdb_last_error().decode()}")
    return int(n)
