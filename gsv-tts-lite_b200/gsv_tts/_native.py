"""ctypes binding of ``libgsv_b200.so`` (C ABI declared in ``include/gsv_b200.h``).

The library is built in-tree by ``__graft_entry__.build()`` / ``csrc/Makefile``.  There is no
CPU or PyTorch fallback: if the library is missing, or the device is not sm_100, importing
the compute path raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgsv_b200.so")

GSV_F16, GSV_BF16 = 0, 1


class GptDims(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "d_model", "n_head", "n_layer", "d_ff", "vocab", "eos", "n_phoneme", "d_bert", "n_pos",
        "dtype", "max_slots", "max_seq")]


class GptWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "w_qkv", "b_qkv", "w_o", "b_o", "w_1", "b_1", "w_2", "b_2", "ln1_g", "ln1_b", "ln2_g", "ln2_b",
        "w_head", "emb_audio", "pe_audio", "emb_text", "pe_text", "w_bert", "b_bert")]


class GptSampling(C.Structure):
    _fields_ = [("top_k", C.c_int32), ("top_p", C.c_float), ("temperature", C.c_float),
                ("repetition_penalty", C.c_float), ("suppress_steps", C.c_int32),
                ("max_new_tokens", C.c_int32), ("mask_eos", C.c_int32), ("max_kv", C.c_int32),
                ("suppress_first", C.c_int32), ("reserved", C.c_int32), ("seed", C.c_uint64)]


class VocDims(C.Structure):
    _fields_ = [("inter_channels", C.c_int32), ("hidden_channels", C.c_int32), ("gin_channels", C.c_int32),
                ("n_flows", C.c_int32), ("wn_layers", C.c_int32), ("wn_kernel", C.c_int32),
                ("upsample_initial_channel", C.c_int32), ("n_ups", C.c_int32),
                ("upsample_rates", C.c_int32 * 8), ("upsample_kernel_sizes", C.c_int32 * 8),
                ("n_resblock_kernels", C.c_int32), ("resblock_kernel_sizes", C.c_int32 * 4),
                ("resblock_dilations", (C.c_int32 * 3) * 4), ("dtype", C.c_int32)]


class EncpDims(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "hidden_channels", "filter_channels", "inter_channels", "n_heads", "n_layers", "kernel_size", "ssl_dim", "n_codes",
        "n_symbols", "mrte_channels", "mrte_heads", "gin_channels", "dtype")]


# every symbol include/gsv_b200.h declares: (name, restype, argtypes)
_P = C.c_void_p
SYMBOLS = [
    ("gsv_last_error", C.c_char_p, []),
    ("gsv_version", C.c_int, []),
    ("gsv_device_check", C.c_int, [C.c_int]),
    ("gsv_gpt_create", C.c_int, [C.POINTER(GptDims), C.POINTER(GptWeights), C.POINTER(_P)]),
    ("gsv_gpt_destroy", C.c_int, [_P]),
    ("gsv_gpt_prefill", C.c_int, [_P, C.c_int, _P, C.c_int, _P, C.c_int, _P, C.POINTER(GptSampling), _P]),
    ("gsv_gpt_prefill_begin", C.c_int, [_P, C.c_int, _P, C.c_int, _P, C.c_int, _P, _P]),
    ("gsv_gpt_prefill_finish", C.c_int, [_P, C.c_int, _P, C.c_int, C.POINTER(GptSampling), _P]),
    ("gsv_gpt_prefill_begin_many", C.c_int, [_P, C.c_int, C.POINTER(C.c_int), C.POINTER(_P), C.POINTER(C.c_int), C.POINTER(_P),
                                            C.POINTER(C.c_int), C.POINTER(_P), _P]),
    ("gsv_gpt_prefill_capacity", C.c_int, [_P]),
    ("gsv_gpt_decode", C.c_int, [_P, C.c_int, _P]),
    ("gsv_gpt_read", C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, _P]),
    ("gsv_gpt_state_ptrs", C.c_int, [_P, C.POINTER(_P), C.POINTER(_P), C.POINTER(_P)]),
    ("gsv_gpt_release_slot", C.c_int, [_P, C.c_int, _P]),
    ("gsv_gpt_set_noise", C.c_int, [_P, _P, C.c_int]),
    ("gsv_gpt_set_forced", C.c_int, [_P, _P, C.c_int]),
    ("gsv_gpt_set_logits_trace", C.c_int, [_P, _P, C.c_int]),
    ("gsv_gpt_set_slot_hooks", C.c_int, [_P, C.c_int, _P, C.c_int, _P, C.c_int, _P, C.c_int, _P]),
    ("gsv_gpt_launch_count", C.c_int64, [_P]),
    ("gsv_gpt_set_decode_sms", C.c_int, [_P, C.c_int]),
    ("gsv_gpt_wait_resident", C.c_int, [_P, _P]),
    ("gsv_gpt_set_timeline", C.c_int, [_P, _P, C.c_int, C.c_int]),
    ("gsv_voc_create", C.c_int, [C.POINTER(VocDims), C.POINTER(_P)]),
    ("gsv_voc_set_weight", C.c_int, [_P, C.c_char_p, _P, _P]),
    ("gsv_voc_destroy", C.c_int, [_P]),
    ("gsv_voc_flow_dec", C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int, _P, _P]),
    ("gsv_voc_set_debug_z", C.c_int, [_P, _P]),
    ("gsv_voc_launch_count", C.c_int64, [_P]),
    ("gsv_voc_graph_count", C.c_int, [_P]),
    ("gsv_encp_create", C.c_int, [C.POINTER(EncpDims), C.POINTER(_P)]),
    ("gsv_encp_set_weight", C.c_int, [_P, C.c_char_p, _P, _P]),
    ("gsv_encp_destroy", C.c_int, [_P]),
    ("gsv_encp_output_frames", C.c_int, [_P, C.c_int, C.c_float, C.c_int, C.c_int]),
    ("gsv_encp_forward", C.c_int, [_P, _P, C.c_int, _P, C.c_int, _P, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, _P, C.c_int,
                                   _P, C.c_float, C.c_uint64, _P, _P, _P, _P, C.POINTER(C.c_int), _P]),
    ("gsv_encp_reset_stream", C.c_int, [_P]),
    ("gsv_encp_stream_rollback", C.c_int, [_P]),
    ("gsv_encp_reuse_text", C.c_int, [_P, C.c_int]),
    ("gsv_encp_launch_count", C.c_int64, [_P]),
    ("gsv_glue_viterbi_monotonic", C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P, _P, _P]),
    ("gsv_glue_silence_offset", C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P]),
    ("gsv_glue_sola", C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P]),
]

_lib = None


class NativeError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Load the shared library (once) and bind every declared symbol; raises if absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback for the B200 hot path)")
        l = C.CDLL(LIB_PATH)
        for name, res, args in SYMBOLS:
            fn = getattr(l, name)          # AttributeError if the .so does not export it
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        msg = lib().gsv_last_error()
        raise NativeError(f"libgsv_b200 error {rc}: {msg.decode() if msg else ''}")


def dtype_code(torch_dtype) -> int:
    import torch
    if torch_dtype == torch.float16:
        return GSV_F16
    if torch_dtype == torch.bfloat16:
        return GSV_BF16
    raise NativeError(f"the B200 path computes in fp16 or bf16 storage with fp32 accumulation, not {torch_dtype}")
