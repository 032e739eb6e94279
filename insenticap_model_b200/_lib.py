"""ctypes binding of libisc_b200.so (include/isc.h). PyTorch supplies device memory and streams;
every compute call goes through the C ABI with raw device pointers. There is NO fallback: a missing
library or a non-sm_100 device raises immediately.
"""
from __future__ import annotations

import ctypes as C
import os

import torch  # noqa: F401  (also makes sure libcudart.so.12 is mapped before our library)

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libisc_b200.so")

PREC_FP32, PREC_BF16X3, PREC_BF16 = 0, 1, 2
PRECISIONS = {"fp32": PREC_FP32, "bf16x3": PREC_BF16X3, "bf16": PREC_BF16}


class Dims(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("vocab", "hidden", "feat_dim", "n_regions", "n_senti", "n_labels",
                 "pad_id", "sos_id", "eos_id", "unk_id", "att_tile")]


WEIGHT_FIELDS = [
    # (struct field, state_dict key) in include/isc.h order
    ("word_embed", "word_embed.0.weight"),
    ("senti_label_embed", "senti_label_embed.0.weight"),
    ("fc_embed_w", "fc_embed.0.weight"), ("fc_embed_b", "fc_embed.0.bias"),
    ("cpt2fc_w", "cpt2fc.0.weight"), ("cpt2fc_b", "cpt2fc.0.bias"),
    ("att_embed_w", "att_embed.0.weight"), ("att_embed_b", "att_embed.0.bias"),
    ("att_lstm_w_ih", "att_lstm.weight_ih"), ("att_lstm_w_hh", "att_lstm.weight_hh"),
    ("att_lstm_b_ih", "att_lstm.bias_ih"), ("att_lstm_b_hh", "att_lstm.bias_hh"),
    ("att2att_w", "att2att.0.weight"), ("att2att_b", "att2att.0.bias"),
    ("senti2att_w", "senti2att.0.weight"), ("senti2att_b", "senti2att.0.bias"),
    ("ca_h2att_w", "attention.cont_att.h2att.weight"), ("ca_h2att_b", "attention.cont_att.h2att.bias"),
    ("ca_alpha_w", "attention.cont_att.att_alpha.weight"), ("ca_alpha_b", "attention.cont_att.att_alpha.bias"),
    ("sa_h2word_w", "attention.senti_att.h2word.weight"), ("sa_h2word_b", "attention.senti_att.h2word.bias"),
    ("sa_label2word_w", "attention.senti_att.label2word.weight"),
    ("sa_label2word_b", "attention.senti_att.label2word.bias"),
    ("sa_alpha_w", "attention.senti_att.word_alpha.weight"), ("sa_alpha_b", "attention.senti_att.word_alpha.bias"),
    ("g_h2att_w", "attention.h2att.weight"), ("g_h2att_b", "attention.h2att.bias"),
    ("g_cont2att_w", "attention.cont2att.weight"), ("g_cont2att_b", "attention.cont2att.bias"),
    ("g_senti2att_w", "attention.senti2att.weight"), ("g_senti2att_b", "attention.senti2att.bias"),
    ("g_alpha_w", "attention.att_alpha.weight"), ("g_alpha_b", "attention.att_alpha.bias"),
    ("lang_lstm_w_ih", "lang_lstm.weight_ih"), ("lang_lstm_w_hh", "lang_lstm.weight_hh"),
    ("lang_lstm_b_ih", "lang_lstm.bias_ih"), ("lang_lstm_b_hh", "lang_lstm.bias_hh"),
    ("classifier_w", "classifier.weight"), ("classifier_b", "classifier.bias"),
]


class Weights(C.Structure):
    _fields_ = [(f, C.c_void_p) for f, _ in WEIGHT_FIELDS]


FEAT_FIELDS = ("fc", "att", "p_att", "sw", "p_sw", "sl", "pre_gates", "pre_word", "cpt_feats", "att16", "p_att16",
               "feat_flags")


class Feats(C.Structure):
    _fields_ = [(f, C.c_void_p) for f in FEAT_FIELDS]


class Dropout(C.Structure):
    _fields_ = [(f, C.c_void_p) for f in ("fc", "att", "sw", "sl", "out")] + [("scale", C.c_float)]


class Grads(C.Structure):
    _fields_ = [(f, C.c_void_p) for f, _ in WEIGHT_FIELDS]


class SchedSampling(C.Structure):
    _fields_ = [("prob", C.c_float), ("uniform", C.c_void_p), ("noise", C.c_void_p), ("seed", C.c_uint64)]


MODE_XE, MODE_SEQ2SEQ, MODE_RL = 0, 1, 2

_vp, _i32, _i64, _sz, _u64, _dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_size_t, C.c_uint64, C.c_double
_PD, _PW, _PF = C.POINTER(Dims), C.POINTER(Weights), C.POINTER(Feats)
_PDR, _PG, _f32 = C.POINTER(Dropout), C.POINTER(Grads), C.c_float

# name -> (restype, argtypes); must list every symbol include/isc.h declares (tests/test_abi.py)
SIGNATURES = {
    "isc_version": (C.c_char_p, []),
    "isc_last_error": (C.c_char_p, []),
    "isc_check_device": (C.c_int, []),
    "isc_packed_weights_bytes": (_sz, [_PD, C.c_int]),
    "isc_pack_weights": (C.c_int, [_PD, _PW, C.c_int, _vp, _sz, _vp]),
    "isc_prologue_workspace_bytes": (_sz, [_PD, C.c_int, C.c_int]),
    "isc_prologue": (C.c_int, [_PD, _vp, C.c_int, _vp, _vp, _vp, C.c_int, _vp, _vp, C.c_int, C.c_int, _PF,
                               _vp, _sz, _vp, _PDR]),
    "isc_convert_features": (C.c_int, [C.c_int, C.c_int, _vp, _vp, _i64, _vp]),
    "isc_hoist": (C.c_int, [_PD, _vp, C.c_int, C.c_int, _PF, _vp, _sz, _vp]),
    "isc_decode_workspace_bytes": (_sz, [_PD, C.c_int, C.c_int]),
    "isc_decode_step": (C.c_int, [_PD, _vp, C.c_int, _PF, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _i64,
                                  _vp, _vp, _vp, _vp, _sz, _vp]),
    "isc_decode_greedy": (C.c_int, [_PD, _vp, C.c_int, _PF, C.c_int, C.c_int, C.c_int, _vp, _u64, _vp, _vp, _vp,
                                    _vp, _vp, _vp, _vp, _sz, _vp, _vp, _f32]),
    "isc_decode_beam": (C.c_int, [_PD, _vp, C.c_int, _PF, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp,
                                  _vp, _sz, _vp]),
    "isc_teacher_forced": (C.c_int, [_PD, _vp, C.c_int, _PF, C.c_int, C.c_int, _vp, _i64, _vp, _vp, _sz, _vp]),
    "isc_train_workspace_bytes": (_sz, [_PD, C.c_int, C.c_int, C.c_int]),
    "isc_train_forward": (C.c_int, [_PD, _vp, C.c_int, C.c_int, _vp, _vp, _vp, C.c_int, _vp, _vp, C.c_int, _vp, _i64, C.c_int,
                                    _PDR, C.POINTER(SchedSampling), _vp, _vp, _vp, _vp, _sz, _vp]),
    "isc_train_forward_sample": (C.c_int, [_PD, _vp, C.c_int, C.c_int, _vp, _vp, _vp, C.c_int, _vp, _vp, C.c_int, C.c_int,
                                           _PDR, C.c_int, _vp, _u64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "isc_expand_f16": (C.c_int, [_vp, _vp, _i64, _vp]),
    "isc_train_backward_marks": (C.c_int, [_vp, _vp]),
    "isc_train_backward": (C.c_int, [_PD, _vp, C.c_int, C.c_int, _vp, _vp, _vp, C.c_int, _vp, _vp, C.c_int, _vp, _i64, C.c_int,
                                     _PDR, _vp, _vp, _vp, _i64, _vp, _vp, _PG, _vp, _sz, _vp]),
    "isc_adam_step": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _f32, _f32, _f32, _f32, _f32, _f32, C.c_int, _f32, _vp]),
    "isc_senti_packed_bytes": (_sz, [C.c_int]),
    "isc_senti_pack": (C.c_int, [C.c_int, _vp, _vp, _vp, _sz, _vp]),
    "isc_senti_workspace_bytes": (_sz, [C.c_int, C.c_int]),
    "isc_senti_detect": (C.c_int, [C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int, _vp, C.c_int, _f32, C.c_int,
                                   _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "isc_prologue_bf16in": (C.c_int, [C.POINTER(Dims), _vp, C.c_int, _vp, _vp, _vp, C.c_int, _vp, _vp, C.c_int,
                                      C.POINTER(Feats), _vp, _sz, _vp]),
    "isc_shard_write": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.c_int, _i64, C.POINTER(C.c_char_p), _vp, _vp]),
    "isc_shard_open": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "isc_shard_close": (C.c_int, [_vp]),
    "isc_shard_info": (C.c_int, [_vp, C.POINTER(C.c_int64), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "isc_shard_find": (C.c_int64, [_vp, C.c_char_p]),
    "isc_shard_name": (C.c_char_p, [_vp, _i64]),
    "isc_shard_gather": (C.c_int, [_vp, _vp, _i64, _vp, _vp, C.c_int]),
    "isc_shard_pin": (C.c_int, [_vp]),
    "isc_shard_copy_to_device": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _vp]),
    "isc_sentcls_packed_bytes": (_sz, [C.c_int, C.c_int]),
    "isc_sentcls_pack": (C.c_int, [C.c_int, C.c_int] + [_vp] * 14 + [_sz, _vp]),
    "isc_sentcls_workspace_bytes": (_sz, [C.c_int, C.c_int]),
    "isc_sentcls_forward": (C.c_int, [C.c_int, C.c_int, _vp, _vp, _i64, _vp, C.c_int, C.c_int, _vp, _vp, _vp, _sz, _vp]),
    "isc_gemm_workspace_bytes": (_sz, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "isc_gemm_tn": (C.c_int, [C.c_int, _vp, _i64, _vp, _i64, _vp, _vp, _i64, C.c_int, C.c_int, C.c_int, C.c_int,
                              _vp, _sz, _vp]),
    "isc_cider_table_bytes": (_sz, [_i64]),
    "isc_cider_build_df": (C.c_int, [_vp, _vp, _i32, _vp, _i32, _vp, _i64, _vp, _vp]),
    "isc_cider_score": (C.c_int, [_vp, _i64, _dbl, _vp, _i32, _vp, _i32, _vp, _vp, _i32, _vp, _i32, _i32, _vp, _vp]),
    "isc_self_critical_reward": (C.c_int, [_vp, _i32, _i32, _vp, _vp]),
    "isc_cider_ngram_counts": (C.c_int, [_vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp]),
    "isc_launch_count": (_u64, []),
    "isc_profile_enable": (C.c_int, [C.c_int]),
    "isc_profile_reset": (C.c_int, []),
    "isc_profile_read": (C.c_int, [C.c_int, C.POINTER(_dbl), C.POINTER(_dbl), C.POINTER(_i64)]),
}

KERNEL_CLASSES = ["gemm_tc", "gemm_simt", "attention", "lstm", "pointwise", "select", "cider", "train"]

_lib = None


def load():
    """dlopen the in-tree library (RTLD_NOW) and attach signatures. Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("ISC_B200_LIB", LIB_PATH)  # another build of the same library, for A/B measurements
    if not os.path.exists(path):
        raise RuntimeError(
            "libisc_b200.so is not built: run `python -m insenticap_model_b200.build` "
            "(or __graft_entry__.build()). There is no CPU / PyTorch fallback for this path.")
    lib = C.CDLL(path, mode=os.RTLD_NOW)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code: int, what: str = "libisc_b200"):
    if code != 0:
        msg = load().isc_last_error().decode("utf-8", "replace")
        raise RuntimeError("%s failed with code %d: %s" % (what, code, msg))


def ptr(t):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("libisc_b200 takes CUDA tensors only (got a %s tensor); there is no CPU path" % t.device)
    if not t.is_contiguous():
        raise RuntimeError("libisc_b200 needs contiguous tensors")
    return C.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def to_device_async(values, dtype, device):
    """Small host data (a python list of lengths / indices, or a CPU tensor) -> device tensor WITHOUT a host
    synchronisation: staged in page-locked memory (torch's caching pinned allocator keeps the block alive until the copy
    has run) and copied with non_blocking=True. ``torch.tensor(list, device=...)`` copies from pageable memory, which
    makes torch synchronise the stream — one full pipeline drain per call."""
    import torch
    t = values if torch.is_tensor(values) else torch.tensor(values, dtype=dtype)
    if t.is_cuda or torch.device(device).type != "cuda":
        return t.to(device=device, dtype=dtype)
    return t.to(dtype).pin_memory().to(device, non_blocking=True)
