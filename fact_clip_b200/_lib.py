"""ctypes binding of libfactk.so (include/factk.h).  No fallback: if the library is missing or a call
fails, a RuntimeError is raised -- the product never routes through a CPU or library path."""
import ctypes as C
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libfactk.so')

F32, BF16 = 0, 1
MAX_SRC = 6
_DT = {torch.float32: F32, torch.bfloat16: BF16}

vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float


class Src(C.Structure):
    _fields_ = [('A', vp), ('a_dtype', i32), ('lda', i32), ('a_slot', i32), ('K', i32), ('row_off', i32),
                ('pos_ld', i32), ('pos_d', i32), ('ldw', i32), ('w_dtype', i32), ('reserved_', i32), ('gather', vp), ('pos', vp), ('pos_idx', vp),
                ('W', vp), ('w_bstride', i64)]


class Gemm(C.Structure):
    _fields_ = [('B', i32), ('slot', i32), ('N', i32), ('nsrc', i32), ('len', vp), ('src', Src * MAX_SRC),
                ('bias', vp), ('bias_bstride', i64), ('alpha', f32), ('relu', i32), ('pre', vp), ('pre_idx', vp),
                ('pre_bstride', i64), ('pre_dtype', i32), ('ldpre', i32), ('res', vp),
                ('res_dtype', i32), ('ldres', i32), ('Y', vp), ('y_dtype', i32), ('ldy', i32)]


class HeadsMM(C.Structure):
    _fields_ = [('A', vp), ('a_dtype', i32), ('lda', i32), ('a_bstride', i64), ('a_hstride', i32), ('a_kmajor', i32),
                ('Bm', vp), ('b_dtype', i32), ('ldb', i32), ('b_bstride', i64), ('b_hstride', i32), ('b_kmajor', i32),
                ('C', vp), ('c_dtype', i32), ('ldc', i32), ('c_bstride', i64), ('c_hstride', i32), ('accumulate', i32),
                ('M', i32), ('N', i32), ('K', i32), ('batch', i32), ('nhead', i32), ('len_mode', i32),
                ('len', vp), ('ws', vp), ('alpha', f32), ('reserved_', i32)]


class TokenLayer(C.Structure):
    _fields_ = [('x', vp), ('B', i32), ('M', i32), ('A', i32), ('nhead', i32), ('ff', i32), ('eps', f32),
                ('w_in', vp), ('b_in', vp), ('pre_qk', vp), ('o_in', vp), ('w_o', vp), ('b_o', vp), ('ln1_w', vp), ('ln1_b', vp),
                ('w_q', vp), ('b_q', vp), ('pre_q', vp), ('cq_out', vp), ('w_1', vp), ('b_1', vp), ('w_2', vp), ('b_2', vp),
                ('ln2_w', vp), ('ln2_b', vp)]


_SIGS = {
    'factk_version': (i32, []),
    'factk_last_error': (C.c_char_p, []),
    'factk_device_check': (i32, []),
    'factk_gemm': (i32, [C.POINTER(Gemm), vp]),
    'factk_heads_mm_ws_floats': (C.c_size_t, [i32, i32, i32, i32, i32]),
    'factk_heads_mm': (i32, [C.POINTER(HeadsMM), vp]),
    'factk_token_layer_supported': (i32, [i32, i32, i32, i32]),
    'factk_token_layer': (i32, [C.POINTER(TokenLayer), vp]),
    'factk_token_layer_debug': (i32, [vp]),
    'factk_gru_bwd_debug': (i32, [vp]),
    'factk_gemm_tc': (i32, [C.POINTER(Gemm), vp]),
    'factk_gemm_tc_supported': (i32, [C.POINTER(Gemm)]),
    'factk_gemm_pair_supported': (i32, [i32, i32]),
    'factk_gemm_pair': (i32, [vp, i32, i32, vp, i32, i32, i32, vp, vp, i32, i64, vp, i32, vp, i32, i32, i32, vp, vp]),
    'factk_tcn_layer_supported': (i32, [i32]),
    'factk_tcn_layer': (i32, [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, vp, i32, vp]),
    'factk_tcn_layer_dbg': (i32, [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, vp, i32, vp, vp]),
    'factk_softmax_splice': (i32, [vp, i32, i32, i32, vp, i32, i32, i32, vp, vp, vp]),
    'factk_layernorm': (i32, [vp, i32, i32, vp, i32, i32, vp, vp, f32, i32, vp, i32, i32, i32, i32, vp, i32, vp]),
    'factk_l2norm': (i32, [vp, i32, i32, vp, i32, i32, i32, i32, vp, i32, f32, vp]),
    'factk_row_softmax': (i32, [vp, i32, vp, i32, i32, i32, vp, i32, f32, vp, i32, i32, vp]),
    'factk_mha_tokens': (i32, [vp, vp, vp, i32, vp, i32, i32, i32, i32, i32, i32, vp]),
    'factk_attn_rows_ws_floats': (C.c_size_t, [i32, i32, i32, i32, i32]),
    'factk_attn_rows': (i32, [vp, i32, vp, vp, i32, i32, vp, i32, i32, i32, vp, i32, i32, i32, vp, vp]),
    'factk_attn_tc_debug': (i32, [vp]),
    'factk_col_softmax_ws_floats': (C.c_size_t, [i32, i32, i32, i32]),
    'factk_col_softmax_apply': (i32, [vp, i32, vp, i32, i32, vp, i32, vp, i32, i32, i32, vp, i32, i32, vp, vp]),
    'factk_a2f_fused_supported': (i32, [i32, i32, i32, i32]),
    'factk_a2f_fused': (i32, [vp, i32, vp, i32, C.c_longlong, vp, C.c_longlong, vp, i32, vp, i32, C.c_longlong, vp, vp, i32, vp, vp, i32, i32, i32,
                              vp, i32, i32, i32, vp]),
    'factk_f2a_fused_supported': (i32, [i32, i32, i32]),
    'factk_f2a_fused_ws_floats': (C.c_size_t, [i32, i32, i32, i32]),
    'factk_f2a_debug': (i32, [vp]),
    'factk_f2a_fused': (i32, [vp, i32, vp, i32, C.c_longlong, vp, i32, i32, i32, vp, i32, i32, vp, vp]),
    'factk_tdu_segment': (i32, [vp, i32, i32, vp, vp, vp, vp, vp, vp, vp]),
    'factk_segment_mean_ws_floats': (C.c_size_t, [i32, i32, i32]),
    'factk_segment_mean': (i32, [vp, i32, i32, vp, i32, i32, vp, vp, vp, vp, i32, i32, i32, vp, vp]),
    'factk_gru_bidir': (i32, [vp, vp, vp, vp, vp, i32, vp, i32, i32, i32, i32, i32, vp, vp]),
    'factk_gru_bidir_mma': (i32, [vp, vp, vp, vp, vp, i32, vp, i32, i32, i32, i32, i32, vp, vp]),
    'factk_gru_bidir_mma_sorted': (i32, [vp, vp, vp, vp, vp, i32, vp, i32, i32, i32, i32, i32, vp, vp, vp]),
    'factk_gru_max_clusters': (i32, []),
    'factk_gru_bidir_mma_dbg': (i32, [vp, vp, vp, vp, vp, i32, vp, i32, i32, i32, i32, i32, vp, vp, vp]),
    'factk_gather_rows': (i32, [vp, i32, i32, vp, vp, i32, i32, i32, vp, i32, vp]),
    'factk_fuse_eval_transcript': (i32, [vp, i32, i32, vp, vp, i32, f32, vp, i32, vp, vp, i32, i32, vp, i32, vp]),
    'factk_embed_tokens': (i32, [vp, i32, vp, vp, i32, vp, i32, i32, i32, vp]),
    'factk_label_prep': (i32, [vp, vp, vp, vp, vp, i32, vp, vp, i32, vp, vp, i32, i32, vp, vp]),
    'factk_match_cost': (i32, [vp, i32, i32, vp, vp, i32, i32, vp, vp, vp, vp, i32, f32, f32, vp, i32, vp, i32, i32, i32, vp]),
    'factk_loss_pick': (i32, [vp, i32, i32, i32, vp, i32, vp, vp, vp, vp, i32, vp, vp, i32, vp, i32, vp, vp, i32, i32, vp, i32, vp]),
    'factk_loss_smooth': (i32, [vp, i32, i32, vp, i32, i32, vp, i32, i32, vp]),
    'factk_col_lse_ws_floats': (C.c_size_t, [i32, i32, i32]),
    'factk_col_lse': (i32, [vp, i32, i32, i32, vp, vp, vp, i32, vp, i32, i32, vp, vp]),
    'factk_token_loss': (i32, [vp, i32, i32, vp, vp, vp, i32, vp, i32, vp, vp, i32, i32, i32, vp]),
    'factk_loss_combine': (i32, [vp, i32, vp, i32, i32, vp, vp, i32, i32, f32, i32, f32, f32, i32, vp, vp, i32, vp]),
    'factk_fuse_eval': (i32, [vp, vp, i32, i32, vp, vp, i32, f32, vp, i32, i32, vp, i32, i32, i32, vp]),
    'factk_transpose_rows': (i32, [vp, C.c_longlong, vp, i32, i32, i32, i32, i32, vp, vp]),
    'factk_vn_splice': (i32, [vp, i32, i32, i32, vp, i32, i32, i32, i32, vp, vp, vp, i32, vp, vp]),
    'factk_vn_combine': (i32, [vp, i32, i32, i32, vp, vp, i32, i32, vp, i32, i32, i32, vp, vp]),
    # training step (csrc/train*.cu)
    'factk_wgrad_ws_floats': (C.c_size_t, [i32, i32, i32, i32]),
    'factk_wgrad': (i32, [vp, i32, i32, vp, i32, i32, i32, i32, vp, i32, i32, vp, i32, i32, vp, i32, C.c_longlong, f32, i32, i32, i32,
                          vp, vp, vp]),
    'factk_wgrad_tc_supported': (i32, [i32, i32, i32, i32, i32, i32, i32]),
    'factk_wgrad_tc_ws_floats': (C.c_size_t, [i32, i32, i32, i32]),
    'factk_wgrad_tc': (i32, [vp, i32, vp, i32, i32, i32, i32, i32, vp, i32, C.c_longlong, f32, i32, i32, i32, vp, vp, vp]),
    'factk_colsum_ws_floats': (C.c_size_t, [i32, i32, i32]),
    'factk_colsum': (i32, [vp, i32, i32, vp, i32, i32, i32, vp, C.c_longlong, f32, i32, i32, i32, vp, vp, vp]),
    'factk_rows_elementwise': (i32, [i32, vp, i32, i32, vp, i32, i32, vp, i32, i32, i32, i32, i32, vp, f32, f32, C.c_ulonglong,
                                     C.c_uint, i32, vp, vp, vp]),
    'factk_transpose': (i32, [vp, i32, C.c_longlong, vp, i32, C.c_longlong, i32, i32, i32, vp]),
    'factk_splice_bwd': (i32, [vp, i32, i32, vp, i32, i32, vp, i32, vp, i32, i32, i32, i32, i32, i32, vp, vp]),
    'factk_row_softmax_bwd': (i32, [vp, i32, vp, i32, vp, i32, i32, i32, i32, i32, vp, vp]),
    'factk_layernorm_bwd_ws_floats': (C.c_size_t, [i32, i32, i32]),
    'factk_layernorm_bwd': (i32, [vp, i32, i32, vp, i32, i32, vp, vp, f32, i32, vp, i32, i32, vp, i32, i32, i32, vp, vp, i32, i32, vp,
                                  i32, vp, vp]),
    'factk_l2norm_bwd': (i32, [vp, i32, i32, vp, i32, i32, vp, i32, i32, i32, i32, vp, i32, f32, vp]),
    'factk_col_softmax_train_ws_floats': (C.c_size_t, [i32, i32, i32]),
    'factk_col_softmax': (i32, [vp, i32, i32, vp, i32, i32, i32, f32, i32, i32, vp, vp, vp]),
    'factk_col_softmax_bwd': (i32, [vp, i32, i32, vp, i32, i32, vp, i32, i32, i32, f32, i32, i32, i32, vp, vp, vp]),
    'factk_segment_reduce': (i32, [vp, i32, i32, vp, i32, i32, vp, vp, vp, i32, i32, i32, i32, i32, vp]),
    'factk_segment_expand': (i32, [vp, i32, i32, vp, vp, vp, i32, i32, i32, i32, vp, i32, i32, i32, vp]),
    'factk_gru_bwd': (i32, [vp, vp, vp, i32, i32, vp, i32, i32, vp, vp, i32, vp, vp, i32, i32, vp, vp]),
    'factk_loss_grad_ce_rows': (i32, [vp, i32, i32, vp, i32, vp, vp, vp, vp, vp, vp, i32, i32, vp]),
    'factk_loss_grad_smooth': (i32, [vp, i32, i32, vp, i32, vp, vp, f32, i32, i32, vp]),
    'factk_loss_grad_token': (i32, [vp, i32, i32, vp, vp, vp, vp, i32, vp, i32, vp, vp, i32, vp]),
    'factk_loss_grad_xattn': (i32, [i32, vp, i32, i32, vp, i32, vp, vp, vp, vp, vp, vp, i32, vp, vp, vp, i32, vp, vp, vp, vp, vp, i32,
                                    i32, vp]),
    'factk_loss_grad_infonce': (i32, [vp, i32, i32, vp, i32, vp, vp, vp, vp, i32, vp, i32, vp, vp, i32, i32, vp]),
}

_lib = None


def exported_symbols():
    return sorted(_SIGS)


def load():
    """Load libfactk.so once; raise loudly if it is absent (build with ``python -m fact_clip_b200.build``)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f'{LIB_PATH} not found: build it with `python -m fact_clip_b200.build` '
                               '(there is no CPU / PyTorch fallback for the FACT forward)')
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(rc, what=''):
    if rc != 0:
        raise RuntimeError(f'factk {what} failed ({rc}): {load().factk_last_error().decode()}')


def dt(t):
    return _DT[t.dtype]


def ptr(t):
    return None if t is None else t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def call(name, *args):
    check(getattr(load(), name)(*args), name)
