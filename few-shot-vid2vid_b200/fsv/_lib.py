"""ctypes binding of libfsv_b200.so (the C ABI declared in include/fsv_b200.h).

There is NO fallback: if the shared library is missing or a symbol is absent the import
fails loudly -- the product path never routes through PyTorch/CPU substitutes.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libfsv_b200.so')

c_int, c_ll, c_float, c_double, c_vp = ctypes.c_int, ctypes.c_longlong, ctypes.c_float, ctypes.c_double, ctypes.c_void_p

ACT_NONE, ACT_LRELU, ACT_TANH, ACT_SIGMOID, ACT_RELU = 0, 1, 2, 3, 4
NORM_BATCH, NORM_INSTANCE = 0, 1
SPADE_MAX_MAPS = 3


class ConvDesc(ctypes.Structure):
    _fields_ = [('N', c_int), ('H', c_int), ('W', c_int), ('Cin', c_int), ('x_ld', c_int), ('x_coff', c_int), ('up', c_int),
                ('Cout', c_int), ('kh', c_int), ('kw', c_int), ('stride', c_int), ('pad', c_int), ('Ho', c_int), ('Wo', c_int),
                ('y_ld', c_int), ('y_coff', c_int), ('act', c_int), ('out_scale', c_float), ('w_nstride', c_ll),
                ('b_nstride', c_ll), ('res_ld', c_int), ('res_coff', c_int), ('in_act', c_int), ('use_tc', c_int)]


class SpadeDesc(ctypes.Structure):
    _fields_ = [('N', c_int), ('H', c_int), ('W', c_int), ('C', c_int), ('up', c_int), ('mode', c_int), ('act', c_int),
                ('nmaps', c_int), ('K', c_int * SPADE_MAX_MAPS), ('m_ld', c_int * SPADE_MAX_MAPS),
                ('m_coff', c_int * SPADE_MAX_MAPS), ('w_nstride', c_ll * SPADE_MAX_MAPS), ('dgb_ld', c_int * SPADE_MAX_MAPS)]


class SnItem(ctypes.Structure):
    _fields_ = [('w_orig', c_vp), ('u', c_vp), ('v', c_vp), ('R', c_int), ('Cin', c_int), ('taps', c_int), ('want_wt', c_int),
                ('K', c_int), ('nchunks', c_int), ('rs', c_int), ('rps', c_int), ('nblk2', c_int), ('nblk3', c_int), ('ticket_off', c_int),
                ('blk1', c_int), ('blk2', c_int), ('blk3', c_int),
                ('out_off', c_ll), ('wt_off', c_ll), ('uvs_off', c_ll), ('work_off', c_ll), ('nb1', c_int), ('nb2', c_int), ('bwd_work_off', c_ll)]


class FlowMaskDesc(ctypes.Structure):
    _fields_ = [(n, c_vp) for n in ('warp0', 'mask0', 'warp1', 'mask1', 'tgt', 'fake', 'ref_body_warp', 'body', 'ref_fg_warp', 'fg', 'face_avg',
                                     'fg_diff')] + [('B', c_int), ('H', c_int), ('W', c_int)]


class AdamItem(ctypes.Structure):
    _fields_ = [('param', c_vp), ('grad', c_vp), ('exp_avg', c_vp), ('exp_avg_sq', c_vp), ('numel', c_ll)]


PtrArray = c_vp * SPADE_MAX_MAPS
_CD, _SD = ctypes.POINTER(ConvDesc), ctypes.POINTER(SpadeDesc)

# name -> argtypes; every symbol include/fsv_b200.h declares must be listed here
SIGNATURES = {
    'fsv_version': [],
    'fsv_device_info': [ctypes.POINTER(c_int)] * 3,
    'fsv_nchw_to_nhwc': [c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_vp],
    'fsv_nhwc_to_nchw': [c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_vp],
    'fsv_copy_channels': [c_vp, c_int, c_int, c_vp, c_int, c_int, c_ll, c_int, c_int, c_vp],
    'fsv_upsample2x_fwd': [c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp],
    'fsv_upsample2x_bwd': [c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp],
    'fsv_maxpool2_fwd': [c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp],
    'fsv_maxpool2_bwd': [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp],
    'fsv_avgpool3s2_fwd': [c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp],
    'fsv_avgpool3s2_bwd': [c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp],
    'fsv_act_bwd': [c_vp, c_vp, c_vp, c_ll, c_int, c_float, c_vp],
    'fsv_conv2d_fwd': [_CD, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp],
    'fsv_conv2d_dgrad': [_CD, c_vp, c_vp, c_vp, c_int, c_vp],
    'fsv_conv2d_wgrad': [_CD, c_vp, c_vp, c_vp, c_vp, c_int, c_vp],
    'fsv_conv2d_tc_eligible': [_CD],
    'fsv_conv2d_fwd_tc_up2_eligible': [_CD],
    'fsv_up2_weights': [c_vp, c_vp, c_int, c_int, c_vp],
    'fsv_conv2d_fwd_tc_up2': [_CD, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp],
    'fsv_up2_dgrad_weights': [c_vp, c_vp, c_int, c_int, c_vp],
    'fsv_up2_wgrad_fold': [c_vp, c_vp, c_int, c_int, c_int, c_vp],
    'fsv_conv2d_dgrad_tc_eligible': [_CD],
    'fsv_conv2d_dgrad_tc': [_CD, c_vp, c_vp, c_vp, c_vp],
    'fsv_conv2d_wgrad_tc_eligible': [_CD],
    'fsv_conv2d_wgrad_tc_workspace': [_CD],
    'fsv_conv2d_wgrad_tc': [_CD, c_vp, c_vp, c_vp, c_vp, c_int, c_vp],
    'fsv_norm_work_doubles': [c_int, c_int, c_ll],
    'fsv_set_reduction_lane': [c_int],
    'fsv_norm_stats': [c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp],
    'fsv_norm_stats_finalize': [c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_double, c_float, c_float, c_vp, c_vp, c_int, c_vp, c_vp, c_vp, c_vp],
    'fsv_norm_finalize': [c_vp, c_vp, c_int, c_int, c_double, c_double, c_float, c_float, c_vp, c_vp, c_int, c_vp, c_vp, c_vp],
    'fsv_norm_from_running': [c_vp, c_vp, c_int, c_float, c_vp, c_vp, c_vp],
    'fsv_norm_apply_fwd': [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_vp],
    'fsv_norm_apply_bwd': [c_vp] * 10 + [c_int] * 6 + [c_vp],
    'fsv_norm_apply_bwd2': [c_vp] * 10 + [c_int] * 6 + [c_vp],
    'fsv_spade_fwd': [_SD, c_vp, c_vp, c_vp, PtrArray, PtrArray, PtrArray, PtrArray, PtrArray, c_vp, c_vp],
    'fsv_spade_fwd_tc_eligible': [_SD],
    'fsv_spade_fwd_tc': [_SD, c_vp, c_vp, c_vp, PtrArray, PtrArray, PtrArray, PtrArray, PtrArray, c_vp, c_vp],
    'fsv_spade_bwd_tc': [_SD, c_vp, c_vp, c_vp, PtrArray, PtrArray, PtrArray, PtrArray, PtrArray, c_vp, c_vp, PtrArray, PtrArray, c_vp],
    'fsv_spade_bwd': [_SD, c_vp, c_vp, c_vp, PtrArray, PtrArray, PtrArray, PtrArray, PtrArray, c_vp, c_vp, PtrArray, PtrArray, c_vp],
    'fsv_spade_norm_bwd': [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_vp],
    'fsv_warp_fwd': [c_vp] * 5 + [c_int] * 7 + [c_vp],
    'fsv_warp_bwd': [c_vp] * 9 + [c_int] * 7 + [c_vp],
    'fsv_softmax_rows_fwd': [c_vp, c_vp, c_ll, c_int, c_vp],
    'fsv_softmax_rows_bwd': [c_vp, c_vp, c_vp, c_ll, c_int, c_vp],
    'fsv_spectral_workspace': [c_int, c_int],
    'fsv_spectral_fwd': [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_float, c_vp, c_vp, c_vp, c_vp, c_vp],
    'fsv_spectral_bwd': [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_vp, c_vp, c_vp],
    'fsv_spectral_group_plan': [c_vp, c_int, c_vp],
    'fsv_spectral_group_fwd': [c_vp, c_vp, c_vp, c_int, c_float, c_int, c_vp, c_vp, c_vp, c_vp],
    'fsv_spectral_group_bwd': [c_vp] * 9,
    'fsv_adam_chunks': [c_ll],
    'fsv_adam_step': [c_vp, c_vp, c_ll, c_vp, c_float, c_float, c_float, c_float, c_vp],
    'fsv_flow_mask_loss_work_doubles': [],
    'fsv_flow_mask_loss_fwd': [c_vp, c_vp, c_vp, c_vp],
    'fsv_flow_mask_loss_bwd': [c_vp] * 10 + [c_vp],
    'fsv_halves_l1_fwd': [c_vp, c_ll, c_vp, c_vp, c_vp],
    'fsv_halves_l1_bwd': [c_vp, c_ll, c_vp, c_vp, c_vp],
    'fsv_fg_mask': [c_vp, c_ll, c_vp, c_int, c_int, c_int, c_float, c_vp],
    'fsv_face_mask_avg15': [c_vp, c_ll, c_vp, c_int, c_int, c_int, c_vp],
    'fsv_part_masks': [c_vp, c_ll, c_vp, c_int, c_int, c_int, c_vp],
    'fsv_face_bbox': [c_vp, c_vp, c_vp, c_ll, c_ll, c_ll, c_float, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp],
    'fsv_crop_resize_fwd': [c_vp, c_ll, c_ll, c_ll, c_ll, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_vp],
    'fsv_crop_resize_bwd': [c_vp, c_int, c_int, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_vp],
}


def _load():
    if not os.path.exists(LIB_PATH):
        raise RuntimeError('fsv_b200: %s is missing -- build it with `python few-shot-vid2vid_b200/build.py` '
                           '(or __graft_entry__.build()); there is no CPU / PyTorch fallback' % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.argtypes = argtypes
        fn.restype = c_int
    lib.fsv_conv2d_wgrad_tc_workspace.restype = c_ll
    lib.fsv_spectral_workspace.restype = c_ll
    lib.fsv_norm_work_doubles.restype = c_ll
    lib.fsv_adam_chunks.restype = c_ll
    lib.fsv_flow_mask_loss_work_doubles.restype = c_ll
    lib.fsv_last_error.argtypes = []
    lib.fsv_last_error.restype = ctypes.c_char_p
    return lib


lib = _load()


class FsvError(RuntimeError):
    pass


def check(rc, what=''):
    if rc != 0:
        raise FsvError('%s failed (code %d): %s' % (what, rc, lib.fsv_last_error().decode()))


def ptr(t):
    return None if t is None else c_vp(t.data_ptr())


def stream():
    return c_vp(torch.cuda.current_stream().cuda_stream)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise FsvError('fsv_b200 ops run on CUDA tensors only (got a %s tensor); there is no CPU fallback' % t.device)
