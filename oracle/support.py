"""TEST / BASELINE INFRASTRUCTURE: ctypes mirrors of the structs taken by oracle/ref_driver.c and the
numpy input preparation (padded HME pyramid) shared by tests/ and bench.py's reference arm."""
import ctypes as ct

import numpy as np


class RefMePicture(ct.Structure):
    _fields_ = [("plane", ct.c_void_p * 3), ("stride", ct.c_int32 * 3), ("org_x", ct.c_int32 * 3), ("org_y", ct.c_int32 * 3),
                ("width", ct.c_int32 * 3), ("height", ct.c_int32 * 3), ("reserved", ct.c_int32 * 2)]


class RefMeParams(ct.Structure):
    _fields_ = [(n, ct.c_int32) for n in ("hme_l0_sa_w", "hme_l0_sa_h", "hme_l1_sa_w", "hme_l1_sa_h", "hme_l2_sa_w", "hme_l2_sa_h", "me_sa_w",
                                          "me_sa_h", "hme_sub_sad", "me_sub_sad", "check_zero_centre", "reserved")]


class RefCdefFrame(ct.Structure):
    _fields_ = [("recon_y", ct.c_void_p), ("recon_cb", ct.c_void_p), ("recon_cr", ct.c_void_p), ("src_y", ct.c_void_p),
                ("src_cb", ct.c_void_p), ("src_cr", ct.c_void_p), ("recon_stride_y", ct.c_int32), ("recon_stride_c", ct.c_int32),
                ("src_stride_y", ct.c_int32), ("src_stride_c", ct.c_int32), ("width", ct.c_int32), ("height", ct.c_int32),
                ("bit_depth", ct.c_int32), ("damping", ct.c_int32), ("subsampling_factor", ct.c_int32), ("reserved", ct.c_int32)]


def pad_np(buf, pad, w, h):
    inner = buf[pad:pad + h, pad:pad + w].copy()
    buf[:, :] = np.pad(inner, ((pad, buf.shape[0] - pad - h), (pad, buf.shape[1] - pad - w)), mode="edge")
    return buf


def build_pyramid_np(full, width, height, shapes):
    """numpy restatement of svt_aom_downsample_2d_c + svt_aom_generate_padding
    (pic_analysis_process.c:130-160, 2138-2190), checked against the reference in tests."""
    planes = [None, None, None]
    th, stride, pad, w, h = shapes[2]
    buf = np.zeros((th, stride), np.uint8)
    buf[pad:pad + h, pad:pad + w] = full
    planes[2] = pad_np(buf, pad, w, h)
    src = full.astype(np.uint32)
    for lvl in (1, 0):
        th, stride, pad, w, h = shapes[lvl]
        d = ((src[0::2, 0::2] + src[0::2, 1::2] + src[1::2, 0::2] + src[1::2, 1::2] + 2) >> 2)[:h, :w]
        buf = np.zeros((th, stride), np.uint8)
        buf[pad:pad + h, pad:pad + w] = d
        planes[lvl] = pad_np(buf, pad, w, h)
        src = d.astype(np.uint32)
    return planes


def ref_pic_desc(planes, shapes):
    p = RefMePicture()
    for lvl, (th, stride, pad, w, h) in enumerate(shapes):
        p.plane[lvl] = planes[lvl].ctypes.data
        p.stride[lvl] = stride
        p.org_x[lvl] = p.org_y[lvl] = pad
        p.width[lvl] = w
        p.height[lvl] = h
    return p
