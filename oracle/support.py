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


# ---- oracle/ref_me_b64.c: the reference's own svt_aom_motion_estimation_b64 over a picture ------------------
class RefMeB64Cfg(ct.Structure):
    _fields_ = [(n, ct.c_int32) for n in ("enc_mode", "qp", "hierarchical_levels", "temporal_layer_index", "is_ref", "sc_class1")] + \
               [("n_ref", ct.c_int32 * 2), ("ref_poc_dist_sign", (ct.c_int32 * 4) * 2)] + \
               [(n, ct.c_int32) for n in ("enable_hme_flag", "enable_hme_level0_flag", "enable_hme_level1_flag", "enable_hme_level2_flag",
                                          "max_l0", "max_l1", "only_l_bwd", "safe_limit_nref", "safe_limit_zz_th", "similar_brightness_refs",
                                          "gm_enabled", "gm_use_distance_based_active_th", "frame_rate_q16")] + [("reserved", ct.c_int32 * 3)]


# field order == SvtB200MeControls (include/svt_b200.h) == RefMeControls (oracle/ref_me_b64.c)
ME_CONTROL_FIELDS = [("n_list", 1), ("n_ref", 2), ("temporal_layer_index", 1), ("is_ref", 1), ("hierarchical_levels", 1), ("dist", 8),
                     ("enable_hme", 1), ("enable_l0", 1), ("enable_l1", 1), ("enable_l2", 1), ("hme_sub_sad", 1), ("me_sub_sad", 1),
                     ("hme_l0_min_w", 1), ("hme_l0_min_h", 1), ("hme_l0_max_w", 1), ("hme_l0_max_h", 1), ("hme_l1_w", 1), ("hme_l1_h", 1),
                     ("hme_l2_w", 1), ("hme_l2_h", 1), ("me_min_w", 1), ("me_min_h", 1), ("me_max_w", 1), ("me_max_h", 1),
                     ("prehme_enable", 1), ("prehme_sa", 8), ("prehme_skip_search_line", 1), ("prehme_l1_early_exit", 1),
                     ("prune_enable", 1), ("prune_hme_th", 1), ("prune_me_th", 1), ("zz_sad_th", 1), ("zz_sad_pct", 1), ("phme_sad_th", 1),
                     ("phme_sad_pct", 1), ("sr_enable", 1), ("sr_mv_length_th", 1), ("sr_stationary_hme_sad_abs_th", 1),
                     ("sr_stationary_divisor", 1), ("sr_hme_sad_abs_th", 1), ("sr_low_hme_sad_divisor", 1), ("sr_distance_based_hme_resizing", 1),
                     ("var_enable", 1), ("var_div4_th", 1), ("var_div2_th", 1), ("var_mult2_th", 1),
                     ("mvsa_enable", 1), ("mvsa_nearest_ref_only", 1), ("mvsa_mv_size_th", 1), ("mvsa_multiplier", 1),
                     ("reduce_hme_l0_sr_th_min", 1), ("reduce_hme_l0_sr_th_max", 1),
                     ("me_early_exit_th", 1), ("me_safe_limit_zz_th", 1), ("prev_me_stage_based_exit_th", 1), ("prune_me_candidates_th", 1),
                     ("use_best_unipred_cand_only", 1), ("similar_brightness_refs", 1), ("only_l_bwd", 1), ("enable_me_8x8", 1),
                     ("enable_me_16x16", 1), ("max_cand", 1), ("max_refs", 1), ("max_l0", 1), ("gm_enabled", 1),
                     ("gm_use_distance_based_active_th", 1), ("resolution_le_480p", 1), ("reserved", 5)]


class RefMeControls(ct.Structure):
    _fields_ = [(n, ct.c_int32 if k == 1 else ct.c_int32 * k) for n, k in ME_CONTROL_FIELDS]

    def as_dict(self):
        return {n: (int(getattr(self, n)) if k == 1 else [int(v) for v in getattr(self, n)]) for n, k in ME_CONTROL_FIELDS if n != "reserved"}


class RefMeB64Out(ct.Structure):
    _fields_ = [(n, ct.c_void_p) for n in ("total_me_candidate_index", "me_candidate_array", "me_mv_array", "distortion", "flags", "do_ref",
                                           "hme_centre", "zz_sad", "best_sad", "best_mv")]


def me_b64_cfg(preset=8, qp=30, n_ref=(2, 2), poc_dist=((-1, -3, 0, 0), (1, 3, 0, 0)), temporal_layer_index=3, hierarchical_levels=4,
               is_ref=0, max_l=(3, 2), only_l_bwd=1, safe_limit_nref=2, gm_enabled=0):
    """defaults: a non-base, non-reference B picture of the preset-8 CRF random-access configuration (MRP level 10,
    enc_handle.c:3559-3577: 2+2 references tried on non-base pictures, storage sized for 3+2; 5-layer hierarchy)"""
    c = RefMeB64Cfg()
    c.enc_mode, c.qp, c.hierarchical_levels, c.temporal_layer_index, c.is_ref, c.sc_class1 = preset, qp, hierarchical_levels, temporal_layer_index, is_ref, 0
    c.n_ref[0], c.n_ref[1] = n_ref
    for li in range(2):
        for r in range(4):
            c.ref_poc_dist_sign[li][r] = poc_dist[li][r]
    c.enable_hme_flag = c.enable_hme_level0_flag = c.enable_hme_level1_flag = c.enable_hme_level2_flag = 1
    c.max_l0, c.max_l1 = max_l
    c.only_l_bwd, c.safe_limit_nref, c.safe_limit_zz_th, c.similar_brightness_refs = only_l_bwd, safe_limit_nref, 0, 0
    c.gm_enabled, c.gm_use_distance_based_active_th, c.frame_rate_q16 = gm_enabled, 0, 30 << 16
    return c


def ref_me_b64_picture(ref, cur_pyr, ref_pyrs, shapes, cfg, run=True):
    """-> (controls dict, outputs dict of numpy arrays or None)"""
    W, H = shapes[2][3], shapes[2][4]
    nb = ((W + 63) // 64) * ((H + 63) // 64)
    ref.ref_me_b64_num_pus.restype = ct.c_int
    n_pu = ref.ref_me_b64_num_pus(cfg.enc_mode, W, H)
    cd = ref_pic_desc(cur_pyr, shapes)
    rd = (RefMePicture * len(ref_pyrs))(*[ref_pic_desc(p, shapes) for p in ref_pyrs])
    ctrl = RefMeControls()
    ref.ref_me_b64_picture.restype = ct.c_int
    ref.ref_me_b64_picture.argtypes = [ct.c_void_p] * 5
    if not run:
        ref.ref_me_b64_picture(ct.byref(cd), rd, ct.byref(cfg), ct.byref(ctrl), None)
        return ctrl, None
    ref.ref_me_b64_picture(ct.byref(cd), rd, ct.byref(cfg), ct.byref(ctrl), None)  # sizes first
    mc, mr = ctrl.max_cand, ctrl.max_refs
    arrs = {"total_me_candidate_index": np.zeros((nb, n_pu), np.uint8), "me_candidate_array": np.zeros((nb, n_pu * mc), np.uint8),
            "me_mv_array": np.zeros((nb, n_pu * mr), np.uint32), "distortion": np.zeros((nb, 6), np.uint32), "flags": np.zeros((nb, 2), np.uint8),
            "do_ref": np.zeros((nb, 2, 4), np.uint8), "hme_centre": np.zeros((nb, 2, 4, 2), np.int16), "zz_sad": np.zeros((nb, 2, 4), np.uint32),
            "best_sad": np.zeros((nb, 2, 4, 85), np.uint32), "best_mv": np.zeros((nb, 2, 4, 85), np.uint32)}
    out = RefMeB64Out()
    for n, _ in RefMeB64Out._fields_:
        setattr(out, n, arrs[n].ctypes.data)
    assert ref.ref_me_b64_picture(ct.byref(cd), rd, ct.byref(cfg), ct.byref(ctrl), ct.byref(out)) == 0
    return ctrl, arrs
