"""Pure-numpy description of the C ABI's data layout (include/svt_b200.h): struct dtypes of the T2 work
items, the reference's enumerator values and the plane geometry of the ME pyramid.  No library is loaded
here -- the module is shared by the product's host mirror (dsp.py), the synthetic workload (workload.py) and,
loaded stand-alone, by bench.py's CPU reference arm, which must not map libsvtav1_b200.so."""
import numpy as np

SAD_ITEM_DTYPE = np.dtype([("src_off", "<u8"), ("ref_off", "<u8"), ("src_stride", "<u4"), ("ref_stride", "<u4"),
                           ("ref_step", "<u4"), ("block_w", "<u2"), ("block_h", "<u2"), ("sa_w", "<i2"),
                           ("sa_h", "<i2"), ("skip_search_line", "<u2"), ("reserved", "<u2")])
SAD_RESULT_DTYPE = np.dtype([("best_sad", "<u4"), ("x", "<i2"), ("y", "<i2")])

TX_W = [4, 8, 16, 32, 64, 4, 8, 8, 16, 16, 32, 32, 64, 4, 16, 8, 32, 16, 64]
TX_H = [4, 8, 16, 32, 64, 8, 4, 16, 8, 32, 16, 64, 32, 16, 4, 32, 8, 64, 16]
TX_NAME = ["%dx%d" % (w, h) for w, h in zip(TX_W, TX_H)]
TXFM_CLASSES = 5  # SVT_B200_TXFM_CLASSES

def txfm_team_class(tx_size):
    """log2(max(W, H)) - 2: the order key of the transform batch calls (svt_b200_txfm_team_class)."""
    return {4: 0, 8: 1, 16: 2, 32: 3, 64: 4}[max(TX_W[tx_size], TX_H[tx_size])]

FWD_ITEM_DTYPE = np.dtype([("src_off", "<u8"), ("dst_off", "<u8"), ("src_stride", "<u4"), ("tx_size", "u1"),
                           ("tx_type", "u1"), ("reserved", "<u2")])
INV_ITEM_DTYPE = np.dtype([("coef_off", "<u8"), ("pred_off", "<u8"), ("recon_off", "<u8"), ("pred_stride", "<u4"),
                           ("recon_stride", "<u4"), ("tx_size", "u1"), ("tx_type", "u1"), ("bd", "u1"),
                           ("reserved", "u1"), ("reserved2", "<u4")])
assert FWD_ITEM_DTYPE.itemsize == 24 and INV_ITEM_DTYPE.itemsize == 40

QUANT_B_LBD, QUANT_B_HBD, QUANT_FP_LBD, QUANT_FP_HBD = 0, 1, 2, 3
NO_QM = 0xffffffff
QUANT_ITEM_DTYPE = np.dtype([("coeff_off", "<u8"), ("q_off", "<u8"), ("dq_off", "<u8"), ("scan_off", "<u4"),
                             ("qm_off", "<u4"), ("iqm_off", "<u4"), ("n_coeffs", "<u4"), ("zbin", "<i2", 2),
                             ("round", "<i2", 2), ("quant", "<i2", 2), ("quant_shift", "<i2", 2), ("dequant", "<i2", 2),
                             ("mode", "u1"), ("log_scale", "u1"), ("reserved", "<u2")])
assert QUANT_ITEM_DTYPE.itemsize == 64
TRIO_ITEM_DTYPE = np.dtype([("fwd", FWD_ITEM_DTYPE), ("quant", QUANT_ITEM_DTYPE), ("inv", INV_ITEM_DTYPE)])  # SvtB200TrioItem
assert TRIO_ITEM_DTYPE.itemsize == 128

HADAMARD_ITEM_DTYPE = np.dtype([("src_off", "<u8"), ("coeff_off", "<u8"), ("src_stride", "<u4"), ("size", "<u4")])
FULLPEL_ITEM_DTYPE = np.dtype([("src_off", "<u8"), ("ref_off", "<u8"), ("src_stride", "<u4"), ("ref_stride", "<u4"),
                               ("sa_w", "<i2"), ("sa_h", "<i2"), ("org_x", "<i2"), ("org_y", "<i2"), ("sub_sad", "u1"),
                               ("seeded", "u1"), ("seed_x", "<i2"), ("seed_y", "<i2"), ("reserved", "u1", 2)])
assert HADAMARD_ITEM_DTYPE.itemsize == 24 and FULLPEL_ITEM_DTYPE.itemsize == 40

WIENER_UNIT_DTYPE = np.dtype([("src_off", "<u8"), ("dst_off", "<u8"), ("src_stride", "<i4"), ("dst_stride", "<i4"), ("w", "<u2"),
                              ("h", "<u2"), ("reserved", "<u4"), ("hfilter", "<i2", 8), ("vfilter", "<i2", 8)])
STATS_ITEM_DTYPE = np.dtype([("dgd_off", "<u8"), ("src_off", "<u8"), ("dgd_stride", "<i4"), ("src_stride", "<i4"),
                             ("h_start", "<i4"), ("h_end", "<i4"), ("v_start", "<i4"), ("v_end", "<i4"), ("wiener_win", "<i4"),
                             ("reserved", "<i4")])
assert WIENER_UNIT_DTYPE.itemsize == 64 and STATS_ITEM_DTYPE.itemsize == 48

SGR_UNIT_DTYPE = np.dtype([("dgd_off", "<u8"), ("flt0_off", "<u8"), ("flt1_off", "<u8"), ("dgd_stride", "<i4"), ("flt_stride", "<i4"),
                           ("w", "<u2"), ("h", "<u2"), ("params_idx", "<u2"), ("reserved", "<u2")])
assert SGR_UNIT_DTYPE.itemsize == 40

ME_PAD = (16, 32, 72)  # padding of the 1/16, 1/4 and full luma planes (the reference uses 16 / 32 / 64+)


def me_plane_shapes(width, height):
    """[(h_total, w_total, org, w, h)] for levels 0 (1/16), 1 (1/4), 2 (full); strides are 16-byte multiples"""
    out = []
    for lvl in range(3):
        w, h, pad = width >> (2 - lvl), height >> (2 - lvl), ME_PAD[lvl]
        stride = (w + 2 * pad + 15) & ~15
        out.append((h + 2 * pad, stride, pad, w, h))
    return out

SAD_SIZES = [(128, 128), (128, 64), (64, 128), (64, 64), (64, 32), (64, 16), (32, 64), (32, 32), (32, 16), (32, 8), (16, 64), (16, 32), (16, 16), (16, 8), (16, 4), (8, 32), (8, 16), (8, 8), (8, 4), (4, 16), (4, 8), (4, 4)]  # (width, height) of the svt_aom_sadMxN family

LR_UNIT_DTYPE = np.dtype([("restoration_type", "<i4"), ("sgr_ep", "<i4"), ("sgr_xqd", "<i4", 2), ("hfilter", "<i2", 8), ("vfilter", "<i2", 8)])  # SvtB200LrUnitInfo
assert LR_UNIT_DTYPE.itemsize == 48


# SvtB200MeControls (include/svt_b200.h): the MeContext controls after svt_aom_sig_deriv_me, flattened; (name, int32 count)
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
ME_CONTROL_WORDS = sum(k for _, k in ME_CONTROL_FIELDS)
ME_B64_RESULT_FIELDS = ("total_me_candidate_index", "me_candidate_array", "me_mv_array", "distortion", "flags", "do_ref", "hme_centre", "zz_sad",
                        "best_sad", "best_mv")
# golden-fixture / comparison names of the ME outputs -> SvtB200MeB64Results field (best_sad / best_mv are undefined for pruned references: not compared whole)
ME_OUTPUT_NAMES = {"me_total": "total_me_candidate_index", "me_cand": "me_candidate_array", "me_mvs": "me_mv_array", "me_dist": "distortion",
                   "me_flags": "flags", "me_do_ref": "do_ref", "me_centre": "hme_centre", "me_zz": "zz_sad"}
