#!/usr/bin/env python3
"""Dev-time tool: read the reference's coefficient scan orders (coefficients.h:2197 av1_scan_orders) and quantization
matrices (q_matrices.h wt_matrix_ref / iwt_matrix_ref, laid out as svt_av1_qm_init walks them, md_config_process.c:232)
out of the compiled reference (oracle/_ref/libsvtav1_ref.so, accessors in oracle/ref_tables.c) and write
svt-av1-psy_b200/av1_tables.npz.  The generated file is committed; this needs /root/reference only here.

  scan / iscan : int16, every (tx_size, tx_type) back to back;  scan_off[tx_size, tx_type] -> start, scan_len[tx_size]
  qm / iqm     : uint8 [level 0..14][luma|chroma], every tx_size back to back;  qm_off[tx_size] -> start (len = scan_len)
"""
import ctypes as ct
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402

ref = oracle.ref
TX_SIZES_ALL, TX_TYPES, QM_LEVELS = 19, 16, 15  # level 15 = "no matrix"
i16p, u8p = ct.POINTER(ct.c_int16), ct.POINTER(ct.c_uint8)
ref.ref_scan_order.restype = ct.c_int
ref.ref_scan_order.argtypes = [ct.c_int, ct.c_int, i16p, i16p]
ref.ref_qm_matrix.restype = ct.c_int
ref.ref_qm_matrix.argtypes = [ct.c_int, ct.c_int, ct.c_int, u8p, u8p]

scan, iscan = [], []
scan_off = np.zeros((TX_SIZES_ALL, TX_TYPES), np.int32)
scan_len = np.zeros(TX_SIZES_ALL, np.int32)
pos = 0
for sz in range(TX_SIZES_ALL):
    n = ref.ref_scan_order(sz, 0, None, None)
    scan_len[sz] = n
    for ty in range(TX_TYPES):
        s, i = np.zeros(n, np.int16), np.zeros(n, np.int16)
        assert ref.ref_scan_order(sz, ty, s.ctypes.data_as(i16p), i.ctypes.data_as(i16p)) == n
        assert np.array_equal(np.sort(s), np.arange(n)) and np.array_equal(i[s], np.arange(n)), (sz, ty)
        scan.append(s)
        iscan.append(i)
        scan_off[sz, ty] = pos
        pos += n

qm_off = np.concatenate([[0], np.cumsum(scan_len)[:-1]]).astype(np.int32)
total = int(scan_len.sum())
qm = np.zeros((QM_LEVELS, 2, total), np.uint8)
iqm = np.zeros_like(qm)
for lv in range(QM_LEVELS):
    for pl in range(2):
        for sz in range(TX_SIZES_ALL):
            n, o = int(scan_len[sz]), int(qm_off[sz])
            got = ref.ref_qm_matrix(lv, pl, sz, qm[lv, pl, o:o + n].ctypes.data_as(u8p), iqm[lv, pl, o:o + n].ctypes.data_as(u8p))
            assert got == n, (lv, pl, sz, got, n)
assert ref.ref_qm_matrix(QM_LEVELS, 0, 0, None, None) == 0

out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "svt-av1-psy_b200", "av1_tables.npz")
np.savez_compressed(out, scan=np.concatenate(scan), iscan=np.concatenate(iscan), scan_off=scan_off, scan_len=scan_len,
                    qm=qm, iqm=iqm, qm_off=qm_off)
print("wrote", out, os.path.getsize(out), "bytes;", pos, "scan entries,", total, "QM bytes per (level, plane type)")
