"""Reference-driven CDEF strength search: replays cdef_seg_search (cdef_process.c:106-352) in Python on
top of the UNMODIFIED reference functions svt_cdef_filter_fb (cdef.c:339) and
svt_aom_compute_cdef_dist_c (enc_cdef.c:129), 16-bit pixel containers."""
import ctypes as ct

import numpy as np

BS, VL = 144, 0x7f7f


def P(a, off_elems=0):
    return ct.c_void_p(a.ctypes.data + off_elems * a.itemsize)


def make_frame(r, width, height, bd, skip_prob=0.3):
    """deblocked-recon-like and source planes (uint16 containers), + luma 8x8 skip map"""
    lim = (1 << bd) - 1
    planes_src, planes_rec = [], []
    for pli in range(3):
        w, h = (width >> (pli > 0)), (height >> (pli > 0))
        yy, xx = np.mgrid[0:h, 0:w]
        base = (np.sin(xx / 7.0) * 40 + np.cos(yy / 5.0) * 30 + (xx // 16 + yy // 16) % 2 * 50 + 100) * (lim / 255.0)
        src = np.clip(base + r.normal(0, 2 * lim / 255.0, base.shape), 0, lim).astype(np.uint16)
        rec = np.clip(src.astype(np.int32) + r.integers(-12, 13, src.shape) * (lim // 255), 0, lim).astype(np.uint16)
        planes_src.append(np.ascontiguousarray(src))
        planes_rec.append(np.ascontiguousarray(rec))
    skip = (r.random(((height + 7) // 8, (width + 7) // 8)) < skip_prob).astype(np.uint8)
    return planes_rec, planes_src, skip


def ref_cdef_search(refc, rec, src, skip, width, height, bd, damping, subsampling, str_y, str_uv):
    nhfb, nvfb = (width + 63) // 64, (height + 63) // 64
    nfb, ng, cs = nhfb * nvfb, len(str_y), max(bd - 8, 0)
    mse = np.zeros((2, nfb, ng), np.uint64)
    dirs = np.zeros((nfb, 64), np.uint8)
    vars_ = np.zeros((nfb, 64), np.int32)
    ffb = refc.svt_cdef_filter_fb; ffb.restype = None
    fdist = refc.svt_aom_compute_cdef_dist_c; fdist.restype = ct.c_uint64
    inbuf = np.zeros(BS * 70, np.uint16)
    for fbr in range(nvfb):
        for fbc in range(nhfb):
            fb = fbr * nhfb + fbc
            lst = [(by, bx) for by in range(8) for bx in range(8)
                   if fbr * 8 + by < skip.shape[0] and fbc * 8 + bx < skip.shape[1] and not skip[fbr * 8 + by, fbc * 8 + bx]]
            if not lst:
                continue
            dlist = np.array(lst, np.uint8).reshape(-1)
            dirm = np.zeros((16, 16), np.uint8); varm = np.zeros((16, 16), np.int32)
            dirinit = ct.c_int32(0)
            for pli in range(3):
                dec = 1 if pli else 0
                fbs, pw, ph = 64 >> dec, width >> dec, height >> dec
                hsz, vsz = min(fbs, pw - fbc * fbs), min(fbs, ph - fbr * fbs)
                inbuf[:] = VL
                yoff, xoff = 3 * (fbr != 0), 8 * (fbc != 0)
                ysize = vsz + 3 * (fbr + 1 < nvfb) + yoff
                xsize = hsz + 8 * (fbc + 1 < nhfb) + xoff
                tile = inbuf.reshape(70, BS)
                tile[3 - yoff:3 - yoff + ysize, 8 - xoff:8 - xoff + xsize] = \
                    rec[pli][fbr * fbs - yoff:fbr * fbs - yoff + ysize, fbc * fbs - xoff:fbc * fbs - xoff + xsize]
                subs = min(subsampling, 1 if dec else 4)
                bsize = 0 if dec else 3
                for g in range(ng):
                    sv = str_uv[g] if pli else str_y[g]
                    if sv < 0:
                        mse[1, fb, g] = 1040400 * 64
                        continue
                    pri, sec = sv // 4, sv % 4
                    tmp = np.zeros(64 * 64, np.uint16)
                    ffb(None, P(tmp), 0, P(inbuf, 3 * BS + 8), dec, dec, P(dirm), ct.byref(dirinit), P(varm), pli, P(dlist), len(lst),
                        pri, sec + (sec == 3), damping, damping, cs, ct.c_uint8(subs))
                    sp = src[pli]
                    d = fdist(P(sp, fbr * fbs * sp.shape[1] + fbc * fbs), sp.shape[1], P(tmp), P(dlist), len(lst), bsize, cs, pli,
                              ct.c_uint8(subs))
                    if pli == 2:
                        mse[1, fb, g] += d * subs
                    else:
                        mse[1 if pli else 0, fb, g] = d * subs
            for (by, bx) in lst:
                dirs[fb, by * 8 + bx] = dirm[by, bx]
                vars_[fb, by * 8 + bx] = varm[by, bx]
    return mse, dirs, vars_


def port_cdef_search(port, rec, src, skip, width, height, bd, damping, subsampling, str_y, str_uv):
    nfb, ng = ((width + 63) // 64) * ((height + 63) // 64), len(str_y)
    mse = np.zeros((2, nfb, ng), np.uint64); dirs = np.zeros((nfb, 64), np.uint8); vars_ = np.zeros((nfb, 64), np.int32)
    PA = ct.c_void_p * 3
    IA = ct.c_int * 3
    sy = np.array(str_y, np.int32); su = np.array(str_uv, np.int32)
    port.port_cdef_search_frame.restype = None
    port.port_cdef_search_frame(PA(*[x.ctypes.data for x in rec]), IA(*[x.shape[1] for x in rec]), PA(*[x.ctypes.data for x in src]),
                                IA(*[x.shape[1] for x in src]), width, height, bd, damping, subsampling, P(skip), P(sy), P(su), ng,
                                P(mse), P(dirs), P(vars_))
    return mse, dirs, vars_
