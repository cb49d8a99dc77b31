/*
 * svt_b200.h -- C ABI of libsvtav1_b200.so: the B200 (sm_100a) tier of SVT-AV1-PSY's inner-loop DSP.
 *
 * Two layers:
 *
 *  T1  "pointer" entry points.  Same argument list, argument meaning and output contract as the
 *      reference's run-time dispatched function pointers (declared RTCD_EXTERN in
 *      Source/Lib/Codec/aom_dsp_rtcd.h and common_dsp_rtcd.h, bound in
 *      svt_aom_setup_rtcd_internal, aom_dsp_rtcd.c:188, and svt_aom_setup_common_rtcd_internal,
 *      common_dsp_rtcd.c:466).  Every pointer argument is caller-owned HOST memory; outputs are
 *      fully written on return; the functions are re-entrant and thread safe.  A reference build
 *      installs them by assigning   svt_sad_loop_kernel = svt_b200_sad_loop_kernel;   etc. right
 *      after the two stock rtcd calls (Source/Lib/Globals/enc_handle.c:1444-1445); see
 *      INTEGRATION.md.  They are the parity surface.
 *
 *  T2  batch entry points (new): one call carries a whole picture / segment worth of work items.
 *      "_host" variants take host buffers (copies are part of the call); "_dev" variants take
 *      device pointers + a CUDA stream (passed as void*) and only enqueue work.
 *
 * No CPU fallback exists: svt_b200_init() fails unless an sm_100 device is present and every other
 * entry point aborts if it has not succeeded.
 *
 * Types are plain C (stdint); no CUDA or torch types appear in any signature.
 */
#ifndef SVT_B200_H
#define SVT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define SVT_B200_API __attribute__((visibility("default")))
#else
#define SVT_B200_API
#endif

/* ------------------------------------------------------------------------------------------ */
/* lifecycle                                                                                  */
/* ------------------------------------------------------------------------------------------ */
enum {
    SVT_B200_OK               = 0,
    SVT_B200_ERR_NO_DEVICE    = -1, /* maps to EB_ErrorInsufficientResources at svt_av1_enc_init time */
    SVT_B200_ERR_BAD_ARCH     = -2, /* device is not sm_100 */
    SVT_B200_ERR_ALREADY_INIT = -3,
    SVT_B200_ERR_BAD_ARG      = -4
};

/* Bind the library to CUDA device `device` (<0: device 0).  Replaces the CPU-flag probing of
 * svt_aom_get_cpu_flags_to_use (Source/Lib/Codec/common_dsp_rtcd.c:97). */
SVT_B200_API int                svt_b200_init(int device);
SVT_B200_API void               svt_b200_shutdown(void);
SVT_B200_API int                svt_b200_sm_count(void);
SVT_B200_API unsigned long long svt_b200_launch_count(void); /* kernels launched so far by this library */
SVT_B200_API const char*        svt_b200_version(void);

/* ------------------------------------------------------------------------------------------ */
/* K1/K3  SAD search + single SADs  (reference: Source/Lib/C_DEFAULT/compute_sad_c.c)           */
/* ------------------------------------------------------------------------------------------ */

/* T1: svt_sad_loop_kernel (aom_dsp_rtcd.h:779; C: compute_sad_c.c:58-101).  Full search of a
 * block_width x block_height 8-bit block over search_area_width x search_area_height positions;
 * first minimum in raster order wins (strict '<', best initialised to 0xffffff);
 * x/y_search_center are left untouched when nothing beats 0xffffff. */
SVT_B200_API void svt_b200_sad_loop_kernel(uint8_t* src, uint32_t src_stride, uint8_t* ref, uint32_t ref_stride,
                                           uint32_t block_height, uint32_t block_width, uint64_t* best_sad,
                                           int16_t* x_search_center, int16_t* y_search_center,
                                           uint32_t src_stride_raw, uint8_t skip_search_line,
                                           int16_t search_area_width, int16_t search_area_height);

/* T1: svt_nxm_sad_kernel (aom_dsp_rtcd.h:856; C: svt_nxm_sad_kernel_helper_c / compute_sad_c.c:20-37). */
SVT_B200_API uint32_t svt_b200_nxm_sad_kernel(const uint8_t* src, uint32_t src_stride, const uint8_t* ref,
                                              uint32_t ref_stride, uint32_t height, uint32_t width);

/* T2 work item: one full search.  Offsets are in bytes from the plane base pointers given to the
 * batch call.  src_stride/ref_stride are the row pitches used for BLOCK rows (2x the plane pitch in
 * SUB_SAD mode, motion_estimation.c:463-481); ref_step is the pitch between SEARCH rows (the
 * reference's `src_stride_raw`). */
typedef struct SvtB200SadSearchItem {
    uint64_t src_off;
    uint64_t ref_off;
    uint32_t src_stride;
    uint32_t ref_stride;
    uint32_t ref_step;
    uint16_t block_w;
    uint16_t block_h;
    int16_t  sa_w;
    int16_t  sa_h;
    uint16_t skip_search_line;
    uint16_t reserved;
} SvtB200SadSearchItem;

typedef struct SvtB200SadSearchResult {
    uint32_t best_sad; /* 0xffffff when no position was evaluated */
    int16_t  x;        /* search-area-relative, -1 when none       */
    int16_t  y;
} SvtB200SadSearchResult;

SVT_B200_API int svt_b200_sad_search_batch_host(const uint8_t* src_plane, size_t src_bytes, const uint8_t* ref_plane,
                                                size_t ref_bytes, const SvtB200SadSearchItem* items, int n_items,
                                                SvtB200SadSearchResult* results);
/* device-resident planes/items/results; `max_*` bound the items' geometry (for shared-memory sizing) */
SVT_B200_API int svt_b200_sad_search_batch_dev(const uint8_t* d_src_plane, const uint8_t* d_ref_plane,
                                               const SvtB200SadSearchItem* d_items, int n_items,
                                               SvtB200SadSearchResult* d_results, int max_block_w, int max_block_h,
                                               int max_sa_w, int max_sa_h, int max_row_mult, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SVT_B200_H */
