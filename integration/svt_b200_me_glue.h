/* svt_b200_me_glue.h -- reference-side glue of the T2 open-loop ME call (svt_b200_me_b64_picture_dev, include/svt_b200.h).
 *
 * A reference build adds integration/svt_b200_me_glue.c next to Source/Lib/Codec/me_process.c.  In place of the 64x64-block
 * loop of svt_aom_motion_estimation_kernel (me_process.c:174-291) the ME process then does, once per picture:
 *     svt_aom_sig_deriv_me(scs, pcs, me_ctx);                          // unchanged (me_process.c:120)
 *     ... me_ctx->num_of_list_to_search / num_of_ref_pic_to_search / me_ds_ref_array set as at me_process.c:219-262 ...
 *     svt_b200_me_controls_from_context(pcs, me_ctx, &ctrl);           // this file
 *     svt_b200_me_b64_picture_dev(&cur, refs, &ctrl, &dev_out, stream);  // pyramids are device-resident pictures
 *     (one D2H copy of the result arena)
 *     svt_b200_me_store_results(pcs, n_pu, host arrays...);            // this file
 * Both functions are compiled against the reference's own headers (a renamed or re-typed MeContext field is a compile error)
 * and are exercised by the parity tests: oracle/ref_me_b64.c derives the controls it hands to the tests through the first,
 * and round-trips the reference's results through the second. */
#ifndef SVT_B200_ME_GLUE_H
#define SVT_B200_ME_GLUE_H
#include <stdint.h>
#include "svt_b200.h"
#ifdef __cplusplus
extern "C" {
#endif
struct PictureParentControlSet;
struct MeContext;
/* MeContext controls (after svt_aom_sig_deriv_me) + the picture-level fields the driver reads -> SvtB200MeControls */
void svt_b200_me_controls_from_context(const struct PictureParentControlSet* pcs, const struct MeContext* me_ctx, SvtB200MeControls* ctrl);
/* host copies of the call's outputs -> pcs->pa_me_data->me_results[] and the per-block pcs arrays
 * (total / cand / mv laid out as SvtB200MeB64Results documents; distortion [n_b64][6]; flags [n_b64][2]) */
void svt_b200_me_store_results(struct PictureParentControlSet* pcs, int n_b64, int n_pu, const uint8_t* total, const uint8_t* cand,
                               const uint32_t* mv, const uint32_t* distortion, const uint8_t* flags);
#ifdef __cplusplus
}
#endif
#endif
