/* svt_b200_me_glue.c -- see svt_b200_me_glue.h.  Compiled against the reference's headers (Source/Lib/Codec). */
#include <stdlib.h>
#include <string.h>
#include "definitions.h"
#include "pcs.h"
#include "sequence_control_set.h"
#include "me_context.h"
#include "svt_b200_me_glue.h"

void svt_b200_me_controls_from_context(const PictureParentControlSet* pcs, const MeContext* me, SvtB200MeControls* c) {
    memset(c, 0, sizeof(*c));
    c->n_list = me->num_of_list_to_search;
    c->n_ref[0] = me->num_of_ref_pic_to_search[0];
    c->n_ref[1] = me->num_of_list_to_search > 1 ? me->num_of_ref_pic_to_search[1] : 0;
    c->temporal_layer_index = me->temporal_layer_index;
    c->is_ref = me->is_ref;
    c->hierarchical_levels = pcs->hierarchical_levels;
    for (int l = 0; l < c->n_list; l++)
        for (int r = 0; r < c->n_ref[l]; r++) /* get_me_reference (motion_estimation.c:1232-1235) */
            c->dist[l][r] = (int32_t)llabs((long long)pcs->picture_number - (long long)me->me_ds_ref_array[l][r].picture_number);
    c->enable_hme = me->enable_hme_flag; c->enable_l0 = me->enable_hme_level0_flag;
    c->enable_l1 = me->enable_hme_level1_flag; c->enable_l2 = me->enable_hme_level2_flag;
    c->hme_sub_sad = me->hme_search_method == SUB_SAD_SEARCH; c->me_sub_sad = me->me_search_method == SUB_SAD_SEARCH;
    c->hme_l0_min_w = me->hme_l0_sa.sa_min.width; c->hme_l0_min_h = me->hme_l0_sa.sa_min.height;
    c->hme_l0_max_w = me->hme_l0_sa.sa_max.width; c->hme_l0_max_h = me->hme_l0_sa.sa_max.height;
    c->hme_l1_w = me->hme_l1_sa.width; c->hme_l1_h = me->hme_l1_sa.height; c->hme_l2_w = me->hme_l2_sa.width; c->hme_l2_h = me->hme_l2_sa.height;
    c->me_min_w = me->me_sa.sa_min.width; c->me_min_h = me->me_sa.sa_min.height; c->me_max_w = me->me_sa.sa_max.width; c->me_max_h = me->me_sa.sa_max.height;
    c->prehme_enable = me->prehme_ctrl.enable;
    for (int s = 0; s < 2; s++) {
        c->prehme_sa[s][0] = me->prehme_ctrl.prehme_sa_cfg[s].sa_min.width; c->prehme_sa[s][1] = me->prehme_ctrl.prehme_sa_cfg[s].sa_min.height;
        c->prehme_sa[s][2] = me->prehme_ctrl.prehme_sa_cfg[s].sa_max.width; c->prehme_sa[s][3] = me->prehme_ctrl.prehme_sa_cfg[s].sa_max.height;
    }
    c->prehme_skip_search_line = me->prehme_ctrl.skip_search_line; c->prehme_l1_early_exit = me->prehme_ctrl.l1_early_exit;
    const MeHmeRefPruneCtrls* p = &me->me_hme_prune_ctrls;
    c->prune_enable = p->enable_me_hme_ref_pruning; c->prune_hme_th = p->prune_ref_if_hme_sad_dev_bigger_than_th;
    c->prune_me_th = p->prune_ref_if_me_sad_dev_bigger_than_th;
    c->zz_sad_th = (int32_t)p->zz_sad_th; c->zz_sad_pct = p->zz_sad_pct; c->phme_sad_th = (int32_t)p->phme_sad_th; c->phme_sad_pct = p->phme_sad_pct;
    const MeSrCtrls* s = &me->me_sr_adjustment_ctrls;
    c->sr_enable = s->enable_me_sr_adjustment; c->sr_mv_length_th = s->reduce_me_sr_based_on_mv_length_th;
    c->sr_stationary_hme_sad_abs_th = s->stationary_hme_sad_abs_th; c->sr_stationary_divisor = s->stationary_me_sr_divisor;
    c->sr_hme_sad_abs_th = s->reduce_me_sr_based_on_hme_sad_abs_th; c->sr_low_hme_sad_divisor = s->me_sr_divisor_for_low_hme_sad;
    c->sr_distance_based_hme_resizing = s->distance_based_hme_resizing;
    c->var_enable = me->me_8x8_var_ctrls.enabled; c->var_div4_th = (int32_t)me->me_8x8_var_ctrls.me_sr_div4_th;
    c->var_div2_th = (int32_t)me->me_8x8_var_ctrls.me_sr_div2_th; c->var_mult2_th = (int32_t)me->me_8x8_var_ctrls.me_sr_mult2_th;
    c->mvsa_enable = me->mv_based_sa_adj.enabled; c->mvsa_nearest_ref_only = me->mv_based_sa_adj.nearest_ref_only;
    c->mvsa_mv_size_th = me->mv_based_sa_adj.mv_size_th; c->mvsa_multiplier = me->mv_based_sa_adj.sa_multiplier;
    c->reduce_hme_l0_sr_th_min = me->reduce_hme_l0_sr_th_min; c->reduce_hme_l0_sr_th_max = me->reduce_hme_l0_sr_th_max;
    c->me_early_exit_th = (int32_t)me->me_early_exit_th; c->me_safe_limit_zz_th = (int32_t)me->me_safe_limit_zz_th;
    c->prev_me_stage_based_exit_th = (int32_t)me->prev_me_stage_based_exit_th; c->prune_me_candidates_th = me->prune_me_candidates_th;
    c->use_best_unipred_cand_only = me->use_best_unipred_cand_only;
    c->similar_brightness_refs = pcs->similar_brightness_refs; c->only_l_bwd = pcs->scs->mrp_ctrls.only_l_bwd;
    c->enable_me_8x8 = pcs->enable_me_8x8; c->enable_me_16x16 = pcs->enable_me_16x16;
    c->max_cand = pcs->pa_me_data->max_cand; c->max_refs = pcs->pa_me_data->max_refs; c->max_l0 = pcs->pa_me_data->max_l0;
    c->gm_enabled = pcs->gm_ctrls.enabled; c->gm_use_distance_based_active_th = pcs->gm_ctrls.use_distance_based_active_th;
    c->resolution_le_480p = pcs->scs->input_resolution <= INPUT_SIZE_480p_RANGE;
}

void svt_b200_me_store_results(PictureParentControlSet* pcs, int n_b64, int n_pu, const uint8_t* total, const uint8_t* cand, const uint32_t* mv,
                               const uint32_t* distortion, const uint8_t* flags) {
    const int max_cand = pcs->pa_me_data->max_cand, max_refs = pcs->pa_me_data->max_refs;
    for (int b = 0; b < n_b64; b++) {
        MeSbResults* r = pcs->pa_me_data->me_results[b];
        memcpy(r->total_me_candidate_index, total + (size_t)b * n_pu, (size_t)n_pu);
        memcpy(r->me_candidate_array, cand + (size_t)b * n_pu * max_cand, (size_t)n_pu * max_cand); /* MeCandidate is one byte of bit fields */
        memcpy(r->me_mv_array, mv + (size_t)b * n_pu * max_refs, (size_t)n_pu * max_refs * sizeof(uint32_t));
        const uint32_t* d = distortion + (size_t)b * 6;
        pcs->rc_me_distortion[b] = d[0]; pcs->me_64x64_distortion[b] = d[1]; pcs->me_32x32_distortion[b] = d[2];
        pcs->me_16x16_distortion[b] = d[3]; pcs->me_8x8_distortion[b] = d[4]; pcs->me_8x8_cost_variance[b] = d[5];
        pcs->stationary_block_present_sb[b] = flags[2 * b];
        pcs->rc_me_allow_gm[b] = flags[2 * b + 1];
    }
}
