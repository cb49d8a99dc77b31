// TEST INFRASTRUCTURE (oracle side): accessors for two constant tables of the reference that are `static const` in its
// headers and therefore invisible to the dynamic linker.  tools/dump_av1_tables.py calls these once, in the container that
// has /root/reference, and writes svt-av1-psy_b200/av1_tables.npz; the synthetic workload then quantizes with the
// reference's real scan orders and quantization matrices instead of stand-ins.  Nothing in the product path links this.
//
//   scan orders   : Source/Lib/Codec/coefficients.h:2197  av1_scan_orders[TX_SIZES_ALL][TX_TYPES]
//   QM weights    : Source/Lib/Codec/q_matrices.h:24 / :6657  wt_matrix_ref / iwt_matrix_ref, laid out per transform size the
//                   way svt_av1_qm_init walks them (Source/Lib/Codec/md_config_process.c:232-256)
#include <stdint.h>
#include <string.h>
#include "definitions.h"
#include "coefficients.h"
#include "q_matrices.h"
#include "inv_transforms.h"

// number of coded coefficients of (tx_size): 64-point dimensions code their 32 low-frequency rows/columns only
static int coded_count(int tx_size) {
    int w = tx_size_wide[tx_size], h = tx_size_high[tx_size];
    if (w > 32) w = 32;
    if (h > 32) h = 32;
    return w * h;
}

// scan[i] = raster position of the i-th coefficient in coding order; iscan = its inverse.  Returns the count.
int ref_scan_order(int tx_size, int tx_type, int16_t* scan, int16_t* iscan) {
    if (tx_size < 0 || tx_size >= TX_SIZES_ALL || tx_type < 0 || tx_type >= TX_TYPES) return -1;
    const ScanOrder* so = &av1_scan_orders[tx_size][tx_type];
    const int        n  = coded_count(tx_size);
    if (scan) memcpy(scan, so->scan, n * sizeof(int16_t));
    if (iscan) memcpy(iscan, so->iscan, n * sizeof(int16_t));
    return n;
}

// the weight (qm) and inverse weight (iqm) matrix of (level, plane, tx_size); 0 when the level has none (level 15 = flat)
int ref_qm_matrix(int level, int plane, int tx_size, uint8_t* qm, uint8_t* iqm) {
    if (level < 0 || level >= NUM_QM_LEVELS || tx_size < 0 || tx_size >= TX_SIZES_ALL || plane < 0 || plane > 2) return -1;
    if (level == NUM_QM_LEVELS - 1) return 0;
    const int target  = av1_get_adjusted_tx_size((TxSize)tx_size);  // 64-point sizes share the 32-point matrices
    int       current = 0;
    for (int t = 0; t < TX_SIZES_ALL; ++t) {
        if (t != (int)av1_get_adjusted_tx_size((TxSize)t)) continue;  // borrows another size's matrix: no storage of its own
        const int size = tx_size_2d[t];
        if (t == target) {
            if (qm) memcpy(qm, &wt_matrix_ref[level][plane >= 1][current], size);
            if (iqm) memcpy(iqm, &iwt_matrix_ref[level][plane >= 1][current], size);
            return size;
        }
        current += size;
    }
    return -1;
}

/* the input-resolution class the preset tables key on (sequence_control_set.c:113) */
#include "sequence_control_set.h"
int ref_input_resolution_class(uint32_t input_size) {
    EbInputResolution r;
    svt_aom_derive_input_resolution(&r, input_size);
    return (int)r;
}
