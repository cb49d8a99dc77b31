/* oracle/ref_glue.c -- TEST INFRASTRUCTURE.  Glue compiled INTO oracle/_ref/libsvtav1_ref.so next
 * to the unmodified reference objects.  It (1) satisfies the three symbols the kernel files
 * reference from translation units we do not build (Source/Lib/Globals, third_party/fastfeat),
 * none of which is reachable from the DSP path, and (2) binds the reference's own run-time
 * dispatch tables the way svt_av1_enc_init does (Source/Lib/Globals/enc_handle.c:1444-1445),
 * for the C tier ("--asm c").  No reference source text is copied here. */
#include <stdint.h>
#include <stdlib.h>

void set_segments_numbers(void) { abort(); }
void svt_aom_fast9_detect_nonmax(void) { abort(); }
int  svt_aom_tf_max_ref_per_struct(void) { abort(); return 0; }

extern void svt_aom_setup_common_rtcd_internal(uint64_t flags);
extern void svt_aom_setup_rtcd_internal(uint64_t flags);

/* flags = 0 -> every pointer is the *_c function (these objects are built without ARCH_X86_64) */
void ref_glue_init(void) {
    svt_aom_setup_common_rtcd_internal(0);
    svt_aom_setup_rtcd_internal(0);
}
