/*
 * ref_shim.c -- TEST INFRASTRUCTURE.  Compiles the UNMODIFIED reference header into a shared
 * object (oracle/_ref/libgs_ref.so) so tests and bench.py can run the real reference.
 *
 * Nothing is copied: the reference is included by path at build time (-I$(REF) from
 * oracle/Makefile).  `#define GS_API` (empty) gives every gs_* function external linkage --
 * the same switch the reference's own wasm build uses (examples/wasm/grayskull.c:33-34).
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may load this.
 */
#define GS_API
#include "grayskull.h"
#include "examples/nanomagick/frontalface.h"

const struct gs_lbp_cascade *ref_frontalface(void) { return &frontalface; }
const int *ref_brief_pattern(void) { return &gs_brief_pattern[0][0]; }
/* static in the reference (grayskull.h:639-649): exported for the stable-sort parity test */
void ref_sort_keypoints(struct gs_keypoint *kps, unsigned n) {
  if (n > 1) gs_sort_keypoints(kps, n);
}
unsigned ref_sizeof(int which) {
  switch (which) {
    case 0: return sizeof(struct gs_image);
    case 1: return sizeof(struct gs_rect);
    case 2: return sizeof(struct gs_keypoint);
    case 3: return sizeof(struct gs_lbp_cascade);
    default: return 0;
  }
}
