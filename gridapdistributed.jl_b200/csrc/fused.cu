// fused.cu -- fused affine route (placeholder until the kernel lands in this file)
#include "internal.h"
bool fused_affine_available(graft_ctx*) { return false; }
void numeric_fused_affine(graft_ctx*, int) { graft_throw("fused route not built"); }
