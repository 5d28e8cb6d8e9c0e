// nb200_xcor_tc.cu -- tcgen05 (5th-generation tensor core) version of the dense background
// cross-correlation.  Placeholder until the kernel lands: reports "not available" so that
// xcor_mode 2 fails loudly instead of silently running the fp64 kernel.
#include "nb200_dev.cuh"

int nb200_tc_setup(nb200_ctx *) { return NB200_OK; }
int nb200_tc_available(nb200_ctx *) { return 0; }
int nb200_nuc_bx_tc(nb200_ctx *ctx, nb200_dbatch *) { return nb200_fail(ctx, NB200_ERR_STATE, "tcgen05 xcor not built"); }
