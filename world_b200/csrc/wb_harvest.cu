#include "wb_internal.h"
namespace wb {
int harvest_run(Ctx *ctx, const Batch &, const HarvestParams &, double *, double *) { ctx->last_error = "harvest: not built yet"; return 3; }
}
