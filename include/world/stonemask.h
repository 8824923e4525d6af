/* world/stonemask.h -- F0 refinement by instantaneous frequency, legacy entry point
 * (reference: src/world/stonemask.h:27).  Runs world_b200_stonemask_batch with n_utts = 1. */
#ifndef WORLD_STONEMASK_H_
#define WORLD_STONEMASK_H_
#include "world/macrodefinitions.h"
WORLD_BEGIN_C_DECLS

WORLD_API void StoneMask(const double *x, int x_length, int fs, const double *temporal_positions,
                         const double *f0, int f0_length, double *refined_f0);

WORLD_END_C_DECLS
#endif /* WORLD_STONEMASK_H_ */
