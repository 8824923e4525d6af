/* world/d4c.h -- band aperiodicity, legacy entry point (reference: src/world/d4c.h:16-47).
 * Runs world_b200_d4c_batch with n_utts = 1. */
#ifndef WORLD_D4C_H_
#define WORLD_D4C_H_
#include "world/macrodefinitions.h"
WORLD_BEGIN_C_DECLS

typedef struct {
  double threshold;
} D4COption;

/* aperiodicity: f0_length row pointers, each to fft_size/2+1 doubles; fft_size is CheapTrick's. */
WORLD_API void D4C(const double *x, int x_length, int fs, const double *temporal_positions,
                   const double *f0, int f0_length, int fft_size, const D4COption *option,
                   double **aperiodicity);
/* threshold = 0.85 (d4c.cpp:405-407) */
WORLD_API void InitializeD4COption(D4COption *option);

WORLD_END_C_DECLS
#endif /* WORLD_D4C_H_ */
