/* world/harvest.h -- Harvest F0 estimator, legacy single-utterance entry point
 * (reference: src/world/harvest.h:16-55).  Runs world_b200_harvest_batch with n_utts = 1. */
#ifndef WORLD_HARVEST_H_
#define WORLD_HARVEST_H_
#include "world/macrodefinitions.h"
WORLD_BEGIN_C_DECLS

typedef struct {
  double f0_floor;
  double f0_ceil;
  double frame_period; /* ms */
} HarvestOption;

WORLD_API void Harvest(const double *x, int x_length, int fs, const HarvestOption *option,
                       double *temporal_positions, double *f0);
/* floor 71, ceil 800, 5 ms (harvest.cpp:1257-1262) */
WORLD_API void InitializeHarvestOption(HarvestOption *option);
WORLD_API int GetSamplesForHarvest(int fs, int x_length, double frame_period);

WORLD_END_C_DECLS
#endif /* WORLD_HARVEST_H_ */
