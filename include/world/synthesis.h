/* world/synthesis.h -- waveform synthesis from f0 / spectral envelope / aperiodicity, legacy entry
 * point (reference: src/world/synthesis.h:30).  Runs world_b200_synthesis_batch with n_utts = 1. */
#ifndef WORLD_SYNTHESIS_H_
#define WORLD_SYNTHESIS_H_
#include "world/macrodefinitions.h"
WORLD_BEGIN_C_DECLS

/* spectrogram / aperiodicity: f0_length row pointers of fft_size/2+1 doubles; y: y_length doubles. */
WORLD_API void Synthesis(const double *f0, int f0_length, const double *const *spectrogram,
                         const double *const *aperiodicity, int fft_size, double frame_period, int fs,
                         int y_length, double *y);

WORLD_END_C_DECLS
#endif /* WORLD_SYNTHESIS_H_ */
