/* world/cheaptrick.h -- spectral envelope, legacy entry point
 * (reference: src/world/cheaptrick.h:16-81).  Runs world_b200_cheaptrick_batch with n_utts = 1
 * and scatters the rows into the caller's double** . */
#ifndef WORLD_CHEAPTRICK_H_
#define WORLD_CHEAPTRICK_H_
#include "world/macrodefinitions.h"
WORLD_BEGIN_C_DECLS

typedef struct {
  double q1;
  double f0_floor;
  int fft_size;
} CheapTrickOption;

/* spectrogram: f0_length row pointers, each to fft_size/2+1 doubles (caller-allocated). */
WORLD_API void CheapTrick(const double *x, int x_length, int fs, const double *temporal_positions,
                          const double *f0, int f0_length, const CheapTrickOption *option,
                          double **spectrogram);
/* q1 = -0.15, f0_floor = 71, fft_size = GetFFTSizeForCheapTrick (cheaptrick.cpp:231-240) */
WORLD_API void InitializeCheapTrickOption(int fs, CheapTrickOption *option);
/* 2^(1 + int(log(3 fs / f0_floor + 1) / log 2)) */
WORLD_API int GetFFTSizeForCheapTrick(int fs, const CheapTrickOption *option);
/* 3 fs / (fft_size - 3) */
WORLD_API double GetF0FloorForCheapTrick(int fs, int fft_size);

WORLD_END_C_DECLS
#endif /* WORLD_CHEAPTRICK_H_ */
