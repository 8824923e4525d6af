/* world/dio.h -- DIO F0 estimator, legacy single-utterance entry point.
 * Same symbols and struct layout as the reference's src/world/dio.h:16-57; the work runs on the
 * GPU (world_b200_dio_batch with n_utts = 1).  Host pointers in, host pointers out. */
#ifndef WORLD_DIO_H_
#define WORLD_DIO_H_
#include "world/macrodefinitions.h"
WORLD_BEGIN_C_DECLS

typedef struct {
  double f0_floor;
  double f0_ceil;
  double channels_in_octave;
  double frame_period; /* ms */
  int speed;           /* decimation ratio 1..12 */
  double allowed_range;
} DioOption;

/* f0 and temporal_positions must hold GetSamplesForDIO() doubles. */
WORLD_API void Dio(const double *x, int x_length, int fs, const DioOption *option,
                   double *temporal_positions, double *f0);
/* floor 71, ceil 800, 2 channels/octave, 5 ms, speed 1, allowed_range 0.1 (dio.cpp:650-666) */
WORLD_API void InitializeDioOption(DioOption *option);
WORLD_API int GetSamplesForDIO(int fs, int x_length, double frame_period);

WORLD_END_C_DECLS
#endif /* WORLD_DIO_H_ */
