/* world/matlabfunctions.h -- the small public helpers of the reference (src/world/matlabfunctions.h:21-149)
 * that callers use next to the analysis API (test/test.cpp:244 stretches a spectrum with interp1): host
 * functions, same names and semantics, results bit-identical to the reference's.  fast_fftfilt() is not
 * provided: it takes the reference's internal FFT plan structs (common.h), which this library does not have. */
#ifndef WORLD_MATLABFUNCTIONS_H_
#define WORLD_MATLABFUNCTIONS_H_
#include <stdint.h>
#include "world/macrodefinitions.h"
WORLD_BEGIN_C_DECLS

/* swaps the two halves of x (x_length even) */
WORLD_API void fftshift(const double *x, int x_length, double *y);
/* index[i] = number of the bin [x[k-1], x[k]) that holds edges[i], 1-based, both vectors ascending */
WORLD_API void histc(const double *x, int x_length, const double *edges, int edges_length, int *index);
/* piecewise-linear interpolation of (x, y) at ascending xi; extrapolates with the end segments */
WORLD_API void interp1(const double *x, const double *y, int x_length, const double *xi, int xi_length,
                       double *yi);
/* zero-phase decimation by r in 2..12 (other r: all-zero filter, like the reference);
 * y receives (x_length - 1) / r + 1 samples (one more when r does not divide x_length + 8) */
WORLD_API void decimate(const double *x, int x_length, int r, double *y);
/* round half away from zero */
WORLD_API int matlab_round(double x);
/* y[i] = x[i + 1] - x[i], x_length - 1 values */
WORLD_API void diff(const double *x, int x_length, double *y);
/* interp1 on the uniform grid x, x + shift, ...; no bounds checks, last segment has zero slope */
WORLD_API void interp1Q(double x, double shift, const double *y, int x_length, const double *xi,
                        int xi_length, double *yi);

typedef struct {
  uint32_t g_randn_x;
  uint32_t g_randn_y;
  uint32_t g_randn_z;
  uint32_t g_randn_w;
} RandnState;
/* xorshift128, twelve steps per value: sum of (w >> 4) / 2^28 - 6 */
WORLD_API double randn(RandnState *state);
WORLD_API void randn_reseed(RandnState *state);

/* sample standard deviation (divides by x_length - 1) */
WORLD_API double matlab_std(const double *x, int x_length);

WORLD_END_C_DECLS
#endif /* WORLD_MATLABFUNCTIONS_H_ */
