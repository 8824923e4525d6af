// wb_f0common.cuh -- building blocks shared by the two F0 estimators (DIO and Harvest):
//   * decimate(): MATLAB-style zero-phase IIR decimator (matlabfunctions.cpp:27-125, 178-204)
//   * band sweep: one CTA per (utterance, band) runs the band's FIR over the whole utterance in
//     shared-memory tiles and picks the four zero-crossing event trains on the fly
//     (dio.cpp:296-435 / harvest.cpp:99-238), then interpolates the trains onto the frame grid
//     (interp1 of dio.cpp:471-519 / harvest.cpp:262-298).
//
// The reference filters with whole-utterance FFTs (2^17..2^21 points) - far beyond shared
// memory.  Its filters are short FIRs (<= ~1000 taps), its FFT size is chosen so that the
// circular convolution never wraps (dio.cpp:592-594, harvest.cpp:1164-1165), and the spectral
// mirroring quirk is inert, so the filtered signal IS the linear
// convolution; it is evaluated directly, register-tiled, FP64 FMA bound.  Filter taps are
// computed on the host with the same libm expressions as the reference and uploaded.
#pragma once
#include "wb_platform.cuh"
#include "wb_block.cuh"

namespace wb {

#define WB_SWEEP_T 2048      // outputs per tile
#define WB_SWEEP_R 8         // outputs per thread group (register tile)
#define WB_SWEEP_THREADS 256 // = T / R

// padded shared layout: one spare double per 8 keeps stride-8 accesses conflict free
WB_DEV int pad8(int i) { return i + (i >> 3); }

// ------------------------------------------------------------------------------ decimate
// IIR coefficients per ratio (matlabfunctions.cpp:29-113); unsupported ratios give zeros like
// the reference's default branch.
WB_HD inline void decimate_coefficients(int r, double a[3], double b[2]) {
  switch (r) {
    case 11: a[0] = 2.450743295230728; a[1] = -2.06794904601978; a[2] = 0.59574774438332101;
      b[0] = 0.0026822508007163792; b[1] = 0.0080467524021491377; break;
    case 12: a[0] = 2.4981398605924205; a[1] = -2.1368928194784025; a[2] = 0.62187513816221485;
      b[0] = 0.0021097275904709001; b[1] = 0.0063291827714127002; break;
    case 10: a[0] = 2.3936475118069387; a[1] = -1.9873904075111861; a[2] = 0.5658879979027055;
      b[0] = 0.0034818622251927556; b[1] = 0.010445586675578267; break;
    case 9: a[0] = 2.3236003491759578; a[1] = -1.8921545617463598; a[2] = 0.53148928133729068;
      b[0] = 0.0046331164041389372; b[1] = 0.013899349212416812; break;
    case 8: a[0] = 2.2357462340187593; a[1] = -1.7780899984041358; a[2] = 0.49152555365968692;
      b[0] = 0.0063522763407111993; b[1] = 0.019056829022133598; break;
    case 7: a[0] = 2.1225239019534703; a[1] = -1.6395144861046302; a[2] = 0.44469707800587366;
      b[0] = 0.0090366882681608418; b[1] = 0.027110064804482525; break;
    case 6: a[0] = 1.9715352749512141; a[1] = -1.4686795689225347; a[2] = 0.3893908434965701;
      b[0] = 0.013469181309343825; b[1] = 0.040407543928031475; break;
    case 5: a[0] = 1.7610939654280557; a[1] = -1.2554914843859768; a[2] = 0.3237186507788215;
      b[0] = 0.021334858522387423; b[1] = 0.06400457556716227; break;
    case 4: a[0] = 1.4499664446880227; a[1] = -0.98943497080950582; a[2] = 0.24578252340690215;
      b[0] = 0.036710750339322612; b[1] = 0.11013225101796784; break;
    case 3: a[0] = 0.95039378983237421; a[1] = -0.67429146741526791; a[2] = 0.15412211621346475;
      b[0] = 0.071221945171178636; b[1] = 0.21366583551353591; break;
    case 2: a[0] = 0.041156734567757189; a[1] = -0.42599112459189636; a[2] = 0.041037215479961225;
      b[0] = 0.16797464681802227; b[1] = 0.50392394045406674; break;
    default: a[0] = a[1] = a[2] = 0.0; b[0] = b[1] = 0.0;
  }
}

// One thread decimates one (virtually edge-padded) signal:  xin(i) = x[clamp(i - lag, 0, n-1)]
// for i in [0, n + 2 lag)  (harvest.cpp:43-66; lag = 0 gives plain decimate()).
// tmp1/tmp2: scratch of n + 2 lag + 18 doubles each.  Writes out[0..n_out) = decimated samples
// starting at decimated index `first`; returns how many decimated samples exist.
WB_DEV int decimate_one(const double *__restrict__ x, int n, int lag, int r, double *tmp1, double *tmp2,
                        int first, int n_out, double *out) {
  const int kNFact = 9;
  const int nx = n + 2 * lag;
  const int nt = nx + 2 * kNFact;
#define WB_XIN(i) x[imin(n - 1, imax(0, (i) - lag))]
  for (int i = 0; i < kNFact; ++i) tmp1[i] = 2 * WB_XIN(0) - WB_XIN(kNFact - i);
  for (int i = kNFact; i < kNFact + nx; ++i) tmp1[i] = WB_XIN(i - kNFact);
  for (int i = kNFact + nx; i < nt; ++i) tmp1[i] = 2 * WB_XIN(nx - 1) - WB_XIN(nx - 2 - (i - (kNFact + nx)));
#undef WB_XIN
  double a[3], b[2];
  decimate_coefficients(r, a, b);
  for (int pass = 0; pass < 2; ++pass) {
    double w0 = 0.0, w1 = 0.0, w2 = 0.0;
    for (int i = 0; i < nt; ++i) {
      const double wt = tmp1[i] + a[0] * w0 + a[1] * w1 + a[2] * w2;
      tmp2[i] = b[0] * wt + b[1] * w0 + b[1] * w1 + b[0] * w2;
      w2 = w1; w1 = w0; w0 = wt;
    }
    for (int i = 0; i < nt; ++i) tmp1[i] = tmp2[nt - i - 1];
  }
  const int nout = (nx - 1) / r + 1;
  const int nbeg = r - r * nout + nx;
  int count = 0;
  for (int i = nbeg; i < nx + kNFact; i += r, ++count) {
    const int k = count - first;
    if (k >= 0 && k < n_out) out[k] = tmp1[i + kNFact - 1];
  }
  return count;
}

// ------------------------------------------------------------------------------ plain FIR
// out(q) = sum_k h[k] * in(q - k), q in [0, q_len[u]); `in` is zero padded on both sides.
// grid (tiles, utterances); same register tiling as the band sweep.
struct FirParams {
  const double *in; size_t in_stride; int in_origin;
  double *out; size_t out_stride; int out_origin;
  const int *base_len; int extra_len;   // q_len[u] = base_len[u] + extra_len
  const double *taps_rev; int ntaps;
};

WB_HD inline size_t fir_plain_smem_bytes(int ntaps) {
  const int seg = 2048 + ntaps + 16;
  return (size_t)(seg + (seg >> 3) + 8) * 8 + (size_t)(ntaps + 8) * 8;
}

// ------------------------------------------------------------------------------ band sweep
struct SweepParams {
  const double *sig; size_t sig_stride; int sig_origin;  // sig[u*stride + origin + m] = s(m), zero padded
  const int *y_len;                                      // [n] samples to filter per utterance
  int n_bands;
  const double *taps_rev; const int *tap_off; const int *ntaps; const int *shift;  // per band
  const double *boundary;                                // per band boundary f0
  double afs;                                            // sampling rate of sig
  double *edges; size_t edge_cap;                        // [(u*nb+b)*4+train][edge_cap] fine edges
  const int *n_frames; int frame_stride; double frame_period;  // frame grid: t_i = i*frame_period/1000
  int mode;                                              // 0 = DIO (candidate + score), 1 = Harvest
  double f0_floor, f0_ceil;
  double *cand; double *score;                           // [(u*nb+b)][frame_stride]
  int max_taps;
  int *status;
};

WB_HD inline size_t sweep_smem_bytes(int max_taps) {
  const int seg = WB_SWEEP_T + max_taps + 16;
  return (size_t)(seg + (seg >> 3) + 8) * 8 + (size_t)(max_taps + 8) * 8 + (size_t)(WB_SWEEP_T + 8) * 8 +
         (size_t)(WB_SWEEP_T / WB_SWEEP_R + 40) * 8;
}

// launchers (wb_f0common.cu)
struct Ctx;
void launch_fir_plain(Ctx *ctx, const FirParams &p, unsigned tiles, unsigned n_utts);
void launch_band_sweep(Ctx *ctx, const SweepParams &p, unsigned n_utts);

}  // namespace wb
