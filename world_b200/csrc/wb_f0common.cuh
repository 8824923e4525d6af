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

WB_KERNEL(256, 3) fir_plain_kernel(FirParams p) {
  WB_DYN_SMEM(double, smem);
  const int tid = WB_TID, nth = WB_NTH;
  const int T = 2048, R = 8, G = T / R;
  const int u = blockIdx.y, n0 = blockIdx.x * T;
  const int qlen = p.base_len[u] + p.extra_len;
  if (n0 >= qlen) return;
  const int ntaps = p.ntaps;
  const int seg_len = T + ntaps - 1;
  const int seg_cap = seg_len + 16;
  double *seg = smem;
  double *hrev = seg + (seg_cap + (seg_cap >> 3) + 8);
  const double *in = p.in + (size_t)u * p.in_stride + p.in_origin;
  double *out = p.out + (size_t)u * p.out_stride + p.out_origin;
  for (int j = tid; j < ntaps + 8; j += nth) hrev[j] = j < ntaps ? __ldg(&p.taps_rev[j]) : 0.0;
  const int m0 = n0 - ntaps + 1;
  for (int i = tid; i < seg_len + 8; i += nth) seg[pad8(i)] = (i < seg_len) ? in[m0 + i] : 0.0;
  WB_SYNC();
  for (int g = tid; g < G; g += nth) {
    const int base = R * g;
    double acc[8], win[8];
#pragma unroll
    for (int r = 0; r < R; ++r) { acc[r] = 0.0; win[r] = seg[pad8(base + r)]; }
    for (int j0 = 0; j0 < ntaps; j0 += R) {
#pragma unroll
      for (int jj = 0; jj < R; ++jj) {
        const double hj = hrev[j0 + jj];
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = fma(hj, win[(r + jj) & (R - 1)], acc[r]);
        win[jj] = seg[pad8(base + R + j0 + jj)];
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r)
      if (n0 + base + r < qlen) out[n0 + base + r] = acc[r];
  }
}

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

// interp1 (matlabfunctions.cpp:157-176) of one event train at time t; edges = fine edge
// positions (n_edges of them), sample (x, y) pairs are (location, interval) of consecutive edges.
WB_DEV double train_interp(const double *e, int n_int, double afs, double t) {
  // k = #{j : loc[j] <= t} clamped to [1, n_int-1]; loc[j] = (e[j] + e[j+1]) / 2 / afs
  int lo = 0, hi = n_int;  // first j with loc[j] > t
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    const double loc = (e[mid] + e[mid + 1]) / 2.0 / afs;
    if (loc <= t) lo = mid + 1; else hi = mid;
  }
  const int k = imin(n_int - 1, imax(1, lo));
  const double e0 = e[k - 1], e1 = e[k], e2 = e[k + 1];
  const double x0 = (e0 + e1) / 2.0 / afs, x1 = (e1 + e2) / 2.0 / afs;
  const double y0 = afs / (e1 - e0), y1 = afs / (e2 - e1);
  const double s = (t - x0) / (x1 - x0);
  return y0 + s * (y1 - y0);
}

// Exclusive scan of G packed counters (4 x 16 bit) held in shared memory, in place; adds the
// running totals in `carry` (also packed) and returns the new running total to every thread.
// CUDA path: requires blockDim.x == G (one counter per thread).
WB_DEV unsigned long long scan_packed(unsigned long long *c, int G, unsigned long long carry,
                                      unsigned long long *warp_tot /* >= 33 */) {
#ifdef WB_EMU
  (void)warp_tot;
  unsigned long long run = carry;
  for (int g = 0; g < G; ++g) { const unsigned long long v = c[g]; c[g] = run; run += v; }
  return run;
#else
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, nw = G >> 5;
  const unsigned long long v = c[tid];
  unsigned long long inc = v;
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_tot[w] = inc;
  __syncthreads();
  unsigned long long base = carry, all = carry;
  for (int i = 0; i < nw; ++i) { const unsigned long long t = warp_tot[i]; if (i < w) base += t; all += t; }
  c[tid] = base + inc - v;
  __syncthreads();
  return all;
#endif
}

WB_KERNEL(WB_SWEEP_THREADS, 3) band_sweep_kernel(SweepParams p) {
  WB_DYN_SMEM(double, smem);
  const int tid = WB_TID, nth = WB_NTH;
  const int b = blockIdx.x, u = blockIdx.y;
  const int T = WB_SWEEP_T, R = WB_SWEEP_R, G = T / R;
  const int ntaps = p.ntaps[b], shift = p.shift[b];
  const int seg_len = T + ntaps - 1;
  const int seg_cap = T + p.max_taps + 16;
  double *seg = smem;                                   // padded: pad8(seg_cap)
  double *hrev = seg + (seg_cap + (seg_cap >> 3) + 8);  // max_taps + 8
  double *st = hrev + (p.max_taps + 8);                 // T + 8: [0..1] carry, [2..T+2) this tile
  unsigned long long *cnt = reinterpret_cast<unsigned long long *>(st + (T + 8));  // G + 40

  const int ylen = p.y_len[u];
  const double *sig = p.sig + (size_t)u * p.sig_stride + p.sig_origin;
  double *edges = p.edges + ((size_t)u * p.n_bands + b) * 4 * p.edge_cap;
  const int cap = (int)p.edge_cap;
  for (int j = tid; j < ntaps; j += nth) hrev[j] = __ldg(&p.taps_rev[p.tap_off[b] + j]);
  for (int j = ntaps + tid; j < ntaps + 8; j += nth) hrev[j] = 0.0;
  if (tid == 0) { st[0] = 0.0; st[1] = 0.0; }
  int tot[4] = {0, 0, 0, 0};  // running event counts per train (identical in every thread)
  WB_SYNC();

  // Tile k produces outputs n0..n0+T-1 into st[2..]; events are detected for positions
  // i = n0-2 .. n0+T-3 (they need s[i], s[i+1], s[i+2]); the last two outputs carry over.
  for (int n0 = 0; n0 < ylen + 2; n0 += T) {
    const int m0 = n0 + shift - ntaps + 1;  // seg[i] = s(m0 + i)
    for (int i = tid; i < seg_len + 8; i += nth) seg[pad8(i)] = (i < seg_len) ? sig[m0 + i] : 0.0;
    WB_SYNC();
    for (int g = tid; g < G; g += nth) {
      const int base = R * g;
      double acc[WB_SWEEP_R], win[WB_SWEEP_R];
#pragma unroll
      for (int r = 0; r < R; ++r) { acc[r] = 0.0; win[r] = seg[pad8(base + r)]; }
      for (int j0 = 0; j0 < ntaps; j0 += R) {
#pragma unroll
        for (int jj = 0; jj < R; ++jj) {
          const double hj = hrev[j0 + jj];  // zero beyond ntaps
#pragma unroll
          for (int r = 0; r < R; ++r) acc[r] = fma(hj, win[(r + jj) & (R - 1)], acc[r]);
          win[jj] = seg[pad8(base + R + j0 + jj)];
        }
      }
#pragma unroll
      for (int r = 0; r < R; ++r) st[2 + base + r] = acc[r];
    }
    WB_SYNC();
    // train 0: s[i] > 0 >= s[i+1]   train 1: s[i] < 0 <= s[i+1]          (i >= 0, i+1 <= ylen-1)
    // train 2: d[i] > 0 >= d[i+1]   train 3: d[i] < 0 <= d[i+1], d[i] = s[i+1]-s[i]  (i+1 <= ylen-2)
    for (int g = tid; g < G; g += nth) {
      unsigned long long c = 0ull;
      const int base = R * g;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int i = n0 - 2 + base + r;
        const double a = st[base + r], bb = st[base + r + 1], cc = st[base + r + 2];
        const double d0 = bb - a, d1 = cc - bb;
        if (i >= 0 && i + 1 <= ylen - 1) {
          c += (0.0 < a && bb <= 0.0) ? 1ull : 0ull;
          c += (a < 0.0 && 0.0 <= bb) ? (1ull << 16) : 0ull;
        }
        if (i >= 0 && i + 1 <= ylen - 2) {
          c += (0.0 < d0 && d1 <= 0.0) ? (1ull << 32) : 0ull;
          c += (d0 < 0.0 && 0.0 <= d1) ? (1ull << 48) : 0ull;
        }
      }
      cnt[g] = c;
    }
    WB_SYNC();
    const unsigned long long tile_total = scan_packed(cnt, G, 0ull, cnt + G + 4);  // <= 2048 each: fits 16 bit
    for (int g = tid; g < G; g += nth) {
      const unsigned long long o = cnt[g];
      int o0 = tot[0] + (int)(o & 0xffffull), o1 = tot[1] + (int)((o >> 16) & 0xffffull);
      int o2 = tot[2] + (int)((o >> 32) & 0xffffull), o3 = tot[3] + (int)((o >> 48) & 0xffffull);
      const int base = R * g;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int i = n0 - 2 + base + r;
        const double a = st[base + r], bb = st[base + r + 1], cc = st[base + r + 2];
        const double d0 = bb - a, d1 = cc - bb;
        const double e = (double)(i + 1);
        if (i >= 0 && i + 1 <= ylen - 1) {
          if (0.0 < a && bb <= 0.0) { if (o0 < cap) edges[o0] = e - a / (bb - a); ++o0; }
          if (a < 0.0 && 0.0 <= bb) { if (o1 < cap) edges[(size_t)cap + o1] = e - a / (bb - a); ++o1; }
        }
        if (i >= 0 && i + 1 <= ylen - 2) {
          if (0.0 < d0 && d1 <= 0.0) { if (o2 < cap) edges[2 * (size_t)cap + o2] = e - d0 / (d1 - d0); ++o2; }
          if (d0 < 0.0 && 0.0 <= d1) { if (o3 < cap) edges[3 * (size_t)cap + o3] = e - d0 / (d1 - d0); ++o3; }
        }
      }
    }
    tot[0] += (int)(tile_total & 0xffffull); tot[1] += (int)((tile_total >> 16) & 0xffffull);
    tot[2] += (int)((tile_total >> 32) & 0xffffull); tot[3] += (int)((tile_total >> 48) & 0xffffull);
    WB_SYNC();
    if (tid == 0) { st[0] = st[T]; st[1] = st[T + 1]; }
    WB_SYNC();
  }
#ifndef WB_EMU
  __threadfence_block();
#endif
  WB_SYNC();
  // ---- candidates on the frame grid
  const int nf = p.n_frames[u];
  double *cand = p.cand + ((size_t)u * p.n_bands + b) * p.frame_stride;
  double *score = p.score ? p.score + ((size_t)u * p.n_bands + b) * p.frame_stride : nullptr;
  int ni[4];
  bool ok = true;
  for (int q = 0; q < 4; ++q) {
    if (tot[q] > cap) { if (tid == 0) atomicOr_status(p.status, 4); ok = false; }
    ni[q] = tot[q] < 2 ? 0 : tot[q] - 1;  // ZeroCrossingEngine returns count-1 (0 if count<2)
    if (ni[q] - 2 <= 0) ok = false;       // CheckEvent(n - 2), dio.cpp:475-484
  }
  const double bf = p.boundary[b];
  for (int i = tid; i < nf; i += nth) {
    double c = 0.0, sc = 100000.0;  // kMaximumValue
    if (ok) {
      const double t = i * p.frame_period / 1000.0;
      const double v0 = train_interp(edges, ni[0], p.afs, t);
      const double v1 = train_interp(edges + cap, ni[1], p.afs, t);
      const double v2 = train_interp(edges + 2 * (size_t)cap, ni[2], p.afs, t);
      const double v3 = train_interp(edges + 3 * (size_t)cap, ni[3], p.afs, t);
      c = (v0 + v1 + v2 + v3) / 4.0;
      if (p.mode == 0) {
        sc = sqrt(((v0 - c) * (v0 - c) + (v1 - c) * (v1 - c) + (v2 - c) * (v2 - c) + (v3 - c) * (v3 - c)) / 3.0);
        if (c > bf || c < bf / 2.0 || c > p.f0_ceil || c < p.f0_floor) { c = 0.0; sc = 100000.0; }
      } else {
        if (c > bf * 1.1 || c < bf * 0.9 || c > p.f0_ceil || c < p.f0_floor) c = 0.0;
      }
    }
    cand[i] = c;
    if (score) score[i] = sc / (c + kTiny);  // dio.cpp:562-566
  }
}

}  // namespace wb
