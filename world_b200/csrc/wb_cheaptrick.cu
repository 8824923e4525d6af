// wb_cheaptrick.cu -- K-CT: spectral envelope, one CTA per (utterance, frame).
//
// Replaces the frame loop of CheapTrick() (cheaptrick.cpp:200-229) and its helpers
// (GetWindowedWaveform :112-142, GetPowerSpectrum :64-82, AddInfinitesimalNoise :147-151,
// SmoothingWithRecovery :22-57) plus common.cpp's DCCorrection/LinearSmoothing.  Algorithm card:
// SURVEY.md A1.  Everything between the waveform read and the spectrogram row write stays in
// shared memory: three real FFTs (window spectrum, cepstrum, liftered cepstrum back -- the
// last two are transforms of real even sequences, so the reference's c2r is a forward r2c
// here), the index-order running sum of LinearSmoothing on one thread, log/exp fused in.
//
// HBM traffic per frame (algorithmic): (2h+1)*8 B of waveform (L2 hits after the first frame of
// a hop), 16 B of f0/time, (2h+1 + fft/2+1)*4 B of materialised draws, (fft/2+1)*8 B written.
#include "wb_internal.h"
#include "wb_spectral.cuh"
#include <stdlib.h>

namespace wb {

struct CtParams {
  const double *x; const int *x_len; int x_stride;
  const double *time_axis; const double *f0; const int *f_len; int f_stride;
  int fs, fft_size, lg_fft;
  double q1, f0_floor;
  const unsigned *draws; size_t draw_stride; const unsigned *draw_off;
  double *out;            // [n][f_stride][fft/2+1]
  const double2 *tw;
  int *status;
  // coded tail (ct_coded_kernel): mel interpolation + DCT of the log envelope, CodeSpectralEnvelope (codec.cpp:279-295)
  int c_dims; const int *c_idx; const double *c_frac; const double2 *c_weight; double c_norm;
  double *c_out;          // [n][f_stride][c_dims]
};

WB_DEV double ct_effective_f0(double f0, double f0_floor) {
  return f0 <= f0_floor ? 500.0 : f0;  // kDefaultF0, cheaptrick.cpp:218
}

// draws consumed by frame = window samples + spectrum bins
WB_KERNEL_PLAIN ct_count_kernel(const double *__restrict__ f0, const int *__restrict__ f_len,
                                int f_stride, int n_utts, int fs, int fft_size, double f0_floor,
                                unsigned *__restrict__ counts) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (long long)n_utts * f_stride) return;
  const int u = (int)(g / f_stride), i = (int)(g % f_stride);
  unsigned c = 0;
  if (i < f_len[u]) {
    const double f = ct_effective_f0(f0[g], f0_floor);
    const int h = round_half_away(1.5 * fs / f);
    c = (unsigned)(2 * h + 1) + (unsigned)(fft_size / 2 + 1);
  }
  counts[g] = c;
}

// Shared memory: two padded complex buffers of fft_size/2 slots (the FFT ping-pongs between them, wb_fft.cuh)
// + the reduction scratch.  Between transforms the buffer that does not hold FFT input serves as the plain
// double array of the phase (window, power spectrum, smoothing scratch), so nothing else is needed.
WB_HD inline size_t ct_smem_bytes(int fft_size) {
  return (size_t)2 * WB_FPAD_SLOTS(fft_size / 2) * sizeof(double2) + WB_RED_DOUBLES * sizeof(double);
}

template <bool kCoded>
WB_DEV void ct_frame_body(const CtParams &p) {
  WB_DYN_SMEM(double2, smem2);
  const int tid = WB_TID, nth = WB_NTH;
  const int u = blockIdx.y, i = blockIdx.x;
  if (i >= p.f_len[u]) return;
  const int N = p.fft_size, half = N / 2;
  const int slots = WB_FPAD_SLOTS(half);
  double2 *A = smem2, *B = smem2 + slots;
  double *red = reinterpret_cast<double *>(B + slots);   // WB_RED_DOUBLES

  const size_t fidx = (size_t)u * p.f_stride + i;
  const double f = ct_effective_f0(p.f0[fidx], p.f0_floor);
  const double t = p.time_axis[fidx];
  const int fs = p.fs;
  const int x_len = p.x_len[u];
  const double *x = p.x + (size_t)u * p.x_stride;
  const unsigned *draw = p.draws + (size_t)u * p.draw_stride + p.draw_off[fidx];

  const int h = round_half_away(1.5 * fs / f);
  const int nwin = 2 * h + 1;
  if (nwin > N) {  // f0 below the floor implied by fft_size: undefined in the reference
    if (tid == 0) atomicOr_status(p.status, 1);
    return;
  }
  const int origin = round_half_away(t * fs + 0.001);

  // ---- F0-adaptive Hanning window, normalised to unit energy (cheaptrick.cpp:97-106)
  double *za = reinterpret_cast<double *>(A);   // packed FFT input: sample e at rpad(e)
  double *wv = reinterpret_cast<double *>(B);   // window values
  double sq = 0.0;
  for (int j = tid; j < nwin; j += nth) {
    const double pos = (j - h) / 1.5 / fs;
    const double w = 0.5 * cos_small(kPi * pos * f) + 0.5;
    wv[j] = w;
    sq += w * w;
  }
  // first round's sample and draw requested before the reduction's barriers, every later round's one round ahead
  double x_next = 0.0;
  unsigned d_next = 0u;
  if (tid < nwin) { x_next = x[imin(x_len - 1, imax(0, origin + tid - h))]; d_next = draw[tid]; }
  const double norm = sqrt(block_sum(sq, red));
  // ---- windowed waveform + 1e-12 * randn, weighted-mean removal (:126-137)
  double s1 = 0.0, s2 = 0.0;
  for (int j = tid; j < nwin; j += nth) {
    const double x_now = x_next;
    const unsigned d_now = d_next;
    if (j + nth < nwin) { x_next = x[imin(x_len - 1, imax(0, origin + j + nth - h))]; d_next = draw[j + nth]; }
    const double w = wv[j] / norm;
    wv[j] = w;
    const double v = x_now * w + randn_value(d_now) * kTiny;
    za[rpad(j)] = v;
    s1 += v;
    s2 += w;
  }
  block_sum2(s1, s2, red);
  const double coef = s1 / s2;
  for (int j = tid; j < N; j += nth) za[rpad(j)] = (j < nwin) ? za[rpad(j)] - wv[j] * coef : 0.0;
  WB_SYNC();

  // ---- power spectrum (:71-78): the unpack of the real FFT squares straight into the idle buffer
  const int lgc = p.lg_fft - 1;
  double2 *z = sfft_forward(A, B, lgc, p.tw);
  double2 *o = (z == A) ? B : A;
  double *pw = reinterpret_cast<double *>(o);
  rfft_unpack(z, p.lg_fft, p.tw, [&](int k, double2 c) { pw[k] = c.x * c.x + c.y * c.y; });
  WB_SYNC();
  // the draws of the +eps noise below are requested now: they arrive while the smoothing (barriers, one thread's
  // index-order running sum) keeps the CTA waiting anyway.  (half + 1 <= 5 * 128 up to fft_size 1024; beyond that
  // the tail of the loop below loads as before.)
  unsigned dpre[5];
#pragma unroll
  for (int q = 0; q < 5; ++q) { const int k = tid + q * nth; dpre[q] = (k <= half) ? draw[nwin + k] : 0u; }
  dc_correction<true>(pw, f, fs, N, reinterpret_cast<double *>(z));
  if (!linear_smoothing<true>(pw, f * 2.0 / 3.0, fs, N, pw, reinterpret_cast<double *>(z), red)) {
    if (tid == 0) atomicOr_status(p.status, 2);
    return;
  }
  // ---- + |randn| * eps (:147-151), log, mirrored to an even sequence of N (:39-42) = input of the next FFT
  {
    double *zin = reinterpret_cast<double *>(z);
#pragma unroll
    for (int q = 0; q < 5; ++q) {
      const int k = tid + q * nth;
      if (k <= half) {
        const double v = pw[k] + fabs(randn_value(dpre[q])) * kEps;
        const double l = log(v);
        zin[rpad(k)] = l;
        if (k > 0 && k < half) zin[rpad(N - k)] = l;
      }
    }
    for (int k = tid + 5 * nth; k <= half; k += nth) {
      const double v = pw[k] + fabs(randn_value(draw[nwin + k])) * kEps;
      const double l = log(v);
      zin[rpad(k)] = l;
      if (k > 0 && k < half) zin[rpad(N - k)] = l;
    }
  }
  WB_SYNC();
  double2 *z2 = sfft_forward(z, o, lgc, p.tw);
  double2 *o2 = (z2 == A) ? B : A;
  // ---- liftering in the cepstrum domain (:28-37, 45-49); cos(2a) = 1 - 2 sin(a)^2 saves the cosine
  {
    double *zin = reinterpret_cast<double *>(o2);
    const double q1 = p.q1;
    // the lifters only scale floating-point values: quefrency through reciprocals, sin(pi a) by sinpi, 1/N exactly
    const double f_over_fs = f / fs, inv_n = 1.0 / N;
    rfft_unpack(z2, p.lg_fft, p.tw, [&](int k, double2 c) {
      double sl = 1.0, cl = (1.0 - 2.0 * q1) + 2.0 * q1;
      if (k > 0) {
        const double a = f_over_fs * k;          // f * quefrency
        const double sn = sinpi(a);
        sl = sn / (kPi * a);
        cl = (1.0 - 2.0 * q1) + 2.0 * q1 * (1.0 - 2.0 * sn * sn);
      }
      const double v = c.x * sl * cl * inv_n;
      zin[rpad(k)] = v;
      if (k > 0 && k < half) zin[rpad(N - k)] = v;
    });
  }
  WB_SYNC();
  double2 *z3 = sfft_forward(o2, z2, lgc, p.tw);
  if (!kCoded) {
    double *row = p.out + fidx * (size_t)(half + 1);
    rfft_unpack(z3, p.lg_fft, p.tw, [&](int k, double2 c) { row[k] = exp(c.x); });
    return;
  }
  // ---- coded row: CodeSpectralEnvelope takes the log of the envelope (codec.cpp:288-289) -- the transform's real
  // part as it stands, the exp / log pair of the unfused path cancels -- interpolates it onto the mel grid (:121-122)
  // and runs DCTForCodec (:73-89) as one real FFT of fft_size / 2 points
  double2 *o3 = (z3 == A) ? B : A;
  double *lgs = reinterpret_cast<double *>(o3);   // half + 1 plain doubles
  rfft_unpack(z3, p.lg_fft, p.tw, [&](int k, double2 c) { lgs[k] = c.x; });
  WB_SYNC();
  {
    double *zin = reinterpret_cast<double *>(z3);
    const int bias = half / 2;
    for (int j = tid; j < bias; j += nth) {        // even samples ascending, odd samples descending (:77-81)
      const int a = 2 * j, b = half - 2 * j - 1;
      const int ka = __ldg(p.c_idx + a), kb = __ldg(p.c_idx + b);
      zin[rpad(j)] = lgs[ka] + __ldg(p.c_frac + a) * (lgs[ka + 1] - lgs[ka]);
      zin[rpad(j + bias)] = lgs[kb] + __ldg(p.c_frac + b) * (lgs[kb + 1] - lgs[kb]);
    }
  }
  WB_SYNC();
  const double2 *zc = sfft_forward(z3, o3, lgc - 1, p.tw);
  double *crow = p.c_out + fidx * (size_t)p.c_dims;
  const int dims = p.c_dims;
  const double cn = p.c_norm;
  rfft_unpack(zc, lgc, p.tw, [&](int d, double2 X) {     // :85-88
    if (d < dims) {
      const double2 w = __ldg(p.c_weight + d);
      crow[d] = (X.x * w.x - X.y * w.y) / cn;
    }
  });
}

#ifndef WB_EMU
__global__ void __launch_bounds__(128, 8) ct_frame_kernel(CtParams p) { ct_frame_body<false>(p); }
__global__ void __launch_bounds__(128, 8) ct_coded_kernel(CtParams p) { ct_frame_body<true>(p); }
#else
void ct_frame_kernel(CtParams p) { ct_frame_body<false>(p); }
void ct_coded_kernel(CtParams p) { ct_frame_body<true>(p); }
#endif

int cheaptrick_run(Ctx *ctx, const Batch &b, double q1, int fft_size, double *spectrogram,
                   const CodecTables *coded, double *coded_out) {
  if (b.n <= 0 || b.max_f_len <= 0) return 0;
  int lg = 0;
  while ((1 << lg) < fft_size) ++lg;
  if ((1 << lg) != fft_size || fft_size < 16 || fft_size > WB_TW_N) {
    ctx->last_error = "CheapTrick: fft_size must be a power of two in [16, 8192]";
    return 3;
  }
  const double f0_floor = 3.0 * b.fs / (fft_size - 3.0);  // GetF0FloorForCheapTrick, :196-198
  const int bins = fft_size / 2 + 1;
  // a frame draws at most (fft_size - 2) + bins numbers (2h+1 <= fft_size - 2 above the floor)
  const size_t max_per_frame = (size_t)fft_size + bins;
  const size_t draw_stride_full = max_per_frame * (size_t)b.max_f_len;
  // utterances per chunk so that the draw scratch fits the budget
  const size_t per_utt_bytes = draw_stride_full * 4 + (size_t)b.f_stride * 8 + 64;
  int chunk = balanced_chunk(imin(b.n, 65535), (int)dmin(65535.0, (double)ctx->scratch_budget / (double)per_utt_bytes));
  const size_t smem = ct_smem_bytes(fft_size);
#ifndef WB_EMU
  cudaFuncSetAttribute(ct_frame_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(ct_coded_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
#endif
  const size_t n_tab = coded ? coded->idx.size() : 0, n_w = coded ? coded->weight.size() : 0;
  for (int u0 = 0; u0 < b.n; u0 += chunk) {
    const int n = imin(chunk, b.n - u0);
    ArenaPlan plan;
    const size_t o_counts = plan.add((size_t)n * b.f_stride * 4);
    const size_t o_offs = plan.add((size_t)n * b.f_stride * 4);
    const size_t o_totals = plan.add((size_t)n * 4);
    const size_t o_draws = plan.add((size_t)n * draw_stride_full * 4);
    const size_t o_cidx = plan.add(n_tab * 4 + 4), o_cfrac = plan.add(n_tab * 8 + 8), o_cw = plan.add(n_w * 16 + 16);
    unsigned char *blk = arena_block(ctx, plan.total);
    if (!blk) return 2;
    unsigned *counts = (unsigned *)(blk + o_counts), *offs = (unsigned *)(blk + o_offs);
    unsigned *totals = (unsigned *)(blk + o_totals), *draws = (unsigned *)(blk + o_draws);
    const double *f0 = b.f0 + (size_t)u0 * b.f_stride;
    const int *f_len = b.f_len + u0;
    const long long total_slots = (long long)n * b.f_stride;
    WB_LAUNCH_FLAT(ct_count_kernel, dim3((unsigned)((total_slots + 255) / 256)), 256, 0, ctx->stream,
                   f0, f_len, b.f_stride, n, b.fs, fft_size, f0_floor, counts);
    scan_counts(ctx, counts, f_len, b.f_stride, nullptr, offs, totals, n);
    rng_fill(ctx, totals, draws, draw_stride_full, draw_stride_full, n);
    CtParams p;
    p.x = b.x + (size_t)u0 * b.x_stride; p.x_len = b.x_len + u0; p.x_stride = b.x_stride;
    p.time_axis = b.time_axis + (size_t)u0 * b.f_stride; p.f0 = f0; p.f_len = f_len;
    p.f_stride = b.f_stride; p.fs = b.fs; p.fft_size = fft_size; p.lg_fft = lg;
    p.q1 = q1; p.f0_floor = f0_floor;
    p.draws = draws; p.draw_stride = draw_stride_full; p.draw_off = offs;
    p.out = spectrogram ? spectrogram + (size_t)u0 * b.f_stride * bins : nullptr;
    p.tw = ctx->twiddle; p.status = ctx->status_dev;
    p.c_dims = 0; p.c_idx = nullptr; p.c_frac = nullptr; p.c_weight = nullptr; p.c_norm = 1.0; p.c_out = nullptr;
    int ct_threads = 128;
    if (const char *e = getenv("WB_CT_THREADS")) ct_threads = atoi(e);
    if (coded) {
      int rc = dev_memcpy_h2d(ctx, blk + o_cidx, coded->idx.data(), n_tab * 4);
      if (!rc) rc = dev_memcpy_h2d(ctx, blk + o_cfrac, coded->frac.data(), n_tab * 8);
      if (!rc) rc = dev_memcpy_h2d(ctx, blk + o_cw, coded->weight.data(), n_w * 16);
      if (rc) return rc;
      p.c_dims = coded->dims; p.c_idx = (const int *)(blk + o_cidx); p.c_frac = (const double *)(blk + o_cfrac);
      p.c_weight = (const double2 *)(blk + o_cw); p.c_norm = coded->norm;
      p.c_out = coded_out + (size_t)u0 * b.f_stride * coded->dims;
      WB_LAUNCH_COOP(ct_coded_kernel, dim3((unsigned)b.max_f_len, (unsigned)n), ct_threads, smem, ctx->stream, p);
    } else
    WB_LAUNCH_COOP(ct_frame_kernel, dim3((unsigned)b.max_f_len, (unsigned)n), ct_threads, smem, ctx->stream, p);
    int rc = dev_check(ctx, "cheaptrick");
    if (rc) return rc;
  }
  return 0;
}

}  // namespace wb
