// wb_host.cu -- host-pointer entry points built on the device-pointer ABI:
//   * world_b200_analyze_host(): {Dio+StoneMask | Harvest} -> CheapTrick -> D4C for N host
//     waveforms, upload / compute / download pipelined over utterance chunks on three streams;
//   * the reference's own single-utterance functions (Dio, Harvest, StoneMask, CheapTrick, D4C;
//     src/world/*.h) as n_utts = 1 batches on a lazily created process-wide context, so existing
//     callers relink unchanged.  They keep the reference's `void` signature; failures are
//     reported on stderr and leave the outputs zero-filled.
#include "wb_internal.h"
#include "../../include/world_b200.h"
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include <mutex>

using namespace wb;

namespace {

struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
};

int ensure(Ctx *ctx, DevBuf *b, size_t bytes) {
  if (bytes <= b->cap) return 0;
  dev_free(b->p);
  b->p = dev_malloc(ctx, bytes);
  b->cap = b->p ? bytes : 0;
  return b->p ? 0 : WORLD_B200_ENOMEM;
}

Ctx *ctx_of(WorldB200 *h) { return reinterpret_cast<Ctx *>(h); }  // Ctx is the first member

}  // namespace

namespace {

// One pipeline for both host entry points.  nbit == 0: x holds doubles; otherwise little-endian PCM that
// is widened on the device (row f3).  dims == 0: the full spectrogram / aperiodicity rows go back to the
// host; dims > 0: they stay on the device and only their coded rows (row f2) are downloaded.
int analyze_pipeline(WorldB200 *h, const void *x, int nbit, int n_utts, int x_stride, const int *x_lengths, int fs,
                     const WorldB200AnalysisOption *opt, int dims, double *time_axis, double *f0, int f0_stride,
                     double *out_sp, double *out_ap) {
  Ctx *ctx = ctx_of(h);
  const int bins = opt->cheaptrick.fft_size / 2 + 1;
  const int n_ap = GetNumberOfAperiodicities(fs);
  const size_t in_bytes = nbit ? (size_t)(nbit / 8) : 8;
  const size_t sp_row = dims ? (size_t)dims : (size_t)bins, ap_row = dims ? (size_t)n_ap : (size_t)bins;
  const double frame_period =
      opt->f0_method == WORLD_B200_F0_HARVEST ? opt->harvest.frame_period : opt->dio.frame_period;
  // chunk so that two sets of device buffers (double buffering) stay within ~1/3 of the budget
  const size_t per_utt = (size_t)x_stride * (8 + (nbit ? in_bytes : 0)) +
                         (size_t)f0_stride * (16 + 2 * (size_t)bins * 8 + (dims ? (sp_row + ap_row) * 8 : 0));
  int chunk = (int)dmax(1.0, dmin((double)n_utts, (double)(ctx->scratch_budget / 3) / (double)(2 * per_utt)));
  // small chunks keep the un-overlapped tail (download of the last chunk) short; 96 utterances still
  // fill the GPU (the per-utterance kernels see 96 CTAs, the frame kernels ~200 k)
  int cap = 96;
  if (const char *e = getenv("WB_HOST_CHUNK")) cap = atoi(e) > 0 ? atoi(e) : cap;
  if (chunk > cap) chunk = cap;

#ifndef WB_EMU
  cudaStream_t s_compute = ctx->stream, s_in = nullptr, s_out = nullptr;
  if (cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking) != cudaSuccess) {
    ctx->last_error = "cudaStreamCreate failed";
    return WORLD_B200_ECUDA;
  }
  cudaEvent_t ev_in[2], ev_done[2], ev_out[2];
  for (int i = 0; i < 2; ++i) {
    cudaEventCreateWithFlags(&ev_in[i], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ev_done[i], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ev_out[i], cudaEventDisableTiming);
  }
#endif
  DevBuf din[2], dx[2], dt[2], df[2], dsp[2], dap[2], dcs[2], dca[2];
  int rc = 0;
  for (int i = 0; i < 2 && !rc; ++i) {
    rc = ensure(ctx, &din[i], (size_t)chunk * x_stride * in_bytes);
    if (!rc && nbit) rc = ensure(ctx, &dx[i], (size_t)chunk * x_stride * 8);
    if (!rc) rc = ensure(ctx, &dt[i], (size_t)chunk * f0_stride * 8);
    if (!rc) rc = ensure(ctx, &df[i], (size_t)chunk * f0_stride * 8);
    if (!rc && out_sp) rc = ensure(ctx, &dsp[i], (size_t)chunk * f0_stride * bins * 8);
    if (!rc && out_ap) rc = ensure(ctx, &dap[i], (size_t)chunk * f0_stride * bins * 8);
    if (!rc && dims && out_sp) rc = ensure(ctx, &dcs[i], (size_t)chunk * f0_stride * sp_row * 8);
    if (!rc && dims && out_ap && n_ap > 0) rc = ensure(ctx, &dca[i], (size_t)chunk * f0_stride * ap_row * 8);
  }
  std::vector<int> flen(n_utts > 0 ? n_utts : 1);
  for (int i = 0; i < n_utts && !rc; ++i) {
    flen[i] = world_b200_frames(fs, x_lengths ? x_lengths[i] : x_stride, frame_period);
    if (flen[i] > f0_stride) { ctx->last_error = "f0_stride too small"; rc = WORLD_B200_EINVAL; }
  }
  int it = 0;
  for (int u0 = 0; u0 < n_utts && !rc; u0 += chunk, ++it) {
    const int n = imin(chunk, n_utts - u0);
    const int s = it & 1;
    const int *xl = x_lengths ? x_lengths + u0 : nullptr;
    const int *fl = flen.data() + u0;
    const size_t fsz = (size_t)n * f0_stride;
    const unsigned char *src = (const unsigned char *)x + (size_t)u0 * x_stride * in_bytes;
#ifndef WB_EMU
    // buffers of slot s are free once the download issued two iterations ago has finished
    if (it >= 2) cudaStreamWaitEvent(s_in, ev_out[s], 0);
    cudaMemcpyAsync(din[s].p, src, (size_t)n * x_stride * in_bytes, cudaMemcpyHostToDevice, s_in);
    cudaEventRecord(ev_in[s], s_in);
    cudaStreamWaitEvent(s_compute, ev_in[s], 0);
    if (it >= 2) cudaStreamWaitEvent(s_compute, ev_out[s], 0);
#else
    memcpy(din[s].p, src, (size_t)n * x_stride * in_bytes);
#endif
    if (nbit) rc = world_b200_pcm_to_double_batch(h, din[s].p, nbit, n, x_stride, xl, (double *)dx[s].p);
    if (rc) break;
    dev_memset(ctx, dt[s].p, 0, fsz * 8);
    dev_memset(ctx, df[s].p, 0, fsz * 8);
    // whole padded rows are downloaded: frames beyond an utterance's length read as zero on the host
    bool ragged = false;
    for (int i = 0; i < n; ++i) ragged = ragged || fl[i] != f0_stride;
    if (ragged) {
      if (!dims && out_sp) dev_memset(ctx, dsp[s].p, 0, fsz * bins * 8);
      if (!dims && out_ap) dev_memset(ctx, dap[s].p, 0, fsz * bins * 8);
      if (dims && out_sp) dev_memset(ctx, dcs[s].p, 0, fsz * sp_row * 8);
      if (dims && out_ap && n_ap > 0) dev_memset(ctx, dca[s].p, 0, fsz * ap_row * 8);
    }
    const double *xd = (const double *)(nbit ? dx[s].p : din[s].p);
    double *td = (double *)dt[s].p, *fd = (double *)df[s].p;
    if (opt->f0_method == WORLD_B200_F0_HARVEST) {
      rc = world_b200_harvest_batch(h, xd, n, x_stride, xl, fs, &opt->harvest, td, fd, f0_stride);
    } else {
      rc = world_b200_dio_batch(h, xd, n, x_stride, xl, fs, &opt->dio, td, fd, f0_stride);
      if (!rc) rc = world_b200_stonemask_batch(h, xd, n, x_stride, xl, fs, td, fd, fl, f0_stride, fd);
    }
    if (!rc && out_sp)
      rc = world_b200_cheaptrick_batch(h, xd, n, x_stride, xl, fs, td, fd, fl, f0_stride, &opt->cheaptrick,
                                       (double *)dsp[s].p);
    if (!rc && out_ap)
      rc = world_b200_d4c_batch(h, xd, n, x_stride, xl, fs, td, fd, fl, f0_stride, opt->cheaptrick.fft_size,
                                &opt->d4c, (double *)dap[s].p);
    if (!rc && dims && out_sp)
      rc = world_b200_code_spectral_envelope_batch(h, (const double *)dsp[s].p, n, fl, f0_stride, fs,
                                                   opt->cheaptrick.fft_size, dims, (double *)dcs[s].p);
    if (!rc && dims && out_ap && n_ap > 0)
      rc = world_b200_code_aperiodicity_batch(h, (const double *)dap[s].p, n, fl, f0_stride, fs,
                                              opt->cheaptrick.fft_size, (double *)dca[s].p);
    if (rc) break;
    const void *sp_src = dims ? dcs[s].p : dsp[s].p, *ap_src = dims ? dca[s].p : dap[s].p;
    const bool get_ap = out_ap && (!dims || n_ap > 0);
#ifndef WB_EMU
    cudaEventRecord(ev_done[s], s_compute);
    cudaStreamWaitEvent(s_out, ev_done[s], 0);
    if (time_axis) cudaMemcpyAsync(time_axis + (size_t)u0 * f0_stride, td, fsz * 8, cudaMemcpyDeviceToHost, s_out);
    if (f0) cudaMemcpyAsync(f0 + (size_t)u0 * f0_stride, fd, fsz * 8, cudaMemcpyDeviceToHost, s_out);
    if (out_sp)
      cudaMemcpyAsync(out_sp + (size_t)u0 * f0_stride * sp_row, sp_src, fsz * sp_row * 8, cudaMemcpyDeviceToHost, s_out);
    if (get_ap)
      cudaMemcpyAsync(out_ap + (size_t)u0 * f0_stride * ap_row, ap_src, fsz * ap_row * 8, cudaMemcpyDeviceToHost, s_out);
    cudaEventRecord(ev_out[s], s_out);
#else
    if (time_axis) memcpy(time_axis + (size_t)u0 * f0_stride, td, fsz * 8);
    if (f0) memcpy(f0 + (size_t)u0 * f0_stride, fd, fsz * 8);
    if (out_sp) memcpy(out_sp + (size_t)u0 * f0_stride * sp_row, sp_src, fsz * sp_row * 8);
    if (get_ap) memcpy(out_ap + (size_t)u0 * f0_stride * ap_row, ap_src, fsz * ap_row * 8);
#endif
  }
#ifndef WB_EMU
  cudaStreamSynchronize(s_in);
  cudaStreamSynchronize(s_compute);
  cudaStreamSynchronize(s_out);
  for (int i = 0; i < 2; ++i) { cudaEventDestroy(ev_in[i]); cudaEventDestroy(ev_done[i]); cudaEventDestroy(ev_out[i]); }
  cudaStreamDestroy(s_in);
  cudaStreamDestroy(s_out);
  if (!rc) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { ctx->last_error = cudaGetErrorString(e); rc = WORLD_B200_ECUDA; }
  }
#endif
  for (int i = 0; i < 2; ++i) {
    dev_free(din[i].p); dev_free(dx[i].p); dev_free(dt[i].p); dev_free(df[i].p);
    dev_free(dsp[i].p); dev_free(dap[i].p); dev_free(dcs[i].p); dev_free(dca[i].p);
  }
  if (!rc) rc = world_b200_synchronize(h);
  return rc;
}

}  // namespace

extern "C" int world_b200_analyze_host(WorldB200 *h, const double *x, int n_utts, int x_stride,
                                       const int *x_lengths, int fs, const WorldB200AnalysisOption *opt,
                                       double *time_axis, double *f0, int f0_stride, double *spectrogram,
                                       double *aperiodicity) {
  if (!h || !x || !opt || n_utts < 0 || fs <= 0 || x_stride <= 0 || f0_stride <= 0) return WORLD_B200_EINVAL;
  return analyze_pipeline(h, x, 0, n_utts, x_stride, x_lengths, fs, opt, 0, time_axis, f0, f0_stride, spectrogram,
                          aperiodicity);
}

extern "C" int world_b200_analyze_coded_host(WorldB200 *h, const void *x, int nbit, int n_utts, int x_stride,
                                             const int *x_lengths, int fs, const WorldB200AnalysisOption *opt,
                                             int number_of_dimensions, double *time_axis, double *f0, int f0_stride,
                                             double *coded_spectral_envelope, double *coded_aperiodicity) {
  if (!h || !x || !opt || n_utts < 0 || fs <= 0 || x_stride <= 0 || f0_stride <= 0) return WORLD_B200_EINVAL;
  if (nbit != 0 && nbit != 8 && nbit != 16 && nbit != 24 && nbit != 32) return WORLD_B200_EINVAL;
  if (number_of_dimensions < 1 || number_of_dimensions > opt->cheaptrick.fft_size / 4 + 1) {
    ctx_of(h)->last_error = "analyze_coded_host: number_of_dimensions must be in [1, fft_size/4 + 1]";
    return WORLD_B200_EINVAL;
  }
  return analyze_pipeline(h, x, nbit, n_utts, x_stride, x_lengths, fs, opt, number_of_dimensions, time_axis, f0,
                          f0_stride, coded_spectral_envelope, coded_aperiodicity);
}

// ------------------------------------------------------------------ legacy single-utterance API
namespace {
std::mutex g_legacy_mutex;
WorldB200 *g_legacy = nullptr;

WorldB200 *legacy_ctx() {
  if (!g_legacy) {
    int dev = 0;
    if (const char *e = getenv("WORLD_B200_DEVICE")) dev = atoi(e);
    if (world_b200_create(dev, &g_legacy) != 0) {
      fprintf(stderr, "world_b200: cannot create a CUDA context for the legacy API (no CPU path)\n");
      g_legacy = nullptr;
    }
  }
  return g_legacy;
}

void report(WorldB200 *h, const char *fn, int rc) {
  if (rc) fprintf(stderr, "world_b200: %s failed (%d): %s\n", fn, rc, world_b200_last_error(h));
}

// stage(x dev, t dev, f0 dev) helpers share the upload of x / time / f0
struct Legacy1 {
  WorldB200 *h; Ctx *ctx;
  double *x = nullptr, *t = nullptr, *f = nullptr;
  int rc = 0;
  Legacy1(const double *xh, int x_length, const double *th, const double *fh, int f0_length) {
    h = legacy_ctx();
    ctx = h ? ctx_of(h) : nullptr;
    if (!h) { rc = WORLD_B200_ECUDA; return; }
    x = (double *)dev_malloc(ctx, (size_t)x_length * 8);
    t = (double *)dev_malloc(ctx, (size_t)imax(1, f0_length) * 8);
    f = (double *)dev_malloc(ctx, (size_t)imax(1, f0_length) * 8);
    if (!x || !t || !f) { rc = WORLD_B200_ENOMEM; return; }
    rc = dev_memcpy_h2d(ctx, x, xh, (size_t)x_length * 8);
    if (!rc && th) rc = dev_memcpy_h2d(ctx, t, th, (size_t)f0_length * 8);
    if (!rc && fh) rc = dev_memcpy_h2d(ctx, f, fh, (size_t)f0_length * 8);
    if (!rc && !th) rc = dev_memset(ctx, t, 0, (size_t)imax(1, f0_length) * 8);
    if (!rc && !fh) rc = dev_memset(ctx, f, 0, (size_t)imax(1, f0_length) * 8);
  }
  ~Legacy1() { dev_free(x); dev_free(t); dev_free(f); }
};
}  // namespace

extern "C" {

void Dio(const double *x, int x_length, int fs, const DioOption *option, double *temporal_positions, double *f0) {
  std::lock_guard<std::mutex> lock(g_legacy_mutex);
  const int L = GetSamplesForDIO(fs, x_length, option->frame_period);
  Legacy1 d(x, x_length, nullptr, nullptr, L);
  int rc = d.rc;
  if (!rc) rc = world_b200_dio_batch(d.h, d.x, 1, x_length, nullptr, fs, option, d.t, d.f, L);
  if (!rc) rc = dev_memcpy_d2h(d.ctx, temporal_positions, d.t, (size_t)L * 8);
  if (!rc) rc = dev_memcpy_d2h(d.ctx, f0, d.f, (size_t)L * 8);
  if (!rc) rc = world_b200_synchronize(d.h);
  if (d.h) report(d.h, "Dio", rc);
}

void Harvest(const double *x, int x_length, int fs, const HarvestOption *option, double *temporal_positions,
             double *f0) {
  std::lock_guard<std::mutex> lock(g_legacy_mutex);
  const int L = GetSamplesForHarvest(fs, x_length, option->frame_period);
  Legacy1 d(x, x_length, nullptr, nullptr, L);
  int rc = d.rc;
  if (!rc) rc = world_b200_harvest_batch(d.h, d.x, 1, x_length, nullptr, fs, option, d.t, d.f, L);
  if (!rc) rc = dev_memcpy_d2h(d.ctx, temporal_positions, d.t, (size_t)L * 8);
  if (!rc) rc = dev_memcpy_d2h(d.ctx, f0, d.f, (size_t)L * 8);
  if (!rc) rc = world_b200_synchronize(d.h);
  if (d.h) report(d.h, "Harvest", rc);
}

void StoneMask(const double *x, int x_length, int fs, const double *temporal_positions, const double *f0,
               int f0_length, double *refined_f0) {
  std::lock_guard<std::mutex> lock(g_legacy_mutex);
  Legacy1 d(x, x_length, temporal_positions, f0, f0_length);
  int rc = d.rc;
  if (!rc) rc = world_b200_stonemask_batch(d.h, d.x, 1, x_length, nullptr, fs, d.t, d.f, nullptr, f0_length, d.f);
  if (!rc) rc = dev_memcpy_d2h(d.ctx, refined_f0, d.f, (size_t)f0_length * 8);
  if (!rc) rc = world_b200_synchronize(d.h);
  if (d.h) report(d.h, "StoneMask", rc);
}

static void rows_out(Legacy1 &d, const char *name, int rc, double *dev_rows, int f0_length, int bins, double **rows) {
  std::vector<double> flat((size_t)f0_length * bins);
  if (!rc) rc = dev_memcpy_d2h(d.ctx, flat.data(), dev_rows, flat.size() * 8);
  if (!rc) rc = world_b200_synchronize(d.h);
  if (!rc)
    for (int i = 0; i < f0_length; ++i) memcpy(rows[i], flat.data() + (size_t)i * bins, (size_t)bins * 8);
  if (d.h) report(d.h, name, rc);
}

void CheapTrick(const double *x, int x_length, int fs, const double *temporal_positions, const double *f0,
                int f0_length, const CheapTrickOption *option, double **spectrogram) {
  std::lock_guard<std::mutex> lock(g_legacy_mutex);
  Legacy1 d(x, x_length, temporal_positions, f0, f0_length);
  const int bins = option->fft_size / 2 + 1;
  double *rows = d.rc ? nullptr : (double *)dev_malloc(d.ctx, (size_t)imax(1, f0_length) * bins * 8);
  int rc = d.rc ? d.rc : (rows ? 0 : WORLD_B200_ENOMEM);
  if (!rc) rc = world_b200_cheaptrick_batch(d.h, d.x, 1, x_length, nullptr, fs, d.t, d.f, nullptr, f0_length, option, rows);
  rows_out(d, "CheapTrick", rc, rows, f0_length, bins, spectrogram);
  dev_free(rows);
}

void D4C(const double *x, int x_length, int fs, const double *temporal_positions, const double *f0, int f0_length,
         int fft_size, const D4COption *option, double **aperiodicity) {
  std::lock_guard<std::mutex> lock(g_legacy_mutex);
  Legacy1 d(x, x_length, temporal_positions, f0, f0_length);
  const int bins = fft_size / 2 + 1;
  double *rows = d.rc ? nullptr : (double *)dev_malloc(d.ctx, (size_t)imax(1, f0_length) * bins * 8);
  int rc = d.rc ? d.rc : (rows ? 0 : WORLD_B200_ENOMEM);
  if (!rc) rc = world_b200_d4c_batch(d.h, d.x, 1, x_length, nullptr, fs, d.t, d.f, nullptr, f0_length, fft_size, option, rows);
  rows_out(d, "D4C", rc, rows, f0_length, bins, aperiodicity);
  dev_free(rows);
}

void Synthesis(const double *f0, int f0_length, const double *const *spectrogram, const double *const *aperiodicity,
               int fft_size, double frame_period, int fs, int y_length, double *y) {
  std::lock_guard<std::mutex> lock(g_legacy_mutex);
  WorldB200 *h = legacy_ctx();
  if (!h) return;
  Ctx *ctx = ctx_of(h);
  const int bins = fft_size / 2 + 1;
  std::vector<double> flat((size_t)f0_length * bins);
  double *d_f0 = (double *)dev_malloc(ctx, (size_t)f0_length * 8);
  double *d_sp = (double *)dev_malloc(ctx, flat.size() * 8), *d_ap = (double *)dev_malloc(ctx, flat.size() * 8);
  double *d_y = (double *)dev_malloc(ctx, (size_t)y_length * 8);
  int rc = (d_f0 && d_sp && d_ap && d_y) ? 0 : WORLD_B200_ENOMEM;
  if (!rc) rc = dev_memcpy_h2d(ctx, d_f0, f0, (size_t)f0_length * 8);
  for (int pass = 0; pass < 2 && !rc; ++pass) {
    const double *const *rows = pass == 0 ? spectrogram : aperiodicity;
    for (int i = 0; i < f0_length; ++i) memcpy(flat.data() + (size_t)i * bins, rows[i], (size_t)bins * 8);
    rc = dev_memcpy_h2d(ctx, pass == 0 ? d_sp : d_ap, flat.data(), flat.size() * 8);
    if (!rc) rc = dev_sync(ctx);  // flat is reused
  }
  if (!rc) rc = world_b200_synthesis_batch(h, d_f0, nullptr, 1, f0_length, d_sp, d_ap, fft_size, frame_period, fs,
                                           nullptr, y_length, d_y);
  if (!rc) rc = dev_memcpy_d2h(ctx, y, d_y, (size_t)y_length * 8);
  if (!rc) rc = world_b200_synchronize(h);
  report(h, "Synthesis", rc);
  dev_free(d_f0); dev_free(d_sp); dev_free(d_ap); dev_free(d_y);
}

// ---- codec.h: rows in, rows out, one utterance
static void codec_rows(const char *name, const double *const *in_rows, int f0_length, int in_w, int out_w,
                       double **out_rows, int (*run)(WorldB200 *, const double *, double *, void *), void *arg) {
  std::lock_guard<std::mutex> lock(g_legacy_mutex);
  WorldB200 *h = legacy_ctx();
  if (!h || f0_length <= 0 || out_w <= 0) return;
  Ctx *ctx = ctx_of(h);
  std::vector<double> flat_in((size_t)f0_length * imax(1, in_w)), flat_out((size_t)f0_length * out_w);
  for (int i = 0; i < f0_length && in_w > 0; ++i) memcpy(flat_in.data() + (size_t)i * in_w, in_rows[i], (size_t)in_w * 8);
  double *d_in = (double *)dev_malloc(ctx, flat_in.size() * 8), *d_out = (double *)dev_malloc(ctx, flat_out.size() * 8);
  int rc = (d_in && d_out) ? 0 : WORLD_B200_ENOMEM;
  if (!rc) rc = dev_memcpy_h2d(ctx, d_in, flat_in.data(), flat_in.size() * 8);
  if (!rc) rc = run(h, d_in, d_out, arg);
  if (!rc) rc = dev_memcpy_d2h(ctx, flat_out.data(), d_out, flat_out.size() * 8);
  if (!rc) rc = world_b200_synchronize(h);
  if (!rc)
    for (int i = 0; i < f0_length; ++i) memcpy(out_rows[i], flat_out.data() + (size_t)i * out_w, (size_t)out_w * 8);
  report(h, name, rc);
  dev_free(d_in); dev_free(d_out);
}

struct CodecArgs { int f0_length, fs, fft_size, dims; };

void CodeAperiodicity(const double *const *aperiodicity, int f0_length, int fs, int fft_size,
                      double **coded_aperiodicity) {
  CodecArgs a = {f0_length, fs, fft_size, 0};
  codec_rows("CodeAperiodicity", aperiodicity, f0_length, fft_size / 2 + 1, GetNumberOfAperiodicities(fs),
             coded_aperiodicity, [](WorldB200 *h, const double *in, double *out, void *p) {
               const CodecArgs *a = (const CodecArgs *)p;
               return world_b200_code_aperiodicity_batch(h, in, 1, nullptr, a->f0_length, a->fs, a->fft_size, out);
             }, &a);
}

void DecodeAperiodicity(const double *const *coded_aperiodicity, int f0_length, int fs, int fft_size,
                        double **aperiodicity) {
  CodecArgs a = {f0_length, fs, fft_size, 0};
  codec_rows("DecodeAperiodicity", coded_aperiodicity, f0_length, GetNumberOfAperiodicities(fs), fft_size / 2 + 1,
             aperiodicity, [](WorldB200 *h, const double *in, double *out, void *p) {
               const CodecArgs *a = (const CodecArgs *)p;
               return world_b200_decode_aperiodicity_batch(h, in, 1, nullptr, a->f0_length, a->fs, a->fft_size, out);
             }, &a);
}

void CodeSpectralEnvelope(const double *const *spectrogram, int f0_length, int fs, int fft_size,
                          int number_of_dimensions, double **coded_spectral_envelope) {
  CodecArgs a = {f0_length, fs, fft_size, number_of_dimensions};
  codec_rows("CodeSpectralEnvelope", spectrogram, f0_length, fft_size / 2 + 1, number_of_dimensions,
             coded_spectral_envelope, [](WorldB200 *h, const double *in, double *out, void *p) {
               const CodecArgs *a = (const CodecArgs *)p;
               return world_b200_code_spectral_envelope_batch(h, in, 1, nullptr, a->f0_length, a->fs, a->fft_size,
                                                              a->dims, out);
             }, &a);
}

void DecodeSpectralEnvelope(const double *const *coded_spectral_envelope, int f0_length, int fs, int fft_size,
                            int number_of_dimensions, double **spectrogram) {
  CodecArgs a = {f0_length, fs, fft_size, number_of_dimensions};
  codec_rows("DecodeSpectralEnvelope", coded_spectral_envelope, f0_length, number_of_dimensions, fft_size / 2 + 1,
             spectrogram, [](WorldB200 *h, const double *in, double *out, void *p) {
               const CodecArgs *a = (const CodecArgs *)p;
               return world_b200_decode_spectral_envelope_batch(h, in, 1, nullptr, a->f0_length, a->fs, a->fft_size,
                                                                a->dims, out);
             }, &a);
}

}  // extern "C"
