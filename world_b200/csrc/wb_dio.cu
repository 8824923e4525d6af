// wb_dio.cu -- DIO F0 estimation for a batch (replaces Dio()/DioGeneralBody, dio.cpp:578-648).
// Algorithm card: SURVEY.md A3.
//   K-DIOp  dio_prep_kernel      decimate (optional) + DC removal                (dio.cpp:60-79)
//   fir_plain_kernel             zero-phase 50 Hz low-cut FIR                     (dio.cpp:40-53, 86-104)
//   K-DIOf  band_sweep_kernel    per band Nuttall low-pass + 4 zero-crossing trains + interp1 onto
//                                the frame grid -> candidates and scores           (dio.cpp:296-568)
//   K-DIOc  dio_contour_kernel   best candidate + FixStep1..4                     (dio.cpp:112-289)
// Host side computes every size and filter tap with the reference's own double-precision
// expressions (and the host libm), so frame counts, band lists and taps are bit-identical.
#include "wb_internal.h"
#include "wb_f0common.cuh"
#include <stdlib.h>
#include <vector>

namespace wb {

// Event-list capacity per band and train: crossings of a signal band-limited around/below
// `boundary` cannot be denser than ~boundary per second for long; 2.5x margin, hard bound
// ylen/2+2 (a negative-going crossing needs two samples).  The lists are history rings: more events
// than this wrap around; only a look-back beyond the last `cap` events raises status bit 4.
static void plan_edge_caps(const std::vector<double> &boundary, double afs, int max_ylen,
                           std::vector<int> *cap, std::vector<long long> *off, size_t *stride) {
  const int nb = (int)boundary.size();
  cap->resize(nb); off->resize(nb);
  long long run = 0;
  long long floor_cap = 2048;
  if (const char *e = getenv("WB_EDGE_CAP_MIN")) floor_cap = atoll(e) > 0 ? atoll(e) : floor_cap;   // test hook: force wraps
  for (int i = 0; i < nb; ++i) {
    const long long hard = (long long)max_ylen / 2 + 2;
    long long soft = (long long)(2.5 * boundary[i] * max_ylen / afs) + 64;
    if (soft < floor_cap) soft = floor_cap;   // a tile can append up to 1025 events per train; the rings look back 256
    // DIO keeps every event: with the mirroring ripple in, digital silence makes the difference trains fire
    // every sample while the crossing trains may stay silent, so frames cannot be finalised until the silence
    // ends and the look-back is as long as the silence (few bands, so the full lists are affordable)
    soft = hard;
    (*cap)[i] = (int)(soft < hard ? soft : hard);
    (*off)[i] = run;
    run += 4LL * (*cap)[i];
  }
  *stride = (size_t)run;
}

struct DioPrepParams {
  const double *x; const int *x_len; int x_stride;
  int ratio;
  double *y; size_t y_stride; int y_origin;   // mean-removed signal, zero padded
  int *y_len;                                  // out: 1 + x_len / ratio
};

WB_KERNEL(256, 2) dio_prep_kernel(DioPrepParams p) {
  WB_SHARED double red[WB_RED_DOUBLES];
  const int tid = WB_TID, nth = WB_NTH, u = blockIdx.x;
  const int n = p.x_len[u];
  const double *x = p.x + (size_t)u * p.x_stride;
  double *y = p.y + (size_t)u * p.y_stride + p.y_origin;
  const int ylen = 1 + n / p.ratio;
  if (p.ratio != 1) {
    // launch_decimate() already left (n-1)/r+1 (+ a few mirrored-edge) samples in y; the rest of y
    // stays zero (dio.cpp:67-73)
  } else {
    for (int i = tid; i < n; i += nth) y[i] = x[i];
    WB_SYNC();
  }
  double s = 0.0;
  for (int i = tid; i < ylen; i += nth) s += y[i];
  const double mean = block_sum(s, red) / ylen;
  for (int i = tid; i < ylen; i += nth) y[i] = y[i] - mean;
  if (tid == 0) p.y_len[u] = ylen;
}

struct DioContourParams {
  const double *cand; const double *score; int n_bands; int frame_stride;
  const int *f_len; double frame_period, f0_floor, allowed_range;
  double *work;      // [n][4][frame_stride]
  int *sections;     // [n][2][frame_stride]
  double *time_axis; double *f0; int out_stride;
};

WB_DEV double dio_select_best(double cur, double past, const double *cand, int nb, int stride, int target,
                              double allowed) {
  const double reference = (cur * 3.0 - past) / 2.0;
  double best = cand[target];
  double min_err = fabs(reference - best);
  for (int b = 1; b < nb; ++b) {
    const double c = cand[(size_t)b * stride + target];
    const double e = fabs(reference - c);
    if (e < min_err) { min_err = e; best = c; }
  }
  if (fabs(1.0 - best / reference) > allowed) return 0.0;
  return best;
}

WB_KERNEL(256, 2) dio_contour_kernel(DioContourParams p) {
  const int tid = WB_TID, nth = WB_NTH, u = blockIdx.x;
  const int L = p.f_len[u], nb = p.n_bands, fstr = p.frame_stride;
  const double *cand = p.cand + (size_t)u * nb * fstr;
  const double *score = p.score + (size_t)u * nb * fstr;
  double *best = p.work + (size_t)u * 4 * fstr, *s1 = best + fstr, *s2 = s1 + fstr, *s3 = s2 + fstr;
  int *pos = p.sections + (size_t)u * 2 * fstr, *neg = pos + fstr;
  double *f0 = p.f0 + (size_t)u * p.out_stride, *ta = p.time_axis + (size_t)u * p.out_stride;
  for (int i = tid; i < L; i += nth) {
    ta[i] = i * p.frame_period / 1000.0;  // dio.cpp:609-610
    double bs = score[i], bf = cand[i];
    for (int b = 1; b < nb; ++b) {
      const double sc = score[(size_t)b * fstr + i];
      if (bs > sc) { bs = sc; bf = cand[(size_t)b * fstr + i]; }
    }
    best[i] = bf;
    f0[i] = 0.0;
  }
  const int vrm = static_cast<int>(0.5 + 1000.0 / p.frame_period / p.f0_floor) * 2 + 1;
  if (L <= vrm) return;  // dio.cpp:266 (the reference leaves f0 unwritten; zeros here)
  WB_SYNC();
  // step 1: s1 (f0_base is `best` with the first/last vrm frames zeroed)
  for (int i = tid; i < L; i += nth) {
    double v = 0.0;
    if (i >= vrm) {
      const double bi = (i < L - vrm) ? best[i] : 0.0;
      const double bp = (i - 1 >= vrm && i - 1 < L - vrm) ? best[i - 1] : 0.0;
      v = fabs((bi - bp) / (kTiny + bi)) < p.allowed_range ? bi : 0.0;
    }
    s1[i] = v;
  }
  WB_SYNC();
  const int center = (vrm - 1) / 2;
  for (int i = tid; i < L; i += nth) {
    double v = s1[i];
    if (i >= center && i < L - center)
      for (int j = -center; j <= center; ++j)
        if (s1[i + j] == 0) { v = 0.0; break; }
    s2[i] = v;
    s3[i] = v;
  }
  WB_SYNC();
  if (tid == 0) {
    int pc = 0, nc = 0;
    for (int i = 1; i < L; ++i) {
      if (s2[i] == 0 && s2[i - 1] != 0) neg[nc++] = i - 1;
      else if (s2[i - 1] == 0 && s2[i] != 0) pos[pc++] = i;
    }
    // step 3: forward extension (dio.cpp:215-231), in s3
    for (int i = 0; i < nc; ++i) {
      const int limit = i == nc - 1 ? L - 1 : neg[i + 1];
      for (int j = neg[i]; j < limit; ++j) {
        s3[j + 1] = dio_select_best(s3[j], s3[j - 1], cand, nb, fstr, j + 1, p.allowed_range);
        if (s3[j + 1] == 0) break;
      }
    }
    // step 4: backward extension (dio.cpp:237-253), in place on s3 (copy semantics are identical:
    // the reference copies step3 into step4 first and then only reads step4)
    for (int i = pc - 1; i >= 0; --i) {
      const int limit = i == 0 ? 1 : pos[i - 1];
      for (int j = pos[i]; j > limit; --j) {
        s3[j - 1] = dio_select_best(s3[j], s3[j + 1], cand, nb, fstr, j - 1, p.allowed_range);
        if (s3[j - 1] == 0) break;
      }
    }
  }
  WB_SYNC();
  for (int i = tid; i < L; i += nth) f0[i] = s3[i];
}

int dio_run(Ctx *ctx, const Batch &b, const DioParams &opt, double *time_axis_out, double *f0_out) {
  if (b.n <= 0) return 0;
  const int fs = b.fs;
  // band list and sizes, dio.cpp:582-594
  const int nb = 1 + static_cast<int>(log(opt.f0_ceil / opt.f0_floor) / kLog2 * opt.channels_in_octave);
  if (nb < 1 || nb > 4096) { ctx->last_error = "Dio: bad band count"; return 3; }
  std::vector<double> boundary(nb);
  for (int i = 0; i < nb; ++i) boundary[i] = opt.f0_floor * pow(2.0, (i + 1) / opt.channels_in_octave);
  const int ratio = imax(imin(opt.speed, 12), 1);
  const double afs = static_cast<double>(fs) / ratio;
  // low-cut filter (DesignLowCutFilter, dio.cpp:40-53) as a centred FIR of 2c+1 taps
  const int c = round_half_away(afs / 50.0);
  const int nlc = 2 * c + 1;
  std::vector<double> lc(nlc);
  {
    for (int i = 1; i <= nlc; ++i) lc[i - 1] = 0.5 - 0.5 * cos(i * 2.0 * kPi / (nlc + 1));
    double sum = 0.0;
    for (int i = 0; i < nlc; ++i) sum += lc[i];
    for (int i = 0; i < nlc; ++i) lc[i] = -lc[i] / sum;
    lc[c] += 1.0;  // "low_cut_filter[0] += 1.0" after the circular shift that centres the filter
  }
  std::vector<double> lc_rev(lc.rbegin(), lc.rend());
  // per band Nuttall low-pass of 4*ha taps (GetFilteredSignal, dio.cpp:296-337)
  std::vector<int> tap_off(nb), ntaps(nb), shift(nb);
  std::vector<double> taps;
  int max_taps = 0;
  for (int i = 0; i < nb; ++i) {
    const int ha = round_half_away(afs / boundary[i] / 2.0);
    if (ha < 1) {
      // band above afs / 2 (heavy decimation): the reference's window has zero length, its filtered signal
      // is identically zero, so the band yields no events -- the same through an all-zero filter
      tap_off[i] = (int)taps.size(); ntaps[i] = 8; shift[i] = 0;
      for (int j = 0; j < 16; ++j) taps.push_back(0.0);
      if (8 > max_taps) max_taps = 8;
      continue;
    }
    const int len = ha * 4;
    tap_off[i] = (int)taps.size(); ntaps[i] = len; shift[i] = ha * 2;
    std::vector<double> w(len);
    for (int j = 0; j < len; ++j) {
      const double tmp = j / (len - 1.0);
      w[j] = 0.355768 - 0.487396 * cos(2.0 * kPi * tmp) + 0.144232 * cos(4.0 * kPi * tmp) -
             0.012604 * cos(6.0 * kPi * tmp);
    }
    for (int j = len - 1; j >= 0; --j) taps.push_back(w[j]);
    for (int j = 0; j < 8; ++j) taps.push_back(0.0);
    if (len > max_taps) max_taps = len;
  }
  const size_t smem_sweep = sweep_smem_bytes(max_taps), smem_fir = fir_plain_smem_bytes(nlc);
  if (smem_sweep > 200 * 1024 || smem_fir > 200 * 1024) {
    ctx->last_error = "Dio: filters too long for shared memory (lower fs / raise f0_floor / use speed > 1)";
    return 3;
  }
  const int max_ylen = 1 + b.max_x_len / ratio;
  const int T = WB_SWEEP_T;
  const int padl = imax(max_taps, nlc) + 16;
  const size_t y_stride = (size_t)padl + max_ylen + 2 * c + 3 * T + max_taps + nlc + 64;
  std::vector<int> ecap; std::vector<long long> eoff; size_t edge_stride = 0;
  plan_edge_caps(boundary, afs, max_ylen, &ecap, &eoff, &edge_stride);
  const int fstr = b.f_stride;
  const size_t tmp_stride = ratio != 1 ? (size_t)b.max_x_len + 32 : 0;
  const size_t per_utt = y_stride * 16 + edge_stride * 8 + (size_t)nb * fstr * 16 +
                         (size_t)fstr * (4 * 8 + 2 * 4) + tmp_stride * 16 + 256;
  int chunk = balanced_chunk(imin(b.n, 65535), (int)dmin(65535.0, (double)ctx->scratch_budget / (double)per_utt));
#ifndef WB_EMU
#endif
  for (int u0 = 0; u0 < b.n; u0 += chunk) {
    const int n = imin(chunk, b.n - u0);
    ArenaPlan plan;
    const size_t o_y = plan.add((size_t)n * y_stride * 8), o_ylc = plan.add((size_t)n * y_stride * 8);
    const size_t o_ylen = plan.add((size_t)n * 4), o_nyq = plan.add((size_t)n * 32), o_nfft = plan.add((size_t)n * 4);
    const size_t o_edges = plan.add((size_t)n * edge_stride * 8);
    const size_t o_ecap = plan.add(nb * 4), o_eoff = plan.add(nb * 8);
    const size_t o_cand = plan.add((size_t)n * nb * fstr * 8), o_score = plan.add((size_t)n * nb * fstr * 8);
    const size_t o_work = plan.add((size_t)n * 4 * fstr * 8), o_sec = plan.add((size_t)n * 2 * fstr * 4);
    const size_t o_tmp = plan.add((size_t)n * tmp_stride * 8);
    const size_t o_lc = plan.add(lc_rev.size() * 8 + 64), o_taps = plan.add(taps.size() * 8);
    const size_t o_toff = plan.add(nb * 4), o_nt = plan.add(nb * 4), o_sh = plan.add(nb * 4), o_bd = plan.add(nb * 8);
    unsigned char *blk = arena_block(ctx, plan.total);
    if (!blk) return 2;
    double *y = (double *)(blk + o_y), *ylc = (double *)(blk + o_ylc);
    int *ylen = (int *)(blk + o_ylen);
    int rc = dev_memset(ctx, y, 0, (size_t)n * y_stride * 8);
    if (!rc) rc = dev_memset(ctx, ylc, 0, (size_t)n * y_stride * 8);
    if (!rc) rc = dev_memcpy_h2d(ctx, blk + o_lc, lc_rev.data(), lc_rev.size() * 8);
    if (!rc) rc = dev_memcpy_h2d(ctx, blk + o_taps, taps.data(), taps.size() * 8);
    if (!rc) rc = dev_memcpy_h2d(ctx, blk + o_toff, tap_off.data(), nb * 4);
    if (!rc) rc = dev_memcpy_h2d(ctx, blk + o_nt, ntaps.data(), nb * 4);
    if (!rc) rc = dev_memcpy_h2d(ctx, blk + o_sh, shift.data(), nb * 4);
    if (!rc) rc = dev_memcpy_h2d(ctx, blk + o_bd, boundary.data(), nb * 8);
    if (!rc) rc = dev_memcpy_h2d(ctx, blk + o_ecap, ecap.data(), nb * 4);
    if (!rc) rc = dev_memcpy_h2d(ctx, blk + o_eoff, eoff.data(), nb * 8);
    if (rc) return rc;

    DioPrepParams pp;
    pp.x = b.x + (size_t)u0 * b.x_stride; pp.x_len = b.x_len + u0; pp.x_stride = b.x_stride; pp.ratio = ratio;
    pp.y = y; pp.y_stride = y_stride; pp.y_origin = padl; pp.y_len = ylen;
    if (ratio != 1) {
      DecimateParams dp;
      dp.x = pp.x; dp.x_len = pp.x_len; dp.x_stride = pp.x_stride; dp.ratio = ratio; dp.lag = 0;
      dp.tmp = (double *)(blk + o_tmp); dp.tmp_stride = tmp_stride;
      dp.y = y; dp.y_stride = y_stride; dp.y_origin = padl; dp.first = 0; dp.n_out_mode = 0;
      launch_decimate(ctx, dp, b.max_x_len, (unsigned)n);
    }
    WB_LAUNCH_COOP(dio_prep_kernel, dim3((unsigned)n), 256, 0, ctx->stream, pp);

    // ylc(q) = sum_k lc[k] y(q - k), q in [0, ylen + 2c): time index n = q - c
    FirParams fp;
    fp.in = y; fp.in_stride = y_stride; fp.in_origin = padl;
    fp.out = ylc; fp.out_stride = y_stride; fp.out_origin = padl;
    fp.base_len = ylen; fp.extra_len = 2 * c; fp.taps_rev = (const double *)(blk + o_lc); fp.ntaps = nlc;
    const unsigned tiles = (unsigned)((max_ylen + 2 * c + 2047) / 2048);
    launch_fir_plain(ctx, fp, tiles, (unsigned)n);

    // the reference's FFT size per utterance: GetSuitableFFTSize(y_length + 2 c + 1 + 4 int(1 + afs / b0 / 2))
    // (dio.cpp:590-592, common.cpp:51-54), host libm like there
    std::vector<int> nfft(n);
    for (int i = 0; i < n; ++i) {
      const int xl = b.x_len_host ? b.x_len_host[u0 + i] : b.x_stride;
      const int yl = 1 + xl / ratio;
      const int sample = yl + c * 2 + 1 + 4 * static_cast<int>(1.0 + afs / boundary[0] / 2.0);
      nfft[i] = static_cast<int>(pow(2.0, static_cast<int>(log(static_cast<double>(sample)) / kLog2) + 1.0));
    }
    rc = dev_memcpy_h2d(ctx, blk + o_nfft, nfft.data(), (size_t)n * 4);
    if (rc) return rc;
    NyquistParams np_;
    np_.sig = ylc; np_.stride = y_stride; np_.origin = padl; np_.y_len = ylen; np_.c = c;
    np_.nfft = (const int *)(blk + o_nfft); np_.nyq = (double *)(blk + o_nyq);
    launch_nyquist_bins(ctx, np_, (unsigned)n);

    SweepParams sp;
    sp.sig = ylc; sp.sig_stride = y_stride; sp.sig_origin = padl + c; sp.y_len = ylen; sp.n_bands = nb;
    sp.taps_rev = (const double *)(blk + o_taps); sp.tap_off = (const int *)(blk + o_toff);
    sp.ntaps = (const int *)(blk + o_nt); sp.shift = (const int *)(blk + o_sh);
    sp.boundary = (const double *)(blk + o_bd); sp.afs = afs;
    sp.edges = (double *)(blk + o_edges); sp.edge_stride = edge_stride;
    sp.edge_cap = (const int *)(blk + o_ecap); sp.edge_off = (const long long *)(blk + o_eoff);
    sp.n_frames = b.f_len + u0; sp.frame_stride = fstr; sp.frame_period = opt.frame_period;
    sp.mode = 0; sp.f0_floor = opt.f0_floor; sp.f0_ceil = opt.f0_ceil;
    sp.nyq = (const double *)(blk + o_nyq); sp.ripple = getenv("WB_NO_RIPPLE") ? 0 : 1;   // A/B switch for experiments
    sp.cand = (double *)(blk + o_cand); sp.score = (double *)(blk + o_score);
    sp.max_taps = max_taps; sp.status = ctx->status_dev;
    launch_band_sweep(ctx, sp, (unsigned)n);

    DioContourParams cp;
    cp.cand = sp.cand; cp.score = sp.score; cp.n_bands = nb; cp.frame_stride = fstr; cp.f_len = b.f_len + u0;
    cp.frame_period = opt.frame_period; cp.f0_floor = opt.f0_floor; cp.allowed_range = opt.allowed_range;
    cp.work = (double *)(blk + o_work); cp.sections = (int *)(blk + o_sec);
    cp.time_axis = time_axis_out + (size_t)u0 * b.f_stride; cp.f0 = f0_out + (size_t)u0 * b.f_stride;
    cp.out_stride = b.f_stride;
    WB_LAUNCH_COOP(dio_contour_kernel, dim3((unsigned)n), 256, 0, ctx->stream, cp);
    rc = dev_check(ctx, "dio");
    if (rc) return rc;
  }
  return 0;
}

}  // namespace wb
