// wb_f0common.cu -- kernels shared by DIO and Harvest (see wb_f0common.cuh for the design notes).
#include "wb_internal.h"
#include "wb_f0common.cuh"

namespace wb {

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

// interp1 (matlabfunctions.cpp:157-176) of one event train at time t.  Fine edge positions e_j live
// in global memory (complete list) and, for the most recent WB_RING of them, in a shared-memory ring
// (the searches below almost always stay inside the ring).  The (x, y) samples interp1 sees are
// (location, interval) of consecutive edges: x_j = (e_j + e_{j+1}) / 2 / afs, y_j = afs / (e_{j+1} - e_j).
#define WB_RING 512
struct Train { const double *g; const double *ring; int from; };  // ring holds indices >= from
WB_DEV double tr_get(const Train &t, int i) { return i >= t.from ? t.ring[i & (WB_RING - 1)] : t.g[i]; }
WB_DEV double train_location(const Train &t, int j, double afs) { return (tr_get(t, j) + tr_get(t, j + 1)) / 2.0 / afs; }

// x_j <= t, exactly as the reference evaluates it ((e_j + e_{j+1}) / 2.0 / afs <= t), but without
// the division in the common case: the quotient is within a few ulp of s / (2 afs), so unless
// s and 2 afs t agree to ~1e-13 relative the comparison of the products decides; only near ties
// is the reference expression evaluated.
WB_DEV bool location_le(const Train &tr, int j, double afs, double t, double two_afs_t) {
  const double s = tr_get(tr, j) + tr_get(tr, j + 1);
  const double margin = 1e-13 * (fabs(s) + fabs(two_afs_t));
  if (s < two_afs_t - margin) return true;
  if (s > two_afs_t + margin) return false;
  return s / 2.0 / afs <= t;
}

// first j in [lo, n_int) with x_j > t   (the caller guarantees #{j : x_j <= t} >= lo)
WB_DEV int train_count(const Train &tr, int lo, int n_int, double afs, double t) {
  int hi = n_int;
  const double two_afs_t = 2.0 * afs * t;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (location_le(tr, mid, afs, t, two_afs_t)) lo = mid + 1; else hi = mid;
  }
  return lo;
}

WB_DEV double train_interp(const Train &tr, int lo, int n_int, double afs, double t) {
  const int k = imin(n_int - 1, imax(1, train_count(tr, lo, n_int, afs, t)));
  const double e0 = tr_get(tr, k - 1), e1 = tr_get(tr, k), e2 = tr_get(tr, k + 1);
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

// One frame of one band: interp1 of the four trains, mean, (DIO) score, range checks
// (dio.cpp:441-465, 562-566 / harvest.cpp:240-254).
WB_DEV void sweep_candidate(const SweepParams &p, const Train *tr, const int *lo_j, const int *tot,
                            int i, double bf, double *cand, double *score) {
  const double t = i * p.frame_period / 1000.0;
  const double v0 = train_interp(tr[0], lo_j[0], tot[0] - 1, p.afs, t);
  const double v1 = train_interp(tr[1], lo_j[1], tot[1] - 1, p.afs, t);
  const double v2 = train_interp(tr[2], lo_j[2], tot[2] - 1, p.afs, t);
  const double v3 = train_interp(tr[3], lo_j[3], tot[3] - 1, p.afs, t);
  double c = (v0 + v1 + v2 + v3) / 4.0, sc = 0.0;
  if (p.mode == 0) {
    sc = sqrt(((v0 - c) * (v0 - c) + (v1 - c) * (v1 - c) + (v2 - c) * (v2 - c) + (v3 - c) * (v3 - c)) / 3.0);
    if (c > bf || c < bf / 2.0 || c > p.f0_ceil || c < p.f0_floor) { c = 0.0; sc = 100000.0; }
  } else {
    if (c > bf * 1.1 || c < bf * 0.9 || c > p.f0_ceil || c < p.f0_floor) c = 0.0;
  }
  cand[i] = c;
  if (score) score[i] = sc / (c + kTiny);  // dio.cpp:562-566
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
  unsigned long long *cnt = reinterpret_cast<unsigned long long *>(st + (T + 8 + ((T + 8) >> 3) + 8));  // G + 40
  double *ring = reinterpret_cast<double *>(cnt + (G + 40));                                           // 4 * WB_RING

  const int ylen = p.y_len[u];
  const double *sig = p.sig + (size_t)u * p.sig_stride + p.sig_origin;
  double *edges = p.edges + (size_t)u * p.edge_stride + (size_t)p.edge_off[b];
  const int cap = p.edge_cap[b];
  for (int j = tid; j < ntaps; j += nth) hrev[j] = __ldg(&p.taps_rev[p.tap_off[b] + j]);
  for (int j = ntaps + tid; j < ntaps + 8; j += nth) hrev[j] = 0.0;
  if (tid == 0) { st[0] = 0.0; st[1] = 0.0; }
  Train tr[4];
  for (int q = 0; q < 4; ++q) { tr[q].g = edges + (size_t)q * cap; tr[q].ring = ring + q * WB_RING; tr[q].from = 0; }
  int tot[4] = {0, 0, 0, 0};  // running event counts per train (identical in every thread)
  int lo_j[4] = {0, 0, 0, 0}; // per train: intervals below this index lie before every unfinished frame
  int next_frame = 0;         // frames [0, next_frame) are done
  const int nf = p.n_frames[u];
  double *cand = p.cand + ((size_t)u * p.n_bands + b) * p.frame_stride;
  double *score = p.score ? p.score + ((size_t)u * p.n_bands + b) * p.frame_stride : nullptr;
  const double bf = p.boundary[b];
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
      // pad8(8 (g+1) + 8 it + jj) = 9 (g+1) + 9 it + jj: one running pointer, immediate offsets
      const double *sp = seg + 9 * (g + 1);
      const double *hp = hrev;
#pragma unroll
      for (int r = 0; r < R; ++r) { acc[r] = 0.0; win[r] = seg[9 * g + r]; }
      for (int j0 = 0; j0 < ntaps; j0 += R, sp += 9, hp += R) {
#pragma unroll
        for (int jj = 0; jj < R; ++jj) {
          const double hj = hp[jj];  // zero beyond ntaps
#pragma unroll
          for (int r = 0; r < R; ++r) acc[r] = fma(hj, win[(r + jj) & (R - 1)], acc[r]);
          win[jj] = sp[jj];
        }
      }
#pragma unroll
      for (int r = 0; r < R; ++r) st[pad8(2 + base + r)] = acc[r];
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
        const double a = st[pad8(base + r)], bb = st[pad8(base + r + 1)], cc = st[pad8(base + r + 2)];
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
    int keep[4];  // after this tile the ring holds event indices >= keep[q]
    keep[0] = imax(0, tot[0] + (int)(tile_total & 0xffffull) - WB_RING);
    keep[1] = imax(0, tot[1] + (int)((tile_total >> 16) & 0xffffull) - WB_RING);
    keep[2] = imax(0, tot[2] + (int)((tile_total >> 32) & 0xffffull) - WB_RING);
    keep[3] = imax(0, tot[3] + (int)((tile_total >> 48) & 0xffffull) - WB_RING);
    for (int g = tid; g < G; g += nth) {
      const unsigned long long o = cnt[g];
      int o0 = tot[0] + (int)(o & 0xffffull), o1 = tot[1] + (int)((o >> 16) & 0xffffull);
      int o2 = tot[2] + (int)((o >> 32) & 0xffffull), o3 = tot[3] + (int)((o >> 48) & 0xffffull);
      const int base = R * g;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int i = n0 - 2 + base + r;
        const double a = st[pad8(base + r)], bb = st[pad8(base + r + 1)], cc = st[pad8(base + r + 2)];
        const double d0 = bb - a, d1 = cc - bb;
        const double e = (double)(i + 1);
        if (i >= 0 && i + 1 <= ylen - 1) {
          if (0.0 < a && bb <= 0.0) {
            const double v = e - a / (bb - a);
            if (o0 < cap) edges[o0] = v;
            if (o0 >= keep[0]) ring[o0 & (WB_RING - 1)] = v;
            ++o0;
          }
          if (a < 0.0 && 0.0 <= bb) {
            const double v = e - a / (bb - a);
            if (o1 < cap) edges[(size_t)cap + o1] = v;
            if (o1 >= keep[1]) ring[WB_RING + (o1 & (WB_RING - 1))] = v;
            ++o1;
          }
        }
        if (i >= 0 && i + 1 <= ylen - 2) {
          if (0.0 < d0 && d1 <= 0.0) {
            const double v = e - d0 / (d1 - d0);
            if (o2 < cap) edges[2 * (size_t)cap + o2] = v;
            if (o2 >= keep[2]) ring[2 * WB_RING + (o2 & (WB_RING - 1))] = v;
            ++o2;
          }
          if (d0 < 0.0 && 0.0 <= d1) {
            const double v = e - d0 / (d1 - d0);
            if (o3 < cap) edges[3 * (size_t)cap + o3] = v;
            if (o3 >= keep[3]) ring[3 * WB_RING + (o3 & (WB_RING - 1))] = v;
            ++o3;
          }
        }
      }
    }
    tot[0] += (int)(tile_total & 0xffffull); tot[1] += (int)((tile_total >> 16) & 0xffffull);
    tot[2] += (int)((tile_total >> 32) & 0xffffull); tot[3] += (int)((tile_total >> 48) & 0xffffull);
    for (int q = 0; q < 4; ++q) tr[q].from = keep[q];
#ifndef WB_EMU
    __threadfence_block();
#endif
    WB_SYNC();
    if (tid == 0) { st[pad8(0)] = st[pad8(T)]; st[pad8(1)] = st[pad8(T + 1)]; }
    // ---- streaming candidates: every frame whose time lies before the last complete interval of
    // all four trains can be interpolated now; its events are the most recent ones (cache hot).
    bool can = true;
    double t_safe = 0.0;
    for (int q = 0; q < 4; ++q) {
      if (tot[q] > cap || tot[q] < 2) { can = false; break; }
      const double loc = train_location(tr[q], tot[q] - 2, p.afs);
      t_safe = (q == 0 || loc < t_safe) ? loc : t_safe;
    }
    if (can) {
      // first frame index with t_i >= t_safe (t_i = i * frame_period / 1000.0, monotone in i)
      int i_safe = (int)(t_safe * 1000.0 / p.frame_period);
      if (i_safe < 0) i_safe = 0;
      while (i_safe > 0 && !((i_safe - 1) * p.frame_period / 1000.0 < t_safe)) --i_safe;
      while (i_safe < nf && (i_safe * p.frame_period / 1000.0 < t_safe)) ++i_safe;
      if (i_safe > nf) i_safe = nf;
      if (i_safe > next_frame) {
        for (int i = next_frame + tid; i < i_safe; i += nth)
          sweep_candidate(p, tr, lo_j, tot, i, bf, cand, score);
        // all later frames have t >= t_safe: their interval counts are at least those of t_safe
        for (int q = 0; q < 4; ++q) {
          const int c = train_count(tr[q], lo_j[q], tot[q] - 1, p.afs, (i_safe - 1) * p.frame_period / 1000.0);
          lo_j[q] = imax(lo_j[q], c);
        }
        next_frame = i_safe;
      }
    }
    WB_SYNC();
  }
  // ---- frames after the last complete interval (interp1 extrapolates from the last two samples)
  bool ok = true;
  for (int q = 0; q < 4; ++q) {
    if (tot[q] > cap) { if (tid == 0) atomicOr_status(p.status, 4); ok = false; }
    const int ni = tot[q] < 2 ? 0 : tot[q] - 1;  // ZeroCrossingEngine returns count-1 (0 if count<2)
    if (ni - 2 <= 0) ok = false;                 // CheckEvent(n - 2), dio.cpp:475-484
  }
  if (ok) {
    for (int i = next_frame + tid; i < nf; i += nth) sweep_candidate(p, tr, lo_j, tot, i, bf, cand, score);
  } else {
    for (int i = tid; i < nf; i += nth) {
      cand[i] = 0.0;
      if (score) score[i] = 100000.0 / (0.0 + kTiny);
    }
  }
}

// extended input of decimate(): 9 mirrored samples on both sides of the edge-padded signal
WB_DEV double dec_ext(const double *__restrict__ x, int n, int lag, int nx, int i) {
#define WB_XIN(k) x[imin(n - 1, imax(0, (k) - lag))]
  if (i < 9) return 2 * WB_XIN(0) - WB_XIN(9 - i);
  if (i < 9 + nx) return WB_XIN(i - 9);
  return 2 * WB_XIN(nx - 1) - WB_XIN(nx - 2 - (i - (9 + nx)));
#undef WB_XIN
}

// pass 0: tmp[i] = forward IIR of ext;  pass 1: backward IIR over tmp, decimated into y
template <int kPass>
WB_KERNEL_PLAIN decimate_pass_kernel(DecimateParams p) {
  const int u = blockIdx.y;
  const int n = p.x_len[u];
  const int nx = n + 2 * p.lag, nt = nx + 18;
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long begin = g * WB_DEC_BLOCK;
  if (begin >= nt) return;
  const int end = (int)(begin + WB_DEC_BLOCK < nt ? begin + WB_DEC_BLOCK : nt);
  const int start = (int)(begin - WB_DEC_WARM > 0 ? begin - WB_DEC_WARM : 0);
  const double *x = p.x + (size_t)u * p.x_stride;
  double *tmp = p.tmp + (size_t)u * p.tmp_stride;
  double a[3], b[2];
  decimate_coefficients(p.ratio, a, b);
  double w0 = 0.0, w1 = 0.0, w2 = 0.0;
  if (kPass == 0) {
    for (int i = start; i < end; ++i) {
      const double wt = dec_ext(x, n, p.lag, nx, i) + a[0] * w0 + a[1] * w1 + a[2] * w2;
      if (i >= begin) tmp[i] = b[0] * wt + b[1] * w0 + b[1] * w1 + b[0] * w2;
      w2 = w1; w1 = w0; w0 = wt;
    }
  } else {
    // second filter runs over the reversed forward output; its result, reversed again, is tmp1 of
    // decimate(): final[j] with j = nt - 1 - i.  y[k] = final[nbeg + k r + 8]  (matlabfunctions.cpp:196-200)
    const int nout = (nx - 1) / p.ratio + 1;
    const int nbeg = p.ratio - p.ratio * nout + nx;
    const int n_out = p.n_out_mode == 0 ? 1 + n / p.ratio : static_cast<int>(ceil(static_cast<double>(n) / p.ratio));
    double *y = p.y + (size_t)u * p.y_stride + p.y_origin;
    for (int i = start; i < end; ++i) {
      const double wt = tmp[nt - 1 - i] + a[0] * w0 + a[1] * w1 + a[2] * w2;
      if (i >= begin) {
        const int j = nt - 1 - i - 8;          // = nbeg + k r  for a kept sample
        const int d = j - nbeg;
        if (d >= 0 && j < nx + 9 && d % p.ratio == 0) {
          const int k = d / p.ratio - p.first;
          if (k >= 0 && k < n_out) y[k] = b[0] * wt + b[1] * w0 + b[1] * w1 + b[0] * w2;
        }
      }
      w2 = w1; w1 = w0; w0 = wt;
    }
  }
}

void launch_decimate(Ctx *ctx, const DecimateParams &p, int max_x_len, unsigned n_utts) {
  const long long nt = (long long)max_x_len + 2 * p.lag + 18;
  const long long threads = (nt + WB_DEC_BLOCK - 1) / WB_DEC_BLOCK;
  const unsigned blocks = (unsigned)((threads + 63) / 64);
  WB_LAUNCH_FLAT(decimate_pass_kernel<0>, dim3(blocks, n_utts), 64, 0, ctx->stream, p);
  WB_LAUNCH_FLAT(decimate_pass_kernel<1>, dim3(blocks, n_utts), 64, 0, ctx->stream, p);
}

void launch_fir_plain(Ctx *ctx, const FirParams &p, unsigned tiles, unsigned n_utts) {
  const size_t smem = fir_plain_smem_bytes(p.ntaps);
#ifndef WB_EMU
  cudaFuncSetAttribute(fir_plain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
#endif
  WB_LAUNCH_COOP(fir_plain_kernel, dim3(tiles, n_utts), 256, smem, ctx->stream, p);
}

void launch_band_sweep(Ctx *ctx, const SweepParams &p, unsigned n_utts) {
  const size_t smem = sweep_smem_bytes(p.max_taps);
#ifndef WB_EMU
  cudaFuncSetAttribute(band_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
#endif
  WB_LAUNCH_COOP(band_sweep_kernel, dim3((unsigned)p.n_bands, n_utts), WB_SWEEP_THREADS, smem, ctx->stream, p);
}

}  // namespace wb
