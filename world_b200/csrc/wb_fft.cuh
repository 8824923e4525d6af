// wb_fft.cuh -- shared-memory FFTs for the frame kernels (FP64, power-of-two, forward only).
//
// Replaces the reference's L1 layer (src/fft.cpp, an Ooura split-radix FFT behind an
// FFTW-shaped plan API, fft.h:37-44) for every per-frame transform.  It is NOT a port of that
// code: transforms live entirely in shared memory, are executed cooperatively by one CTA, and
// only the conventions are kept:
//   r2c : X[k] = sum_n x[n] exp(-j 2 pi k n / N), k = 0..N/2          (fft.cpp:49-60)
//   c2r of a real, even spectrum == Re r2c(mirror(spectrum))           (fft.cpp:26-35)
// so every transform the analysis path needs is a forward real FFT.
//
// Twiddles come from one table tw[k] = exp(-j 2 pi k / WB_TW_N), k < WB_TW_N/2, computed on
// the host in long double and kept in global memory (L1/L2 resident, read with __ldg).
#pragma once
#include "wb_platform.cuh"

namespace wb {

#define WB_TW_LOG2 13
#define WB_TW_N (1 << WB_TW_LOG2)  // supports complex FFTs up to 8192, real up to 8192

WB_DEV unsigned bit_reverse(unsigned v, int bits) {
#ifdef WB_EMU
  unsigned r = 0;
  for (int i = 0; i < bits; ++i) { r = (r << 1) | (v & 1u); v >>= 1; }
  return r;
#else
  return __brev(v) >> (32 - bits);
#endif
}

// In-place forward complex FFT of z[0..2^lg) (natural order in, natural order out).
// Bit reversal, then decimation-in-time stages taken two at a time as radix-4 butterflies held in
// registers (half the shared-memory passes and barriers of radix-2; one table twiddle per
// butterfly, the second one is its square), plus one radix-2 stage when lg is odd.
// Ends with a barrier.
WB_DEV void cfft_forward(double2 *z, int lg, const double2 *__restrict__ tw) {
  const int tid = WB_TID, nth = WB_NTH;
  const int n = 1 << lg;
  WB_UNROLL4
  for (int i = tid; i < n; i += nth) {
    const int j = (int)bit_reverse((unsigned)i, lg);
    if (i < j) { const double2 a = z[i]; z[i] = z[j]; z[j] = a; }
  }
  WB_SYNC();
  int s = 1;
  if (lg >= 2) {
    // stages 1+2: twiddles are 1 and -j
    WB_UNROLL4
    for (int q = tid; q < (n >> 2); q += nth) {
      double2 *p = z + 4 * q;
      const double2 a = p[0], b = p[1], c = p[2], d = p[3];
      const double2 ab0 = make_double2(a.x + b.x, a.y + b.y);
      const double2 ab1 = make_double2(a.x - b.x, a.y - b.y);
      const double2 cd0 = make_double2(c.x + d.x, c.y + d.y);
      const double2 cd1 = make_double2(c.x - d.x, c.y - d.y);
      p[0] = make_double2(ab0.x + cd0.x, ab0.y + cd0.y);
      p[2] = make_double2(ab0.x - cd0.x, ab0.y - cd0.y);
      p[1] = make_double2(ab1.x + cd1.y, ab1.y - cd1.x);
      p[3] = make_double2(ab1.x - cd1.y, ab1.y + cd1.x);
    }
    WB_SYNC();
    s = 3;
  }
  for (; s + 1 <= lg; s += 2) {
    // stages s and s+1: quarter = 2^(s-1); group of 4*quarter elements
    const int quarter = 1 << (s - 1);
    const int tws = WB_TW_LOG2 - (s + 1);
    WB_UNROLL4
    for (int q = tid; q < (n >> 2); q += nth) {
      const int j = q & (quarter - 1);
      const int i0 = ((q >> (s - 1)) << (s + 1)) + j;
      const int i1 = i0 + quarter, i2 = i1 + quarter, i3 = i2 + quarter;
      const double2 b = __ldg(&tw[j << tws]);                       // exp(-j 2 pi j / (4 quarter))
      const double2 a = make_double2(fma(b.x, b.x, -(b.y * b.y)), 2.0 * b.x * b.y);  // its square
      const double2 z0 = z[i0], z1 = z[i1], z2 = z[i2], z3 = z[i3];
      const double t1r = fma(a.x, z1.x, -(a.y * z1.y)), t1i = fma(a.x, z1.y, a.y * z1.x);
      const double t3r = fma(a.x, z3.x, -(a.y * z3.y)), t3i = fma(a.x, z3.y, a.y * z3.x);
      const double u0r = z0.x + t1r, u0i = z0.y + t1i, u1r = z0.x - t1r, u1i = z0.y - t1i;
      const double u2r = z2.x + t3r, u2i = z2.y + t3i, u3r = z2.x - t3r, u3i = z2.y - t3i;
      const double v2r = fma(b.x, u2r, -(b.y * u2i)), v2i = fma(b.x, u2i, b.y * u2r);   // b * u2
      const double w3r = fma(b.x, u3r, -(b.y * u3i)), w3i = fma(b.x, u3i, b.y * u3r);   // b * u3
      // (-j b) u3 = (w3i, -w3r)
      z[i0] = make_double2(u0r + v2r, u0i + v2i);
      z[i2] = make_double2(u0r - v2r, u0i - v2i);
      z[i1] = make_double2(u1r + w3i, u1i - w3r);
      z[i3] = make_double2(u1r - w3i, u1i + w3r);
    }
    WB_SYNC();
  }
  if (s <= lg) {
    const int half = 1 << (s - 1);
    const int tws = WB_TW_LOG2 - s;
    WB_UNROLL4
    for (int b = tid; b < (n >> 1); b += nth) {
      const int j = b & (half - 1);
      const int i0 = ((b >> (s - 1)) << s) + j;
      const int i1 = i0 + half;
      const double2 w = __ldg(&tw[j << tws]);
      const double2 u = z[i0], v = z[i1];
      const double tr = fma(w.x, v.x, -(w.y * v.y));
      const double ti = fma(w.x, v.y, w.y * v.x);
      z[i0] = make_double2(u.x + tr, u.y + ti);
      z[i1] = make_double2(u.x - tr, u.y - ti);
    }
    WB_SYNC();
  }
}

// Forward real FFT of buf[0..N), N = 2^lg >= 4, in place: on return buf holds N/2+1 complex
// values (buf must have room for N+2 doubles).  Ends with a barrier.
WB_DEV void rfft_forward(double *buf, int lg, const double2 *__restrict__ tw) {
  const int tid = WB_TID, nth = WB_NTH;
  double2 *z = reinterpret_cast<double2 *>(buf);
  const int m = 1 << (lg - 1);
  cfft_forward(z, lg - 1, tw);
  const int tws = WB_TW_LOG2 - lg;
  WB_UNROLL4
  for (int k = tid; k <= (m >> 1); k += nth) {
    if (k == 0) {
      const double2 z0 = z[0];
      z[0] = make_double2(z0.x + z0.y, 0.0);
      z[m] = make_double2(z0.x - z0.y, 0.0);
    } else {
      const double2 a = z[k], b = z[m - k];
      const double er = 0.5 * (a.x + b.x), ei = 0.5 * (a.y - b.y);
      const double orr = 0.5 * (a.y + b.y), oi = -0.5 * (a.x - b.x);
      const double2 w = __ldg(&tw[k << tws]);
      const double pr = fma(w.x, orr, -(w.y * oi));
      const double pi = fma(w.x, oi, w.y * orr);
      z[k] = make_double2(er + pr, ei + pi);
      z[m - k] = make_double2(er - pr, -(ei - pi));
    }
  }
  WB_SYNC();
}

// Two real FFTs from one complex FFT: z[n] = a[n] + j b[n] has been transformed in place by
// cfft_forward (size n = 2^lg); returns A[k] and B[k] for 0 <= k <= n/2.
WB_DEV void split_pair(const double2 *z, int n, int k, double2 &A, double2 &B) {
  const double2 p = z[k], q = z[(n - k) & (n - 1)];
  A = make_double2(0.5 * (p.x + q.x), 0.5 * (p.y - q.y));
  B = make_double2(0.5 * (p.y + q.y), -0.5 * (p.x - q.x));
}

}  // namespace wb
