// wb_matlab.cu -- host-only: the public MATLAB-style helpers of the reference (matlabfunctions.h:21-149) that
// callers use around the analysis API.  Same results, bit for bit; nothing here is on the device path (the
// kernels have their own interp1 / decimate / randn, see wb_spectral.cuh, wb_f0common.cu, wb_rng.cu).
#include "../../include/world/matlabfunctions.h"
#include "wb_f0common.cuh"   // decimate_coefficients(): the same table the device decimator uses
#include <math.h>
#include <vector>

extern "C" {

void fftshift(const double *x, int x_length, double *y) {                       // matlabfunctions.cpp:129-134
  const int half = x_length / 2;
  for (int i = 0; i < half; ++i) { y[i] = x[i + half]; y[i + half] = x[i]; }
}

void histc(const double *x, int x_length, const double *edges, int edges_length, int *index) {   // :136-155
  // edges below x[0] land in bin 1; then the bin number follows the ascending edges; once the last
  // node is reached everything that remains belongs to bin x_length - 1
  int i = 0, bin = 1;
  for (; i < edges_length; ++i) {
    index[i] = 1;
    if (edges[i] >= x[0]) break;
  }
  while (i < edges_length) {
    if (edges[i] < x[bin]) {
      index[i++] = bin;
    } else {
      index[i] = bin++;          // re-examined against the next node unless that was the last one
      if (bin == x_length) { ++i; break; }
      continue;
    }
    if (bin == x_length) break;
  }
  for (; i < edges_length; ++i) index[i] = x_length - 1;
}

void interp1(const double *x, const double *y, int x_length, const double *xi, int xi_length, double *yi) {   // :157-176
  std::vector<int> k(xi_length > 0 ? xi_length : 1, 0);
  histc(x, x_length, xi, xi_length, k.data());
  for (int i = 0; i < xi_length; ++i) {
    const int a = k[i] - 1, b = k[i];
    const double s = (xi[i] - x[a]) / (x[b] - x[a]);
    yi[i] = y[a] + s * (y[b] - y[a]);
  }
}

static void iir_pass(const std::vector<double> &in, const double a[3], const double b[2], std::vector<double> *out) {   // :113-122
  double w0 = 0.0, w1 = 0.0, w2 = 0.0;
  out->resize(in.size());
  for (size_t i = 0; i < in.size(); ++i) {
    const double wt = in[i] + a[0] * w0 + a[1] * w1 + a[2] * w2;
    (*out)[i] = b[0] * wt + b[1] * w0 + b[1] * w1 + b[0] * w2;
    w2 = w1; w1 = w0; w0 = wt;
  }
}

void decimate(const double *x, int x_length, int r, double *y) {                // :178-204
  const int pad = 9;
  double a[3], b[2];
  wb::decimate_coefficients(r, a, b);
  std::vector<double> ext((size_t)x_length + 2 * pad), tmp;
  for (int i = 0; i < pad; ++i) ext[i] = 2 * x[0] - x[pad - i];
  for (int i = 0; i < x_length; ++i) ext[pad + i] = x[i];
  for (int i = 0; i < pad; ++i) ext[pad + x_length + i] = 2 * x[x_length - 1] - x[x_length - 2 - i];
  for (int pass = 0; pass < 2; ++pass) {       // forward, reverse, forward, reverse: zero phase
    iir_pass(ext, a, b, &tmp);
    for (size_t i = 0; i < ext.size(); ++i) ext[i] = tmp[ext.size() - 1 - i];
  }
  const int nout = (x_length - 1) / r + 1;
  const int nbeg = r - r * nout + x_length;
  int count = 0;
  for (int i = nbeg; i < x_length + pad; i += r) y[count++] = ext[i + pad - 1];
}

int matlab_round(double x) { return x > 0 ? static_cast<int>(x + 0.5) : static_cast<int>(x - 0.5); }   // :206-208

void diff(const double *x, int x_length, double *y) {                           // :210-212
  for (int i = 0; i + 1 < x_length; ++i) y[i] = x[i + 1] - x[i];
}

void interp1Q(double x, double shift, const double *y, int x_length, const double *xi, int xi_length,
              double *yi) {                                                      // :214-235
  for (int i = 0; i < xi_length; ++i) {
    const int base = static_cast<int>((xi[i] - x) / shift);
    const double frac = (xi[i] - x) / shift - base;
    const double slope = base == x_length - 1 ? 0.0 : y[base + 1] - y[base];
    yi[i] = y[base] + slope * frac;
  }
}

void randn_reseed(RandnState *s) {                                               // :237-242
  s->g_randn_x = 123456789u; s->g_randn_y = 362436069u; s->g_randn_z = 521288629u; s->g_randn_w = 88675123u;
}

double randn(RandnState *s) {                                                    // :244-264
  uint32_t sum = 0;
  for (int i = 0; i < 12; ++i) {
    const uint32_t t = s->g_randn_x ^ (s->g_randn_x << 11);
    s->g_randn_x = s->g_randn_y; s->g_randn_y = s->g_randn_z; s->g_randn_z = s->g_randn_w;
    s->g_randn_w = (s->g_randn_w ^ (s->g_randn_w >> 19)) ^ (t ^ (t >> 8));
    sum += s->g_randn_w >> 4;
  }
  return sum / 268435456.0 - 6.0;
}

double matlab_std(const double *x, int x_length) {                               // :303-313
  double mean = 0.0;
  for (int i = 0; i < x_length; ++i) mean += x[i];
  mean /= x_length;
  double s = 0.0;
  for (int i = 0; i < x_length; ++i) s += pow(x[i] - mean, 2.0);
  s /= (x_length - 1);
  return sqrt(s);
}

}  // extern "C"
