// oracle/world_oracle.cpp -- CPU restatement of the WORLD analysis path.  TEST INFRASTRUCTURE ONLY.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library; the
// product (world_b200/) never links, loads or calls it.  It restates, in plain single-threaded
// C++ and with its own textbook radix-2 FFT, what the reference computes, each function citing the
// reference lines it follows (mmorise/World @ d625e76).  It exists beside oracle/_ref (the
// unmodified reference compiled from /root/reference) so that the checker does not depend on the
// reference tree being present, and so that the algorithm cards of SURVEY.md App. A are executable.
//
// Pinning: the reference ships no golden outputs.  This restatement is pinned against
// tests/golden/vaiueo2d.npz (generated from oracle/_ref on the reference's own fixture) and against
// oracle/_ref directly on synthetic batches, by tests/test_oracle.py.
//
// Restated here: randn, interp1Q, interp1, DCCorrection, LinearSmoothing, NuttallWindow, the FFT
// conventions, StoneMask, CheapTrick, D4C, the option / sizing helpers, and (rows f2 / f3 of SURVEY.md 8)
// the codec of codec.cpp and the PCM sample conversion of tools/audioio.cpp; Dio (with decimate) as
// direct time-domain filtering instead of the reference's FFT convolutions; Harvest likewise (band-pass
// FIRs in the time domain, the instantaneous-frequency refinement with full FFTs like the reference).
// Synthesis (row f1) too: pulse time base, minimum-phase responses, noise responses, overlap-add.
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>

namespace {

const double kPi = 3.1415926535897932384;        // constantnumbers.h:18
const double kTiny = 0.000000000001;             // kMySafeGuardMinimum :19
const double kEps = 0.00000000000000022204460492503131;  // :20
const double kLog2 = 0.69314718055994529;        // :24

int RoundHalfAway(double x) { return x > 0 ? (int)(x + 0.5) : (int)(x - 0.5); }  // matlabfunctions.cpp:206-208

// ---- randn: xorshift128, 12 steps per draw (matlabfunctions.cpp:237-264)
struct Rng {
  uint32_t x, y, z, w;
  Rng() : x(123456789), y(362436069), z(521288629), w(88675123) {}
  double Next() {
    uint32_t acc = 0;
    for (int i = 0; i < 12; ++i) {
      uint32_t t = x ^ (x << 11);
      x = y; y = z; z = w;
      w = (w ^ (w >> 19)) ^ (t ^ (t >> 8));
      acc += w >> 4;
    }
    return acc / 268435456.0 - 6.0;
  }
};

// ---- FFT: r2c convention of fft.cpp:49-60 (X[k] = sum x[n] exp(-j 2 pi k n / N)), textbook
// iterative radix-2 on a complex copy of the real input.
void RealFFT(const std::vector<double> &x, std::vector<double> *re, std::vector<double> *im) {
  const int n = (int)x.size();
  std::vector<double> ar(x), ai(n, 0.0);
  for (int i = 1, j = 0; i < n; ++i) {
    int bit = n >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) std::swap(ar[i], ar[j]);
  }
  static std::vector<double> tw_r, tw_i;     // exp(-j 2 pi k / n) for the largest n seen (single-threaded use)
  static int tw_n = 0;
  if (tw_n < n) {
    tw_n = n;
    tw_r.resize(n / 2); tw_i.resize(n / 2);
    for (int k = 0; k < n / 2; ++k) { tw_r[k] = cos(-2.0 * kPi * k / n); tw_i[k] = sin(-2.0 * kPi * k / n); }
  }
  for (int len = 2; len <= n; len <<= 1) {
    const int stride = tw_n / len;
    for (int i = 0; i < n; i += len)
      for (int k = 0; k < len / 2; ++k) {
        const double wr = tw_r[k * stride], wi = tw_i[k * stride];
        const int a = i + k, b = i + k + len / 2;
        const double tr = ar[b] * wr - ai[b] * wi, ti = ar[b] * wi + ai[b] * wr;
        ar[b] = ar[a] - tr; ai[b] = ai[a] - ti;
        ar[a] += tr; ai[a] += ti;
      }
  }
  re->assign(ar.begin(), ar.begin() + n / 2 + 1);
  im->assign(ai.begin(), ai.begin() + n / 2 + 1);
}

// ---- interp1Q (matlabfunctions.cpp:214-235)
double Interp1QAt(double x0, double dx, const std::vector<double> &y, int ny, double xi) {
  const double r = (xi - x0) / dx;
  const int base = (int)r;
  const double dy = base + 1 < ny ? y[base + 1] - y[base] : 0.0;
  return y[base] + dy * (r - base);
}

// ---- interp1 + histc for sorted query points (matlabfunctions.cpp:136-176)
double Interp1At(const std::vector<double> &x, const std::vector<double> &y, double xi) {
  const int n = (int)x.size();
  int k = 0;
  while (k < n && x[k] <= xi) ++k;
  k = std::min(n - 1, std::max(1, k));
  const double s = (xi - x[k - 1]) / (x[k] - x[k - 1]);
  return y[k - 1] + s * (y[k] - y[k - 1]);
}

// ---- DCCorrection (common.cpp:56-75), in place
void DCCorrect(std::vector<double> *spec, double f0, int fs, int fft_size) {
  const int upper = 2 + (int)(f0 * fft_size / fs);
  std::vector<double> rep(upper - 1);
  for (int i = 0; i < upper - 1; ++i)
    rep[i] = Interp1QAt(f0, -(double)fs / fft_size, *spec, upper + 1, (double)i * fs / fft_size);
  for (int i = 0; i < upper - 1; ++i) (*spec)[i] += rep[i];
}

// ---- LinearSmoothing (common.cpp:27-46, 77-111): index-order running sum is essential (App. B4)
void LinearSmooth(const std::vector<double> &in, double width, int fs, int fft_size, std::vector<double> *out) {
  const int half = fft_size / 2;
  const int boundary = (int)(width * fft_size / fs) + 1;
  const int n = half + boundary * 2 + 1;
  std::vector<double> seg(n);
  for (int i = 0; i < n; ++i) {
    double v;
    if (i < boundary) v = in[boundary - i];
    else if (i < half + boundary) v = in[i - boundary];
    else v = in[half - (i - (half + boundary))];
    seg[i] = v * fs / fft_size + (i ? seg[i - 1] : 0.0);
  }
  const double origin = -(boundary - 0.5) * fs / fft_size, dx = (double)fs / fft_size;
  std::vector<double> res(half + 1);
  for (int i = 0; i <= half; ++i) {
    const double lo = (double)i / fft_size * fs - width / 2.0, hi = lo + width;
    res[i] = (Interp1QAt(origin, dx, seg, n, hi) - Interp1QAt(origin, dx, seg, n, lo)) / width;
  }
  *out = res;
}

double SampleAt(const double *x, int n, int idx) { return x[std::min(n - 1, std::max(0, idx))]; }

}  // namespace

extern "C" {

typedef struct { double q1; double f0_floor; int fft_size; } CheapTrickOption;   // cheaptrick.h:16-20
typedef struct { double threshold; } D4COption;                                   // d4c.h:16-18

int GetFFTSizeForCheapTrick(int fs, const CheapTrickOption *o) {                  // cheaptrick.cpp:191-194
  return (int)pow(2.0, 1.0 + (int)(log(3.0 * fs / o->f0_floor + 1) / kLog2));
}
double GetF0FloorForCheapTrick(int fs, int fft_size) { return 3.0 * fs / (fft_size - 3.0); }  // :196-198
void InitializeCheapTrickOption(int fs, CheapTrickOption *o) {                    // :231-240
  o->q1 = -0.15; o->f0_floor = 71.0; o->fft_size = GetFFTSizeForCheapTrick(fs, o);
}
void InitializeD4COption(D4COption *o) { o->threshold = 0.85; }                   // d4c.cpp:405-407
int GetSamplesForDIO(int fs, int n, double fp) { return (int)(1000.0 * n / fs / fp) + 1; }      // dio.cpp:639-641
int GetSamplesForHarvest(int fs, int n, double fp) { return (int)(1000.0 * n / fs / fp) + 1; }  // harvest.cpp:1219-1221

// ---- CheapTrick (cheaptrick.cpp:200-229; card SURVEY.md A1)
void CheapTrick(const double *x, int x_length, int fs, const double *t, const double *f0, int f0_length,
                const CheapTrickOption *opt, double **spectrogram) {
  const int N = opt->fft_size, half = N / 2;
  const double floor_f0 = GetF0FloorForCheapTrick(fs, N);
  Rng rng;                                                          // :205-206
  for (int i = 0; i < f0_length; ++i) {
    const double f = f0[i] <= floor_f0 ? 500.0 : f0[i];             // :218
    const int h = RoundHalfAway(1.5 * fs / f);                      // :115
    const int origin = RoundHalfAway(t[i] * fs + 0.001);            // :92
    std::vector<double> w(2 * h + 1), v(N, 0.0);
    double e = 0.0;
    for (int j = 0; j <= 2 * h; ++j) {                              // :97-106
      const double pos = (j - h) / 1.5 / fs;
      w[j] = 0.5 * cos(kPi * pos * f) + 0.5;
      e += w[j] * w[j];
    }
    e = sqrt(e);
    double s1 = 0.0, s2 = 0.0;
    for (int j = 0; j <= 2 * h; ++j) {                              // :126-137
      w[j] /= e;
      v[j] = SampleAt(x, x_length, origin + j - h) * w[j] + rng.Next() * kTiny;
      s1 += v[j]; s2 += w[j];
    }
    for (int j = 0; j <= 2 * h; ++j) v[j] -= w[j] * (s1 / s2);
    std::vector<double> re, im, p(half + 1);
    RealFFT(v, &re, &im);                                           // :71-78
    for (int k = 0; k <= half; ++k) p[k] = re[k] * re[k] + im[k] * im[k];
    DCCorrect(&p, f, fs, N);                                        // :81
    LinearSmooth(p, f * 2.0 / 3.0, fs, N, &p);                      // :176-177
    std::vector<double> l(N);
    for (int k = 0; k <= half; ++k) {                               // :147-151, :39-42
      p[k] += fabs(rng.Next()) * kEps;
      l[k] = log(p[k]);
    }
    for (int k = 1; k < half; ++k) l[N - k] = l[k];
    RealFFT(l, &re, &im);                                           // :43
    std::vector<double> c(N);
    for (int k = 0; k <= half; ++k) {                               // :28-37, :45-49
      double sl = 1.0, cl = (1.0 - 2.0 * opt->q1) + 2.0 * opt->q1;
      if (k > 0) {
        const double q = (double)k / fs;
        sl = sin(kPi * f * q) / (kPi * f * q);
        cl = (1.0 - 2.0 * opt->q1) + 2.0 * opt->q1 * cos(2.0 * kPi * q * f);
      }
      c[k] = re[k] * sl * cl / N;
    }
    for (int k = 1; k < half; ++k) c[N - k] = c[k];
    RealFFT(c, &re, &im);   // c2r of a real even spectrum == Re r2c(mirror)   (:50, fft.cpp:26-35)
    for (int k = 0; k <= half; ++k) spectrogram[i][k] = exp(re[k]);  // :52-53
  }
}

// ---- D4C helpers (d4c.cpp:21-83): F0-adaptive window + noise + weighted mean removal
static int D4CWindowed(const double *x, int n, int fs, double f, double pos, int type, double ratio, Rng *rng,
                       std::vector<double> *v) {
  const int h = RoundHalfAway(ratio * fs / f / 2.0);
  const int origin = RoundHalfAway(pos * fs + 0.001);
  std::vector<double> w(2 * h + 1);
  double s1 = 0.0, s2 = 0.0;
  for (int j = 0; j <= 2 * h; ++j) {
    const double p = (2.0 * (j - h) / ratio) / fs;
    w[j] = type == 1 ? 0.5 * cos(kPi * p * f) + 0.5 : 0.42 + 0.5 * cos(kPi * p * f) + 0.08 * cos(kPi * p * f * 2);
    (*v)[j] = SampleAt(x, n, origin + j - h) * w[j] + rng->Next() * 0.000001;
    s1 += (*v)[j]; s2 += w[j];
  }
  for (int j = 0; j <= 2 * h; ++j) (*v)[j] -= w[j] * (s1 / s2);
  return 2 * h + 1;
}

// ---- D4C (d4c.cpp:342-403; card SURVEY.md A2)
void D4C(const double *x, int x_length, int fs, const double *t, const double *f0, int f0_length, int fft_size,
         const D4COption *opt, double **aperiodicity) {
  const int bins = fft_size / 2 + 1;
  Rng rng;
  for (int i = 0; i < f0_length; ++i)
    for (int k = 0; k < bins; ++k) aperiodicity[i][k] = 1.0 - kTiny;                  // :323-328
  const int Nd = (int)pow(2.0, 1.0 + (int)(log(4.0 * fs / 47.0 + 1) / kLog2));       // :350-352
  const int Nl = (int)pow(2.0, 1.0 + (int)(log(3.0 * fs / 40.0 + 1) / kLog2));       // :262-264
  const int n_ap = (int)(std::min(15000.0, fs / 2.0 - 3000.0) / 3000.0);             // :357-359
  const int wlen = (int)(3000.0 * Nd / fs) * 2 + 1;                                   // :362-363
  std::vector<double> nuttall(wlen);
  for (int i = 0; i < wlen; ++i) {                                                    // common.cpp:113-121
    const double u = i / (wlen - 1.0);
    nuttall[i] = 0.355768 - 0.487396 * cos(2.0 * kPi * u) + 0.144232 * cos(4.0 * kPi * u) - 0.012604 * cos(6.0 * kPi * u);
  }
  // pass A: LoveTrain (:260-285, :227-252)
  const int b0 = (int)ceil(100.0 * Nl / fs), b1 = (int)ceil(4000.0 * Nl / fs), b2 = (int)ceil(7900.0 * Nl / fs);
  std::vector<double> ap0(f0_length, 0.0), re, im;
  for (int i = 0; i < f0_length; ++i) {
    if (f0[i] == 0.0) continue;
    std::vector<double> v(Nl, 0.0);
    D4CWindowed(x, x_length, fs, std::max(f0[i], 40.0), t[i], 2, 3.0, &rng, &v);
    RealFFT(v, &re, &im);
    double lo = 0.0, hi = 0.0;
    for (int k = b0 + 1; k <= std::min(b2, Nl / 2); ++k) {
      const double p = re[k] * re[k] + im[k] * im[k];
      hi += p;
      if (k == b1) lo = hi;
    }
    ap0[i] = lo / hi;
  }
  // pass B (:385-395, :293-321)
  std::vector<double> axis(n_ap + 2), coarse(n_ap + 2);
  for (int i = 0; i <= n_ap; ++i) axis[i] = i * 3000.0;
  axis[n_ap + 1] = fs / 2.0;
  coarse[0] = -60.0; coarse[n_ap + 1] = -kTiny;
  const int half = Nd / 2, bd = RoundHalfAway(Nd * 8.0 / wlen), hw = wlen / 2;
  for (int i = 0; i < f0_length; ++i) {
    if (f0[i] == 0 || ap0[i] <= opt->threshold) continue;
    const double f = std::max(47.0, f0[i]);
    std::vector<double> cent(half + 1, 0.0);
    for (int pass = 0; pass < 2; ++pass) {                                            // GetCentroid :90-120
      std::vector<double> v(Nd, 0.0), r2, i2;
      const int nw = D4CWindowed(x, x_length, fs, f, pass == 0 ? t[i] - 0.25 / f : t[i] + 0.25 / f, 2, 4.0, &rng, &v);
      double pw = 0.0;
      for (int j = 0; j < nw; ++j) pw += v[j] * v[j];
      for (int j = 0; j < nw; ++j) v[j] /= sqrt(pw);
      RealFFT(v, &re, &im);
      for (int j = 0; j < Nd; ++j) v[j] *= j + 1.0;
      RealFFT(v, &r2, &i2);
      for (int k = 0; k <= half; ++k) cent[k] += r2[k] * re[k] + im[k] * i2[k];
    }
    DCCorrect(&cent, f, fs, Nd);                                                      // :137-140
    std::vector<double> v(Nd, 0.0), pwr(half + 1);                                    // :149-166
    D4CWindowed(x, x_length, fs, f, t[i], 1, 4.0, &rng, &v);
    RealFFT(v, &re, &im);
    for (int k = 0; k <= half; ++k) pwr[k] = re[k] * re[k] + im[k] * im[k];
    DCCorrect(&pwr, f, fs, Nd);
    LinearSmooth(pwr, f, fs, Nd, &pwr);
    std::vector<double> g(half + 1), g2;                                              // :172-188
    for (int k = 0; k <= half; ++k) g[k] = cent[k] / pwr[k];
    LinearSmooth(g, f / 2.0, fs, Nd, &g);
    LinearSmooth(g, f, fs, Nd, &g2);
    for (int k = 0; k <= half; ++k) g[k] -= g2[k];
    for (int b = 0; b < n_ap; ++b) {                                                  // :194-225
      const int center = (int)(3000.0 * (b + 1) * Nd / fs);
      std::vector<double> u(Nd, 0.0), ps(half + 1);
      for (int j = 0; j <= hw * 2; ++j) u[j] = g[center - hw + j] * nuttall[j];
      RealFFT(u, &re, &im);
      for (int k = 0; k <= half; ++k) ps[k] = re[k] * re[k] + im[k] * im[k];
      std::sort(ps.begin(), ps.end());
      for (int k = 1; k <= half; ++k) ps[k] += ps[k - 1];
      coarse[b + 1] = std::min(0.0, 10 * log10(ps[half - bd - 1] / ps[half]) + (f - 100) / 50.0);  // :314-316
    }
    for (int k = 0; k < bins; ++k)                                                    // :330-338
      aperiodicity[i][k] = pow(10.0, Interp1At(axis, coarse, (double)k * fs / fft_size) / 20.0);
  }
}

// ---- StoneMask (stonemask.cpp:212-218; card SURVEY.md A4)
static double FixF0(const std::vector<double> &pw, const std::vector<double> &num, int N, int fs, double f, int H) {
  double numer = 0.0, denom = 0.0;                                                    // :96-118
  for (int m = 0; m < H; ++m) {
    const int k = std::min(RoundHalfAway(f * N / fs * (m + 1)), N / 2);
    const double inst = pw[k] == 0.0 ? 0.0 : (double)k * fs / N + num[k] / pw[k] * fs / 2.0 / kPi;
    const double amp = sqrt(pw[k]);
    numer += amp * inst;
    denom += amp * (m + 1);
  }
  return numer / (denom + kTiny);
}

void StoneMask(const double *x, int x_length, int fs, const double *t, const double *f0, int f0_length,
               double *refined) {
  for (int i = 0; i < f0_length; ++i) {
    const double f = f0[i];
    if (f <= 40.0 || f > fs / 12.0) { refined[i] = 0.0; continue; }                   // :187-188
    const int h = (int)(1.5 * fs / f + 1.0), n = 2 * h + 1;
    const int N = (int)pow(2.0, 2.0 + (int)(log(n * 1.0) / kLog2));                    // :195-196
    const double T = (2.0 * h + 1.0) / fs;
    std::vector<int> raw(n);
    std::vector<double> w(n), dw(n), a(N, 0.0), d(N, 0.0);
    for (int j = 0; j < n; ++j) {                                                     // :24-43
      raw[j] = RoundHalfAway((t[i] + (double)(j - h) / fs) * fs);
      const double tau = (raw[j] - 1.0) / fs - t[i];
      w[j] = 0.42 + 0.5 * cos(2.0 * kPi * tau / T) + 0.08 * cos(4.0 * kPi * tau / T);
    }
    dw[0] = -w[1] / 2.0;                                                              // :49-55
    for (int j = 1; j < n - 1; ++j) dw[j] = -(w[j + 1] - w[j - 1]) / 2.0;
    dw[n - 1] = w[n - 2] / 2.0;
    for (int j = 0; j < n; ++j) {                                                     // :61-88
      const double s = SampleAt(x, x_length, raw[j] - 1);
      a[j] = s * w[j]; d[j] = s * dw[j];
    }
    std::vector<double> mr, mi, dr, di, pw(N / 2 + 1), num(N / 2 + 1);
    RealFFT(a, &mr, &mi);
    RealFFT(d, &dr, &di);
    for (int k = 0; k <= N / 2; ++k) {                                                // :159-164
      num[k] = mr[k] * di[k] - mi[k] * dr[k];
      pw[k] = mr[k] * mr[k] + mi[k] * mi[k];
    }
    double mean = 0.0;
    const double tent = FixF0(pw, num, N, fs, f, 2);                                  // :123-132
    if (!(tent <= 0.0 || tent > f * 2)) mean = FixF0(pw, num, N, fs, tent, 6);
    if (fabs(mean - f) > f * 0.2) mean = f;                                           // :203
    refined[i] = mean;
  }
}

// ---- codec (codec.cpp).  The "DCT" there is one real FFT of the even/odd re-ordered mel spectrum
// times a unit phasor; the inverse is written out here as the plain sum it stands for.
static double MelOf(double f) { return 1127.01048 * log(f / 700.0 + 1.0); }        // codec.cpp:60-62, constantnumbers.h:45-46
static double HzOf(double m) { return 700.0 * (exp(m / 1127.01048) - 1.0); }      // :67-69

int GetNumberOfAperiodicities(int fs) {                                            // :216-219
  return (int)(std::min(15000.0, fs / 2.0 - 3000.0) / 3000.0);
}

void CodeAperiodicity(const double *const *ap, int f0_length, int fs, int fft_size, double **coded) {  // :221-240
  const int n_ap = GetNumberOfAperiodicities(fs), bins = fft_size / 2 + 1;
  std::vector<double> db(bins);
  for (int i = 0; i < f0_length; ++i) {
    for (int j = 0; j < bins; ++j) db[j] = 20 * log10(ap[i][j]);
    for (int b = 0; b < n_ap; ++b) coded[i][b] = Interp1QAt(0, (double)fs / fft_size, db, bins, 3000.0 * (b + 1.0));
  }
}

void DecodeAperiodicity(const double *const *coded, int f0_length, int fs, int fft_size, double **ap) {  // :242-272
  const int n_ap = GetNumberOfAperiodicities(fs), bins = fft_size / 2 + 1;
  std::vector<double> axis(n_ap + 2), val(n_ap + 2);
  for (int i = 0; i <= n_ap; ++i) axis[i] = i * 3000.0;
  axis[n_ap + 1] = fs / 2.0;
  val[0] = -60.0;
  val[n_ap + 1] = -kTiny;
  for (int i = 0; i < f0_length; ++i) {
    double mean = 0.0;
    for (int b = 0; b < n_ap; ++b) { mean += coded[i][b]; val[b + 1] = coded[i][b]; }
    mean /= n_ap;
    for (int j = 0; j < bins; ++j)
      ap[i][j] = mean > -0.5 ? 1.0 - kTiny                                          // unvoiced frame (:21-41, :263-264)
                             : pow(10.0, Interp1At(axis, val, (double)fs / fft_size * j) / 20.0);
  }
}

void CodeSpectralEnvelope(const double *const *sp, int f0_length, int fs, int fft_size, int dims, double **coded) {  // :274-301
  const int M = fft_size / 2;
  const double floor_mel = MelOf(40.0), ceil_mel = MelOf(std::min(fs / 2.0, 20000.0));
  std::vector<double> lin_axis(M + 1), lg(M + 1), v(M), re, im;
  for (int j = 0; j <= M; ++j) lin_axis[j] = MelOf((double)j * fs / fft_size);      // :177-180
  for (int i = 0; i < f0_length; ++i) {
    for (int j = 0; j <= M; ++j) lg[j] = log(sp[i][j]);
    std::vector<double> mel(M);
    for (int m = 0; m < M; ++m) mel[m] = Interp1At(lin_axis, lg, (ceil_mel - floor_mel) * m / M + floor_mel);  // :121-122, :169-170
    for (int m = 0; m < M / 2; ++m) { v[m] = mel[2 * m]; v[m + M / 2] = mel[M - 2 * m - 1]; }                   // :77-81
    RealFFT(v, &re, &im);
    for (int d = 0; d < dims; ++d) {                                                                            // :85-88, :171-175
      double wr = 2.0 * cos(d * kPi / fft_size) / sqrt((double)fft_size);
      const double wi = 2.0 * sin(d * kPi / fft_size) / sqrt((double)fft_size);
      if (d == 0) wr /= sqrt(2.0);
      coded[i][d] = (re[d] * wr - im[d] * wi) / sqrt((double)M);
    }
  }
}

void DecodeSpectralEnvelope(const double *const *coded, int f0_length, int fs, int fft_size, int dims, double **sp) {  // :303-324
  const int M = fft_size / 2;
  const double floor_mel = MelOf(40.0), ceil_mel = MelOf(std::min(fs / 2.0, 20000.0));
  std::vector<double> axis(M + 2), mel(M + 2), out(M);
  axis[0] = 0;
  for (int m = 0; m < M; ++m) axis[m + 1] = HzOf((ceil_mel - floor_mel) * m / M + floor_mel);                    // :202-206
  axis[M + 1] = fs / 2.0;
  for (int i = 0; i < f0_length; ++i) {
    // Re conj(FFT(a)) = Re FFT(a) with a[d] = c[d] (wr - j wi) sqrt(M), zero beyond dims (:97-109, fft.cpp:36-46)
    for (int n = 0; n < M; ++n) {
      double acc = 0.0;
      for (int d = 0; d < dims; ++d) {
        double wr = cos(d * kPi / fft_size) * sqrt((double)fft_size);
        const double wi = sin(d * kPi / fft_size) * sqrt((double)fft_size);
        if (d == 0) wr /= sqrt(2.0);
        const double ar = coded[i][d] * wr * sqrt((double)M), ai = -coded[i][d] * wi * sqrt((double)M);
        const double ang = -2.0 * kPi * d * n / M;
        acc += ar * cos(ang) - ai * sin(ang);
      }
      out[n] = acc;
    }
    for (int m = 0; m < M / 2; ++m) { mel[1 + 2 * m] = out[m]; mel[2 + 2 * m] = out[M - m - 1]; }               // :111-115
    mel[0] = mel[1];
    mel[M + 1] = mel[M];
    for (int j = 0; j <= M; ++j) sp[i][j] = exp(Interp1At(axis, mel, (double)j * fs / fft_size) / M);           // :150-154
  }
}

// ---- wavread's sample conversion (tools/audioio.cpp:236-249): little-endian signed PCM -> [-1, 1)
void OraclePcmToDouble(const unsigned char *pcm, int nbit, int n, double *x) {
  const int nb = nbit / 8;
  const double zero_line = pow(2.0, nbit - 1);
  for (int i = 0; i < n; ++i) {
    const unsigned char *s = pcm + (size_t)i * nb;
    double tmp = 0.0, bias = 0.0;
    unsigned char top = s[nb - 1];
    if (top >= 128) { bias = zero_line; top &= 0x7F; }
    tmp = top;
    for (int j = nb - 2; j >= 0; --j) tmp = tmp * 256.0 + s[j];
    x[i] = (tmp - bias) / zero_line;
  }
}

// ---- decimate (matlabfunctions.cpp:178-204) with FilterForDecimate's IIR (:27-125)
static void DecimateCoefficients(int r, double a[3], double b[2]) {
  static const double kTable[11][5] = {   // rows r = 2 .. 12: a0 a1 a2 b0 b1
    {0.041156734567757189, -0.42599112459189636, 0.041037215479961225, 0.16797464681802227, 0.50392394045406674},
    {0.95039378983237421, -0.67429146741526791, 0.15412211621346475, 0.071221945171178636, 0.21366583551353591},
    {1.4499664446880227, -0.98943497080950582, 0.24578252340690215, 0.036710750339322612, 0.11013225101796784},
    {1.7610939654280557, -1.2554914843859768, 0.3237186507788215, 0.021334858522387423, 0.06400457556716227},
    {1.9715352749512141, -1.4686795689225347, 0.3893908434965701, 0.013469181309343825, 0.040407543928031475},
    {2.1225239019534703, -1.6395144861046302, 0.44469707800587366, 0.0090366882681608418, 0.027110064804482525},
    {2.2357462340187593, -1.7780899984041358, 0.49152555365968692, 0.0063522763407111993, 0.019056829022133598},
    {2.3236003491759578, -1.8921545617463598, 0.53148928133729068, 0.0046331164041389372, 0.013899349212416812},
    {2.3936475118069387, -1.9873904075111861, 0.5658879979027055, 0.0034818622251927556, 0.010445586675578267},
    {2.450743295230728, -2.06794904601978, 0.59574774438332101, 0.0026822508007163792, 0.0080467524021491377},
    {2.4981398605924205, -2.1368928194784025, 0.62187513816221485, 0.0021097275904709001, 0.0063291827714127002}};
  for (int i = 0; i < 3; ++i) a[i] = (r >= 2 && r <= 12) ? kTable[r - 2][i] : 0.0;
  for (int i = 0; i < 2; ++i) b[i] = (r >= 2 && r <= 12) ? kTable[r - 2][3 + i] : 0.0;
}

static void IirPass(const std::vector<double> &in, int r, std::vector<double> *out) {   // :113-122
  double a[3], b[2], w0 = 0.0, w1 = 0.0, w2 = 0.0;
  DecimateCoefficients(r, a, b);
  out->resize(in.size());
  for (size_t i = 0; i < in.size(); ++i) {
    const double wt = in[i] + a[0] * w0 + a[1] * w1 + a[2] * w2;
    (*out)[i] = b[0] * wt + b[1] * w0 + b[1] * w1 + b[0] * w2;
    w2 = w1; w1 = w0; w0 = wt;
  }
}

static void Decimate(const double *x, int n, int r, std::vector<double> *y) {
  const int pad = 9;
  std::vector<double> ext(n + 2 * pad), tmp;
  for (int i = 0; i < pad; ++i) ext[i] = 2 * x[0] - x[pad - i];                       // odd reflection at both ends
  for (int i = 0; i < n; ++i) ext[pad + i] = x[i];
  for (int i = 0; i < pad; ++i) ext[pad + n + i] = 2 * x[n - 1] - x[n - 2 - i];
  for (int pass = 0; pass < 2; ++pass) {                                              // forward, then backward: zero phase
    IirPass(ext, r, &tmp);
    for (size_t i = 0; i < ext.size(); ++i) ext[i] = tmp[ext.size() - 1 - i];
  }
  const int nout = (n - 1) / r + 1, nbeg = r - r * nout + n;
  y->clear();
  for (int i = nbeg; i < n + pad; i += r) y->push_back(ext[i + pad - 1]);
}

// ---- DIO (dio.cpp).  The reference filters by FFT; every product there is a circular convolution that
// the chosen fft_size keeps free of wrap-around (:590-592), so the same signals are linear convolutions.
typedef struct { double f0_floor, f0_ceil, channels_in_octave, frame_period; int speed; double allowed_range; } DioOption;  // dio.h:16-23

void InitializeDioOption(DioOption *o) {                                            // dio.cpp:650-663
  o->channels_in_octave = 2.0; o->f0_ceil = 800.0; o->f0_floor = 71.0;
  o->frame_period = 5; o->speed = 1; o->allowed_range = 0.1;
}

// ZeroCrossingEngine (:357-401): negative-going crossings, linearly interpolated; interval i lies
// between crossings i and i+1.  Returns the number of intervals.
static int CrossingIntervals(const std::vector<double> &s, int n, double fs, std::vector<double> *loc, std::vector<double> *f) {
  std::vector<double> fine;
  for (int i = 0; i + 1 < n; ++i)
    if (0.0 < s[i] && s[i + 1] <= 0.0) fine.push_back((i + 1) - s[i] / (s[i + 1] - s[i]));
  loc->clear(); f->clear();
  if (fine.size() < 2) return 0;
  for (size_t i = 0; i + 1 < fine.size(); ++i) {
    f->push_back(fs / (fine[i + 1] - fine[i]));
    loc->push_back((fine[i] + fine[i + 1]) / 2.0 / fs);
  }
  return (int)f->size();
}

// interp1 as the reference evaluates it for a whole sorted query vector (matlabfunctions.cpp:136-176)
static void Interp1Vec(const std::vector<double> &x, const std::vector<double> &y, const double *xi, int n, std::vector<double> *yi) {
  yi->resize(n);
  for (int i = 0; i < n; ++i) (*yi)[i] = Interp1At(x, y, xi[i]);
}

static double DioSelect(double cur, double past, const std::vector<std::vector<double> > &cand, int idx, double allowed) {  // :197-216
  const double ref = (cur * 3.0 - past) / 2.0;
  double best = cand[0][idx], err = fabs(ref - best);
  for (size_t b = 1; b < cand.size(); ++b)
    if (fabs(ref - cand[b][idx]) < err) { err = fabs(ref - cand[b][idx]); best = cand[b][idx]; }
  return fabs(1.0 - best / ref) > allowed ? 0.0 : best;
}

void Dio(const double *x, int x_length, int fs, const DioOption *o, double *t, double *f0) {   // :643-648, :578-635
  const int nb = 1 + (int)(log(o->f0_ceil / o->f0_floor) / kLog2 * o->channels_in_octave);
  std::vector<double> boundary(nb);
  for (int i = 0; i < nb; ++i) boundary[i] = o->f0_floor * pow(2.0, (i + 1) / o->channels_in_octave);
  const int ratio = std::max(std::min(o->speed, 12), 1);
  const int ylen = 1 + x_length / ratio;
  const double afs = (double)fs / ratio;
  const int L = GetSamplesForDIO(fs, x_length, o->frame_period);
  for (int i = 0; i < L; ++i) t[i] = i * o->frame_period / 1000.0;

  // GetSpectrumForEstimation (:61-106): decimate, remove the mean over y_length, low-cut filter
  std::vector<double> y(ylen, 0.0), dec;
  if (ratio != 1) { Decimate(x, x_length, ratio, &dec); for (size_t i = 0; i < dec.size() && (int)i < ylen; ++i) y[i] = dec[i]; }
  else for (int i = 0; i < x_length; ++i) y[i] = x[i];
  double mean = 0.0;
  for (int i = 0; i < ylen; ++i) mean += y[i];
  mean /= ylen;
  for (int i = 0; i < ylen; ++i) y[i] -= mean;
  // DesignLowCutFilter (:40-53): delta minus a unit-sum Hann bump of 2c+1 points, centred on lag 0
  const int c = RoundHalfAway(afs / 50.0), N = 2 * c + 1;
  std::vector<double> lc(N);
  double sum = 0.0;
  for (int i = 1; i <= N; ++i) { lc[i - 1] = 0.5 - 0.5 * cos(i * 2.0 * kPi / (N + 1)); sum += lc[i - 1]; }
  for (int i = 0; i < N; ++i) lc[i] = -lc[i] / sum;
  lc[c] += 1.0;
  // s[n], n in [-c, ylen + c): stored with offset c
  std::vector<double> s(ylen + 2 * c, 0.0);
  for (int n = -c; n < ylen + c; ++n) {
    double acc = 0.0;
    for (int k = -c; k <= c; ++k) { const int m = n - k; if (m >= 0 && m < ylen) acc += lc[k + c] * y[m]; }
    s[n + c] = acc;
  }

  // bins N/2 - 1 and N/2 of the low-cut filtered spectrum, N = the reference's fft_size (dio.cpp:590-592)
  const int NF = (int)pow(2.0, (int)(log((double)(ylen + RoundHalfAway(afs / 50.0) * 2 + 1 + 4 * (int)(1.0 + afs / boundary[0] / 2.0))) / kLog2) + 1.0);
  const int N2 = NF / 2;
  double ys1r = 0.0, ys1i = 0.0, ys2 = 0.0;
  for (int n = -c; n < ylen + c; ++n) {
    const double v = s[n + c];
    ys1r += v * cos(2.0 * kPi * (N2 - 1) * n / NF); ys1i -= v * sin(2.0 * kPi * (N2 - 1) * n / NF);
    ys2 += v * ((n & 1) ? -1.0 : 1.0);
  }
  std::vector<std::vector<double> > cand(nb, std::vector<double>(L)), score(nb, std::vector<double>(L));
  std::vector<double> filt(ylen), work, loc[4], itv[4], yi[4];
  for (int b = 0; b < nb; ++b) {
    // GetFilteredSignal (:296-343): Nuttall window of 4h points as low-pass, delay 2h removed
    const int h = RoundHalfAway(afs / boundary[b] / 2.0), M = 4 * h;
    std::vector<double> w(M);
    for (int i = 0; i < M; ++i) {
      const double u = i / (M - 1.0);
      w[i] = 0.355768 - 0.487396 * cos(2.0 * kPi * u) + 0.144232 * cos(4.0 * kPi * u) - 0.012604 * cos(6.0 * kPi * u);
    }
    // F at bins N/2 - 1 and N/2, and the difference between what the mirroring loop leaves there and the product
    double f1r = 0.0, f1i = 0.0, f2 = 0.0;
    for (int k = 0; k < M; ++k) {
      f1r += w[k] * cos(2.0 * kPi * (N2 - 1) * k / NF); f1i -= w[k] * sin(2.0 * kPi * (N2 - 1) * k / NF);
      f2 += w[k] * ((k & 1) ? -1.0 : 1.0);
    }
    const double p_re = ys1r * f1r - ys1i * f1i, p_im = ys1r * f1i + ys1i * f1r;     // Ys[N/2-1] F[N/2-1]
    const double dq_re = ys2 * p_re - p_re, dq_im = ys2 * p_im - p_im;             // Q - product at N/2 - 1
    const double dn = ys2 * p_re - ys2 * f2;                                        // Re Q - product at N/2
    for (int i = 0; i < ylen; ++i) {
      double acc = 0.0;
      for (int k = 0; k < M; ++k) { const int m = i + 2 * h - k; if (m >= -c && m < ylen + c) acc += w[k] * s[m + c]; }
      // The reference's spectral "mirroring" loop (dio.cpp:319-328) writes product bin i to slot N - i - 1; for
      // i = N/2 - 1 and N/2 those slots are N/2 and N/2 - 1, i.e. inside the half c2r reads: both end up as
      // Q = Ys[N/2] * (Ys[N/2-1] F[N/2-1]).  Negligible next to a real signal when the filter is long (F ~ 0 near
      // Nyquist), visible for the 4..12 tap windows of heavy decimation -- and the ONLY thing left in digital
      // silence, where this near-Nyquist ripple gives the reference a zero crossing every sample or two (which is
      // why it calls silence unvoiced instead of extrapolating the last interval).  Added here as the time-domain
      // signal it amounts to.
      acc += (2.0 * (dq_re * cos(2.0 * kPi * (N2 - 1) * (i + 2 * h) / NF) - dq_im * sin(2.0 * kPi * (N2 - 1) * (i + 2 * h) / NF)) +
              dn * (((i + 2 * h) & 1) ? -1.0 : 1.0)) / NF;
      filt[i] = acc;
    }
    // GetFourZeroCrossingIntervals (:410-444): crossings of s, -s, and of the two signs of the difference
    int cnt[4];
    cnt[0] = CrossingIntervals(filt, ylen, afs, &loc[0], &itv[0]);
    work.assign(filt.begin(), filt.end());
    for (int i = 0; i < ylen; ++i) work[i] = -work[i];
    cnt[1] = CrossingIntervals(work, ylen, afs, &loc[1], &itv[1]);
    for (int i = 0; i + 1 < ylen; ++i) work[i] = work[i] - work[i + 1];
    cnt[2] = CrossingIntervals(work, ylen - 1, afs, &loc[2], &itv[2]);
    for (int i = 0; i + 1 < ylen; ++i) work[i] = -work[i];
    cnt[3] = CrossingIntervals(work, ylen - 1, afs, &loc[3], &itv[3]);
    // GetF0CandidateContour (:483-523) + the score normalisation of :562-567
    const bool usable = cnt[0] > 2 && cnt[1] > 2 && cnt[2] > 2 && cnt[3] > 2;
    if (usable) for (int q = 0; q < 4; ++q) Interp1Vec(loc[q], itv[q], t, L, &yi[q]);
    for (int i = 0; i < L; ++i) {
      double f = 0.0, sc = 100000.0;                                                   // kMaximumValue (constantnumbers.h)
      if (usable) {
        f = (yi[0][i] + yi[1][i] + yi[2][i] + yi[3][i]) / 4.0;
        sc = sqrt(((yi[0][i] - f) * (yi[0][i] - f) + (yi[1][i] - f) * (yi[1][i] - f) +
                   (yi[2][i] - f) * (yi[2][i] - f) + (yi[3][i] - f) * (yi[3][i] - f)) / 3.0);
        if (f > boundary[b] || f < boundary[b] / 2.0 || f > o->f0_ceil || f < o->f0_floor) { f = 0.0; sc = 100000.0; }
      }
      cand[b][i] = f;
      score[b][i] = sc / (f + kTiny);
    }
  }
  // GetBestF0Contour (:112-126)
  std::vector<double> best(L);
  for (int i = 0; i < L; ++i) {
    double sc = score[0][i];
    best[i] = cand[0][i];
    for (int b = 1; b < nb; ++b) if (sc > score[b][i]) { sc = score[b][i]; best[i] = cand[b][i]; }
  }
  // FixF0Contour (:264-289): when the contour is too short the reference leaves f0 untouched; zeros here
  const int vrm = (int)(0.5 + 1000.0 / o->frame_period / o->f0_floor) * 2 + 1;
  for (int i = 0; i < L; ++i) f0[i] = 0.0;
  if (L <= vrm) return;
  std::vector<double> base(L, 0.0), s1(L, 0.0), s2, s3, s4;
  for (int i = vrm; i < L - vrm; ++i) base[i] = best[i];                              // FixStep1 (:132-151)
  for (int i = vrm; i < L; ++i)
    s1[i] = fabs((base[i] - base[i - 1]) / (kTiny + base[i])) < o->allowed_range ? base[i] : 0.0;
  s2 = s1;                                                                            // FixStep2 (:157-171)
  const int center = (vrm - 1) / 2;
  for (int i = center; i < L - center; ++i)
    for (int j = -center; j <= center; ++j)
      if (s1[i + j] == 0) { s2[i] = 0.0; break; }
  std::vector<int> rise, fall;                                                        // GetNumberOfVoicedSections (:176-187)
  for (int i = 1; i < L; ++i) {
    if (s2[i] == 0 && s2[i - 1] != 0) fall.push_back(i - 1);
    else if (s2[i - 1] == 0 && s2[i] != 0) rise.push_back(i);
  }
  s3 = s2;                                                                            // FixStep3 (:222-238): extend forwards
  for (size_t q = 0; q < fall.size(); ++q) {
    const int limit = q + 1 == fall.size() ? L - 1 : fall[q + 1];
    for (int j = fall[q]; j < limit; ++j) {
      s3[j + 1] = DioSelect(s3[j], s3[j - 1], cand, j + 1, o->allowed_range);
      if (s3[j + 1] == 0) break;
    }
  }
  s4 = s3;                                                                            // FixStep4 (:244-260): extend backwards
  for (int q = (int)rise.size() - 1; q >= 0; --q) {
    const int limit = q == 0 ? 1 : rise[q - 1];
    for (int j = rise[q]; j > limit; --j) {
      s4[j - 1] = DioSelect(s4[j], s4[j + 1], cand, j - 1, o->allowed_range);
      if (s4[j - 1] == 0) break;
    }
  }
  for (int i = 0; i < L; ++i) f0[i] = s4[i];
}

// ---- Harvest (harvest.cpp).  Everything on the 1 ms grid, then subsampled (:1237-1251).
typedef struct { double f0_floor, f0_ceil, frame_period; } HarvestOption;            // harvest.h:16-20
void InitializeHarvestOption(HarvestOption *o) { o->f0_ceil = 800.0; o->f0_floor = 71.0; o->frame_period = 5; }  // harvest.cpp:1257-1263

namespace {
typedef std::vector<std::vector<double> > Rows;

// SelectBestF0 (:636-650): the LAST candidate among those with the smallest relative error <= allowed
double HvSelect(double ref, const std::vector<double> &cand, int n, double allowed, double *err) {
  double best = 0.0;
  *err = allowed;
  for (int i = 0; i < n; ++i) {
    const double e = fabs(ref - cand[i]) / ref;
    if (e > *err) continue;
    best = cand[i];
    *err = e;
  }
  return best;
}

// GetBoundaryList (:727-743): alternating rise / fall positions, falls shifted back by one
int HvBoundaries(const std::vector<double> &f0, std::vector<int> *list) {
  const int n = (int)f0.size();
  list->clear();
  int prev = 0;
  for (int i = 1; i < n; ++i) {
    const int v = (i == n - 1) ? 0 : (f0[i] > 0 ? 1 : 0);
    if (v != prev) list->push_back(i - (int)list->size() % 2);
    prev = v;
  }
  return (int)list->size();
}

// GetRefinedF0 (:589-617) with GetMeanF0 (:540-583) and FixF0 (:498-535): two windowed FFTs per candidate
void HvRefine(const std::vector<double> &y, double fs, double t, double f, double f0_floor, double f0_ceil,
              double *out_f, double *out_score) {
  *out_f = 0.0; *out_score = 0.0;
  if (f <= 0.0) return;
  const int n = (int)y.size();
  const int h = (int)(1.5 * fs / f + 1.0), len = 2 * h + 1;
  const double win_time = (2.0 * h + 1.0) / fs;
  const int fft_size = (int)pow(2.0, 2.0 + (int)(log(h * 2.0 + 1.0) / kLog2));
  const int basic = RoundHalfAway((t + (-h) / fs) * fs + 0.001);                       // GetBaseIndex (:434-441)
  std::vector<double> w(len), dw(len), a(fft_size, 0.0), b(fft_size, 0.0), ar, ai, br, bi;
  for (int i = 0; i < len; ++i) {                                                      // GetMainWindow (:446-456)
    const double tau = (basic + i - 1.0) / fs - t;
    w[i] = 0.42 + 0.5 * cos(2.0 * kPi * tau / win_time) + 0.08 * cos(4.0 * kPi * tau / win_time);
  }
  dw[0] = -w[1] / 2.0;                                                                  // GetDiffWindow (:462-468)
  for (int i = 1; i < len - 1; ++i) dw[i] = -(w[i + 1] - w[i - 1]) / 2.0;
  dw[len - 1] = w[len - 2] / 2.0;
  for (int i = 0; i < len; ++i) {                                                      // GetSpectra (:474-496)
    const double v = y[std::max(0, std::min(n - 1, basic + i - 1))];
    a[i] = v * w[i];
    b[i] = v * dw[i];
  }
  RealFFT(a, &ar, &ai);
  RealFFT(b, &br, &bi);
  const int H = std::min((int)(fs / 2.0 / f), 6);
  double num = 0.0, den = 0.0, score = 0.0;
  for (int m = 0; m < H; ++m) {                                                        // FixF0
    const int k = RoundHalfAway(f * fft_size / fs * (m + 1));
    const double pw = ar[k] * ar[k] + ai[k] * ai[k];
    const double ni = ar[k] * bi[k] - ai[k] * br[k];
    const double inst = pw == 0.0 ? 0.0 : (double)k * fs / fft_size + ni / pw * fs / 2.0 / kPi;
    const double amp = sqrt(pw);
    num += amp * inst;
    den += amp * (m + 1.0);
    score += fabs((inst / (m + 1.0) - f) / f);
  }
  double rf = num / (den + kTiny), rs = 1.0 / (score / H + kTiny);
  if (rf < f0_floor || rf > f0_ceil || rs < 2.5) { rf = 0.0; rs = 0.0; }
  *out_f = rf; *out_score = rs;
}

// ExtendF0 (:806-836): follow the candidates frame by frame from `origin` in direction `shift`
int HvExtendOne(int origin, int last, int shift, const Rows &cand, int n, double allowed, std::vector<double> *f0) {
  double cur = (*f0)[origin];
  int reached = origin, misses = 0;
  const int distance = abs(last - origin);
  for (int i = 0; i <= distance; ++i) {
    const int at = origin + shift * i + shift;
    double e;
    (*f0)[at] = HvSelect(cur, cand[at], n, allowed, &e);
    if ((*f0)[at] == 0.0) {
      ++misses;
    } else {
      cur = (*f0)[at];
      misses = 0;
      reached = at;
    }
    if (misses == 4) break;
  }
  return reached;
}

double HvScoreOf(double f0, const std::vector<double> &cand, const std::vector<double> &score, int n) {   // SearchScore (:905-911)
  double s = 0.0;
  for (int i = 0; i < n; ++i)
    if (f0 == cand[i] && s < score[i]) s = score[i];
  return s;
}

// FixStep3 (:967-996) = GetMultiChannelF0 + Extend (+ ExtendSub) + MergeF0, quirks included: the running mean of
// ExtendSub is never reset, MergeF0 starts from channel 0 (not the earliest) and reuses boundary slots 0 / 1
void HvStep3(const std::vector<double> &in, const Rows &cand, const Rows &score, int n, double allowed, std::vector<double> *out) {
  const int L = (int)in.size();
  *out = in;
  std::vector<int> bl;
  const int nsec = HvBoundaries(in, &bl) / 2;
  Rows ch(nsec, std::vector<double>(L, 0.0));
  for (int i = 0; i < nsec; ++i)
    for (int j = bl[2 * i]; j <= bl[2 * i + 1]; ++j) ch[i][j] = in[j];
  for (int i = 0; i < nsec; ++i) {                                                     // Extend (:874-889)
    bl[2 * i + 1] = HvExtendOne(bl[2 * i + 1], std::min(L - 2, bl[2 * i + 1] + 100), 1, cand, n, allowed, &ch[i]);
    bl[2 * i] = HvExtendOne(bl[2 * i], std::max(1, bl[2 * i] - 100), -1, cand, n, allowed, &ch[i]);
  }
  int kept = 0;                                                                        // ExtendSub (:853-869)
  double mean = 0.0;
  for (int i = 0; i < nsec; ++i) {
    const int st = bl[2 * i], ed = bl[2 * i + 1];
    for (int j = st; j < ed; ++j) mean += ch[i][j];
    mean /= ed - st;
    if (2200.0 / mean < ed - st) {
      std::swap(ch[kept], ch[i]);
      std::swap(bl[2 * kept], bl[2 * i]);
      std::swap(bl[2 * kept + 1], bl[2 * i + 1]);
      ++kept;
    }
  }
  if (kept == 0) return;
  std::vector<int> order(kept);                                                        // MakeSortedOrder (:891-903)
  for (int i = 0; i < kept; ++i) order[i] = i;
  for (int i = 1; i < kept; ++i)             // as written there: position i is the comparison partner throughout,
    for (int j = i - 1; j >= 0; --j) {       // so this is not a full insertion sort when an element must move far
      if (bl[order[j] * 2] > bl[order[i] * 2]) std::swap(order[i], order[j]);
      else break;
    }
  std::vector<double> &m = *out;                                                       // MergeF0 (:938-963)
  m = ch[0];
  for (int i = 1; i < kept; ++i) {
    const int st2 = bl[order[i] * 2], ed2 = bl[order[i] * 2 + 1];
    const std::vector<double> &f2 = ch[order[i]];
    if (st2 - bl[1] > 0) {
      for (int j = st2; j <= ed2; ++j) m[j] = f2[j];
      bl[0] = st2; bl[1] = ed2;
    } else {                                                                           // MergeF0Sub (:916-933)
      const int st1 = bl[0], ed1 = bl[1];
      if (st1 <= st2 && ed1 >= ed2) { bl[1] = ed1; continue; }
      double s1 = 0.0, s2 = 0.0;
      for (int j = st2; j <= ed1; ++j) {
        s1 += HvScoreOf(m[j], cand[j], score[j], n);
        s2 += HvScoreOf(f2[j], cand[j], score[j], n);
      }
      if (s1 > s2) for (int j = ed1; j <= ed2; ++j) m[j] = f2[j];
      else for (int j = st2; j <= ed2; ++j) m[j] = f2[j];
      bl[1] = ed2;
    }
  }
}

void HvBody(const double *x, int x_length, int fs, double f0_floor, double f0_ceil, int ratio, std::vector<double> *f0_out) {   // HarvestGeneralBody (:1145-1215)
  const double ch_per_oct = 40.0, afloor = f0_floor * 0.9, aceil = f0_ceil * 1.1;
  const int nch = 1 + (int)(log(aceil / afloor) / kLog2 * ch_per_oct);
  std::vector<double> boundary(nch);
  for (int i = 0; i < nch; ++i) boundary[i] = afloor * pow(2.0, (i + 1) / ch_per_oct);
  ratio = std::max(std::min(ratio, 12), 1);
  const int ylen = (int)ceil((double)x_length / ratio);
  const double afs = (double)fs / ratio;
  const int L = GetSamplesForHarvest(fs, x_length, 1.0);
  std::vector<double> t(L);
  for (int i = 0; i < L; ++i) t[i] = i * 1 / 1000.0;
  // GetWaveformAndSpectrum(Sub) (:43-93): decimation of an edge-padded copy, then DC removal
  std::vector<double> y(ylen, 0.0);
  if (ratio == 1) {
    for (int i = 0; i < x_length; ++i) y[i] = x[i];
  } else {
    const int lag = (int)(ceil(140.0 / ratio) * ratio);
    std::vector<double> padded(x_length + 2 * lag), dec;
    for (int i = 0; i < lag; ++i) padded[i] = x[0];
    for (int i = 0; i < x_length; ++i) padded[lag + i] = x[i];
    for (int i = 0; i < lag; ++i) padded[lag + x_length + i] = x[x_length - 1];
    Decimate(padded.data(), (int)padded.size(), ratio, &dec);
    for (int i = 0; i < ylen; ++i) y[i] = (lag / ratio + i < (int)dec.size()) ? dec[lag / ratio + i] : 0.0;
  }
  double mean = 0.0;
  for (int i = 0; i < ylen; ++i) mean += y[i];
  mean /= ylen;
  for (int i = 0; i < ylen; ++i) y[i] -= mean;

  // bins NF/2 - 1 and NF/2 of the signal's spectrum, NF = the reference's fft_size (:1164-1165)
  const int NF = (int)pow(2.0, (int)(log((double)(ylen + 5 + 2 * (int)(2.0 * afs / boundary[0]))) / kLog2) + 1.0);
  const int N2 = NF / 2;
  double ys1r = 0.0, ys1i = 0.0, ys2 = 0.0;
  for (int n = 0; n < ylen; ++n) {
    ys1r += y[n] * cos(2.0 * kPi * (N2 - 1) * n / NF); ys1i -= y[n] * sin(2.0 * kPi * (N2 - 1) * n / NF);
    ys2 += y[n] * ((n & 1) ? -1.0 : 1.0);
  }
  // raw candidates per channel (:99-147, :162-343): band-pass FIR, four crossing trains, interp1, gate
  Rows raw(nch, std::vector<double>(L, 0.0));
  std::vector<double> filt(ylen), work, loc[4], itv[4], yi[4];
  for (int c = 0; c < nch; ++c) {
    const int h = RoundHalfAway(afs / boundary[c] * 2.0), M = 2 * h + 1;
    std::vector<double> bp(M);
    for (int i = 0; i < M; ++i) {
      const double u = i / (M - 1.0);
      bp[i] = (0.355768 - 0.487396 * cos(2.0 * kPi * u) + 0.144232 * cos(4.0 * kPi * u) - 0.012604 * cos(6.0 * kPi * u)) *
              cos(2 * kPi * boundary[c] * (i - h) / afs);
    }
    // the mirroring loop of harvest.cpp:122-135 (same as DIO's, see Dio above): bins NF/2 - 1 and NF/2 of the
    // product both become Ys[NF/2] * (Ys[NF/2-1] F[NF/2-1]); the resulting ripple is ~1e-20 of the signal for
    // these long band-pass filters and matters only where the input is digitally silent
    double f1r = 0.0, f1i = 0.0, f2 = 0.0;
    for (int k = 0; k < M; ++k) {
      f1r += bp[k] * cos(2.0 * kPi * (N2 - 1) * k / NF); f1i -= bp[k] * sin(2.0 * kPi * (N2 - 1) * k / NF);
      f2 += bp[k] * ((k & 1) ? -1.0 : 1.0);
    }
    const double p_re = ys1r * f1r - ys1i * f1i, p_im = ys1r * f1i + ys1i * f1r;
    const double dq_re = ys2 * p_re - p_re, dq_im = ys2 * p_im - p_im, dn = ys2 * p_re - ys2 * f2;
    for (int i = 0; i < ylen; ++i) {                       // delay compensation h + 1 (:137-139)
      double acc = 0.0;
      const int k_lo = std::max(0, i + h + 1 - (ylen - 1)), k_hi = std::min(M - 1, i + h + 1);
      for (int k = k_lo; k <= k_hi; ++k) acc += bp[k] * y[i + h + 1 - k];
      const int m = i + h + 1;
      acc += (2.0 * (dq_re * cos(2.0 * kPi * (N2 - 1) * m / NF) - dq_im * sin(2.0 * kPi * (N2 - 1) * m / NF)) +
              dn * ((m & 1) ? -1.0 : 1.0)) / NF;
      filt[i] = acc;
    }
    int cnt[4];
    cnt[0] = CrossingIntervals(filt, ylen, afs, &loc[0], &itv[0]);
    work.assign(filt.begin(), filt.end());
    for (int i = 0; i < ylen; ++i) work[i] = -work[i];
    cnt[1] = CrossingIntervals(work, ylen, afs, &loc[1], &itv[1]);
    for (int i = 0; i + 1 < ylen; ++i) work[i] = work[i] - work[i + 1];
    cnt[2] = CrossingIntervals(work, ylen - 1, afs, &loc[2], &itv[2]);
    for (int i = 0; i + 1 < ylen; ++i) work[i] = -work[i];
    cnt[3] = CrossingIntervals(work, ylen - 1, afs, &loc[3], &itv[3]);
    if (!(cnt[0] > 2 && cnt[1] > 2 && cnt[2] > 2 && cnt[3] > 2)) continue;
    for (int q = 0; q < 4; ++q) Interp1Vec(loc[q], itv[q], t.data(), L, &yi[q]);
    for (int i = 0; i < L; ++i) {
      const double f = (yi[0][i] + yi[1][i] + yi[2][i] + yi[3][i]) / 4.0;
      raw[c][i] = (f > boundary[c] * 1.1 || f < boundary[c] * 0.9 || f > f0_ceil || f < f0_floor) ? 0.0 : f;
    }
  }
  // DetectOfficialF0Candidates (:348-412): runs of >= 10 voiced channels -> their mean
  const int max_cand = RoundHalfAway(nch / 10.0) * 7;
  Rows cand(L, std::vector<double>(max_cand, 0.0)), score(L, std::vector<double>(max_cand, 0.0));
  int nc = 0;
  for (int i = 0; i < L; ++i) {
    int count = 0, st = 0, prev = 0;
    for (int c = 1; c < nch; ++c) {
      const int v = (c == nch - 1) ? 0 : (raw[c][i] > 0 ? 1 : 0);
      if (v - prev == 1) st = c;
      if (v - prev == -1 && c - st >= 10) {
        double sum = 0.0;
        for (int j = st; j < c; ++j) sum += raw[j][i];
        cand[i][count++] = sum / (c - st);
      }
      prev = v;
    }
    nc = std::max(nc, count);
  }
  for (int d = 1; d <= 3; ++d)                                                         // OverlapF0Candidates (:417-429)
    for (int j = 0; j < nc; ++j) {
      for (int k = d; k < L; ++k) cand[k][j + nc * d] = cand[k - d][j];
      for (int k = 0; k < L - d; ++k) cand[k][j + nc * (d + 3)] = cand[k + d][j];
    }
  const int n = nc * 7;
  for (int i = 0; i < L; ++i)                                                          // RefineF0Candidates (:622-631)
    for (int j = 0; j < n; ++j) HvRefine(y, afs, t[i], cand[i][j], f0_floor, f0_ceil, &cand[i][j], &score[i][j]);
  {                                                                                    // RemoveUnreliableCandidates (:652-688)
    const Rows before(cand);
    for (int i = 1; i < L - 1; ++i)
      for (int j = 0; j < n; ++j) {
        if (cand[i][j] == 0) continue;
        double e1, e2;
        HvSelect(cand[i][j], before[i + 1], n, 1.0, &e1);
        HvSelect(cand[i][j], before[i - 1], n, 1.0, &e2);
        if (std::min(e1, e2) > 0.05) { cand[i][j] = 0; score[i][j] = 0; }
      }
  }
  // FixF0Contour (:1035-1053)
  std::vector<double> base(L, 0.0), s1(L, 0.0), s2, s3, s4;
  for (int i = 0; i < L; ++i) {                                                        // SearchF0Base (:693-705)
    double best = 0.0;
    for (int j = 0; j < n; ++j) if (score[i][j] > best) { base[i] = cand[i][j]; best = score[i][j]; }
  }
  for (int i = 2; i < L; ++i) {                                                        // FixStep1 (:710-722), allowed 0.008
    if (base[i] == 0.0) continue;
    const double ref = base[i - 1] * 2 - base[i - 2];
    s1[i] = (fabs((base[i] - ref) / ref) > 0.008 && fabs((base[i] - base[i - 1])) / base[i - 1] > 0.008) ? 0.0 : base[i];
  }
  s2 = s1;                                                                             // FixStep2 (:748-762), minimum 6
  std::vector<int> bl;
  int nb = HvBoundaries(s1, &bl);
  for (int i = 0; i < nb / 2; ++i) {
    if (bl[2 * i + 1] - bl[2 * i] >= 6) continue;
    for (int j = bl[2 * i]; j <= bl[2 * i + 1]; ++j) s2[j] = 0.0;
  }
  HvStep3(s2, cand, score, n, 0.18, &s3);
  s4 = s3;                                                                             // FixStep4 (:1001-1030), threshold 9
  nb = HvBoundaries(s3, &bl);
  for (int i = 0; i < nb / 2 - 1; ++i) {
    const int gap = bl[(i + 1) * 2] - bl[2 * i + 1] - 1;
    if (gap >= 9) continue;
    const double a = s3[bl[2 * i + 1]] + 1, b = s3[bl[(i + 1) * 2]] - 1;
    const double slope = (b - a) / (gap + 1.0);
    int count = 1;
    for (int j = bl[2 * i + 1] + 1; j <= bl[(i + 1) * 2] - 1; ++j) s4[j] = a + slope * count++;
  }
  // SmoothF0Contour (:1093-1129): zero-lag 2nd-order Butterworth per voiced section of the padded contour
  const double fb[2] = {0.0078202080334971724, 0.015640416066994345}, fa[2] = {1.7347257688092754, -0.76600660094326412};
  const int lag = 300, N = L + 2 * lag;
  std::vector<double> padded(N, 0.0), xs(N), rev(N), sm(N);
  for (int i = 0; i < L; ++i) padded[lag + i] = s4[i];
  f0_out->assign(L, 0.0);
  nb = HvBoundaries(padded, &bl);
  for (int i = 0; i < nb / 2; ++i) {
    const int st = bl[2 * i], ed = bl[2 * i + 1];
    for (int j = 0; j < N; ++j) xs[j] = padded[std::max(st, std::min(ed, j))];        // FilteringF0 (:1058-1086)
    double w0 = 0.0, w1 = 0.0;
    for (int j = 0; j < N; ++j) {
      const double wt = xs[j] + fa[0] * w0 + fa[1] * w1;
      rev[N - j - 1] = fb[0] * wt + fb[1] * w0 + fb[0] * w1;
      w1 = w0; w0 = wt;
    }
    w0 = w1 = 0.0;
    for (int j = 0; j < N; ++j) {
      const double wt = rev[j] + fa[0] * w0 + fa[1] * w1;
      sm[N - j - 1] = fb[0] * wt + fb[1] * w0 + fb[0] * w1;
      w1 = w0; w0 = wt;
    }
    for (int j = st; j <= ed; ++j) (*f0_out)[j - lag] = sm[j];
  }
}
}  // namespace

void Harvest(const double *x, int x_length, int fs, const HarvestOption *o, double *t, double *f0) {   // :1223-1255
  std::vector<double> basic;
  HvBody(x, x_length, fs, o->f0_floor, o->f0_ceil, RoundHalfAway(fs / 8000.0), &basic);
  const int L = GetSamplesForHarvest(fs, x_length, o->frame_period);
  for (int i = 0; i < L; ++i) {
    t[i] = i * o->frame_period / 1000.0;
    f0[i] = basic[std::min((int)basic.size() - 1, RoundHalfAway(t[i] * 1000.0))];
  }
}

// ---- Synthesis (synthesis.cpp:339-399).  Transform conventions of fft.cpp written out as plain sums over a
// textbook complex FFT: c2c FORWARD(a) = FFT(conj a) (:61-71), c2r(X) = X0.re + (-1)^n X_{N/2}.re +
// 2 sum_k Re(X_k e^{+j 2 pi k n / N}) (:26-35).
namespace {
// in-place complex FFT, sum a[n] exp(-j 2 pi k n / N)
void ComplexFFT(std::vector<double> *re, std::vector<double> *im) {
  std::vector<double> &ar = *re, &ai = *im;
  const int n = (int)ar.size();
  for (int i = 1, j = 0; i < n; ++i) {
    int bit = n >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) { std::swap(ar[i], ar[j]); std::swap(ai[i], ai[j]); }
  }
  for (int len = 2; len <= n; len <<= 1)
    for (int i = 0; i < n; i += len)
      for (int k = 0; k < len / 2; ++k) {
        const double wr = cos(-2.0 * kPi * k / len), wi = sin(-2.0 * kPi * k / len);
        const int a = i + k, b = i + k + len / 2;
        const double tr = ar[b] * wr - ai[b] * wi, ti = ar[b] * wi + ai[b] * wr;
        ar[b] = ar[a] - tr; ai[b] = ai[a] - ti;
        ar[a] += tr; ai[a] += ti;
      }
}

// c2r of a half spectrum (N/2+1 bins) -> N real samples, unnormalised
void HalfToReal(const std::vector<double> &xr, const std::vector<double> &xi, int N, std::vector<double> *out) {
  std::vector<double> re(N), im(N);
  for (int k = 0; k <= N / 2; ++k) { re[k] = xr[k]; im[k] = -xi[k]; }     // conj: turns e^{+j} into the forward kernel
  im[0] = 0.0; im[N / 2] = 0.0;
  for (int k = N / 2 + 1; k < N; ++k) { re[k] = xr[N - k]; im[k] = xi[N - k]; }
  ComplexFFT(&re, &im);
  *out = re;
}

// GetMinimumPhaseSpectrum (common.cpp:192-226): log spectrum (N/2+1) -> minimum-phase spectrum (N/2+1)
void MinimumPhase(const std::vector<double> &log_spec, int N, std::vector<double> *mr, std::vector<double> *mi) {
  std::vector<double> full(N), cr, ci;
  for (int i = 0; i <= N / 2; ++i) full[i] = log_spec[i];
  for (int i = N / 2 + 1; i < N; ++i) full[i] = full[N - i];
  RealFFT(full, &cr, &ci);
  std::vector<double> re(N, 0.0), im(N, 0.0);
  re[0] = cr[0]; im[0] = -ci[0];
  for (int i = 1; i < N / 2; ++i) { re[i] = cr[i] * 2.0; im[i] = ci[i] * -2.0; }
  re[N / 2] = cr[N / 2]; im[N / 2] = -ci[N / 2];
  for (int i = 0; i < N; ++i) im[i] = -im[i];                               // FORWARD = FFT(conj a)
  ComplexFFT(&re, &im);
  mr->resize(N / 2 + 1); mi->resize(N / 2 + 1);
  for (int i = 0; i <= N / 2; ++i) {
    const double e = exp(re[i] / N);
    (*mr)[i] = e * cos(im[i] / N);
    (*mi)[i] = e * sin(im[i] / N);
  }
}

double SafeAp(double a) { return std::max(0.001, std::min(0.999999999999, a)); }     // common.h GetSafeAperiodicity
}  // namespace

void Synthesis(const double *f0, int f0_length, const double *const *sp, const double *const *ap, int fft_size,
               double frame_period_ms, int fs, int y_length, double *y) {
  const int N = fft_size, half = N / 2;
  const double fp = frame_period_ms / 1000.0;
  const double lowest_f0 = fs / fft_size + 1.0;                                        // integer division, as there (:362)
  for (int i = 0; i < y_length; ++i) y[i] = 0.0;
  // GetTimeBase (:283-315): interpolate f0 / vuv to the sample grid, accumulate phase, find the wraps
  std::vector<double> ct(f0_length + 1), cf(f0_length + 1), cv(f0_length + 1), vuv(y_length), wrap(y_length);
  for (int i = 0; i < f0_length; ++i) {
    ct[i] = i * fp;
    cf[i] = f0[i] < lowest_f0 ? 0.0 : f0[i];
    cv[i] = cf[i] == 0.0 ? 0.0 : 1.0;
  }
  ct[f0_length] = f0_length * fp;
  cf[f0_length] = cf[f0_length - 1] * 2 - cf[f0_length - 2];
  cv[f0_length] = cv[f0_length - 1] * 2 - cv[f0_length - 2];
  double total = 0.0;
  for (int i = 0; i < y_length; ++i) {
    const double t = i / (double)fs;
    vuv[i] = Interp1At(ct, cv, t) > 0.5 ? 1.0 : 0.0;
    const double fi = vuv[i] == 0.0 ? 500.0 : Interp1At(ct, cf, t);                     // kDefaultF0
    total = i == 0 ? 2.0 * kPi * fi / fs : total + 2.0 * kPi * fi / fs;
    wrap[i] = fmod(total, 2.0 * kPi);
  }
  std::vector<int> pidx;
  std::vector<double> pshift;
  for (int i = 0; i + 1 < y_length; ++i)
    if (fabs(wrap[i + 1] - wrap[i]) > kPi) {
      const double y1 = wrap[i] - 2.0 * kPi, y2 = wrap[i + 1];
      pidx.push_back(i);
      pshift.push_back(-y1 / (y2 - y1) / fs);
    }
  std::vector<double> dcr(N);                                                           // GetDCRemover (:319-333)
  {
    double dc = 0.0;
    for (int i = 0; i < half; ++i) {
      dcr[i] = 0.5 - 0.5 * cos(2.0 * kPi * (i + 1.0) / (1.0 + N));
      dcr[N - i - 1] = dcr[i];
      dc += dcr[i] * 2.0;
    }
    for (int i = 0; i < half; ++i) { dcr[i] /= dc; dcr[N - i - 1] = dcr[i]; }
  }
  Rng rng;
  const int np = (int)pidx.size();
  std::vector<double> env(half + 1), ratio(half + 1), lg(half + 1), mr, mi, xr(half + 1), xi(half + 1), wave, per(N), aper(N), noise(N), nr, ni;
  for (int p = 0; p < np; ++p) {
    const int noise_size = pidx[std::min(np - 1, p + 1)] - pidx[p];
    const double cur_vuv = vuv[pidx[p]], cur_t = pidx[p] / (double)fs;
    // GetSpectralEnvelope / GetAperiodicRatio (:140-179)
    const int fl = std::min(f0_length - 1, (int)floor(cur_t / fp)), ce = std::min(f0_length - 1, (int)ceil(cur_t / fp));
    const double w = cur_t / fp - fl;
    for (int k = 0; k <= half; ++k) {
      if (fl == ce) {
        env[k] = fabs(sp[fl][k]);
        ratio[k] = pow(SafeAp(ap[fl][k]), 2.0);
      } else {
        env[k] = (1.0 - w) * fabs(sp[fl][k]) + w * fabs(sp[ce][k]);
        ratio[k] = pow((1.0 - w) * SafeAp(ap[fl][k]) + w * SafeAp(ap[ce][k]), 2.0);
      }
    }
    // GetPeriodicResponse (:110-138)
    if (cur_vuv <= 0.5 || ratio[0] > 0.999) {
      std::fill(per.begin(), per.end(), 0.0);
    } else {
      for (int k = 0; k <= half; ++k) lg[k] = log(env[k] * (1.0 - ratio[k]) + kTiny) / 2.0;
      MinimumPhase(lg, N, &mr, &mi);
      const double coef = 2.0 * kPi * pshift[p] * fs / N;
      for (int k = 0; k <= half; ++k) {                                                 // fractional delay (:95-106)
        const double c = cos(coef * k), sn = sqrt(1.0 - c * c);
        xr[k] = mr[k] * c + mi[k] * sn;
        xi[k] = mi[k] * c - mr[k] * sn;
      }
      HalfToReal(xr, xi, N, &wave);
      for (int i = 0; i < half; ++i) { per[i] = wave[i + half]; per[i + half] = wave[i]; }   // fftshift
      double dc = 0.0;                                                                  // RemoveDCComponent (:73-82)
      for (int i = half; i < N; ++i) dc += per[i];
      for (int i = 0; i < half; ++i) per[i] = -dc * dcr[i];
      for (int i = half; i < N; ++i) per[i] -= dc * dcr[i];
    }
    // GetAperiodicResponse (:36-69) with GetNoiseSpectrum (:18-31)
    double avg = 0.0;
    std::fill(noise.begin(), noise.end(), 0.0);
    for (int i = 0; i < noise_size; ++i) { noise[i] = rng.Next(); avg += noise[i]; }
    avg /= noise_size;
    for (int i = 0; i < noise_size; ++i) noise[i] -= avg;
    RealFFT(noise, &nr, &ni);
    for (int k = 0; k <= half; ++k) lg[k] = (cur_vuv != 0.0 ? log(env[k] * ratio[k]) : log(env[k])) / 2.0;
    MinimumPhase(lg, N, &mr, &mi);
    for (int k = 0; k <= half; ++k) {
      xr[k] = mr[k] * nr[k] - mi[k] * ni[k];
      xi[k] = mr[k] * ni[k] + mi[k] * nr[k];
    }
    HalfToReal(xr, xi, N, &wave);
    for (int i = 0; i < half; ++i) { aper[i] = wave[i + half]; aper[i + half] = wave[i]; }
    // GetOneFrameSegment's mix (:203-207) and the overlap-add of :383-390
    const double sq = sqrt((double)noise_size);
    const int offset = pidx[p] - half + 1;
    for (int j = std::max(0, -offset); j < std::min(N, y_length - offset); ++j)
      y[j + offset] += (per[j] * sq + aper[j]) / N;
  }
}

}  // extern "C"

// test hook: the restated decimate() on its own (matlabfunctions.cpp:178-204)
extern "C" int OracleDecimate(const double *x, int n, int r, double *y) {
  std::vector<double> out;
  Decimate(x, n, r, &out);
  for (size_t i = 0; i < out.size(); ++i) y[i] = out[i];
  return (int)out.size();
}
