// Exercises include/world_b200.hpp: the batched C++ overloads must give, for every utterance, exactly
// what the reference-compatible single-utterance entry points give (same kernels, N = 1 vs N = 3).
#include <cmath>
#include <cstdio>
#include <vector>

#include "world_b200.hpp"

static std::vector<double> tone(int n, int fs, double f0, unsigned seed) {
  std::vector<double> x(n);
  unsigned s = seed;
  for (int i = 0; i < n; ++i) {
    s = s * 1664525u + 1013904223u;
    const double noise = ((s >> 8) / 16777216.0 - 0.5) * 0.01;
    double v = 0.0;
    for (int k = 1; k <= 8; ++k) v += std::sin(2.0 * 3.14159265358979323846 * f0 * k * i / fs) / k;
    x[i] = 0.2 * v + noise;
  }
  return x;
}

int main() {
  const int fs = 16000, n_utts = 3;
  const int lens[3] = {8000, 6400, 4000};
  const double f0s_true[3] = {120.0, 180.0, 240.0};
  std::vector<std::vector<double>> x(n_utts);
  const double *xs[3];
  for (int u = 0; u < n_utts; ++u) { x[u] = tone(lens[u], fs, f0s_true[u], 17u + u); xs[u] = x[u].data(); }

  DioOption dopt; InitializeDioOption(&dopt);
  CheapTrickOption copt; InitializeCheapTrickOption(fs, &copt);
  D4COption aopt; InitializeD4COption(&aopt);
  const int bins = copt.fft_size / 2 + 1;
  int fl[3];
  std::vector<std::vector<double>> t(n_utts), f0(n_utts), f0r(n_utts), y(n_utts);
  std::vector<std::vector<double>> sp(n_utts), ap(n_utts);
  std::vector<std::vector<double *>> sp_rows(n_utts), ap_rows(n_utts);
  double *tp[3], *fp[3], *frp[3], *yp[3];
  double **spp[3], **app[3];
  for (int u = 0; u < n_utts; ++u) {
    fl[u] = GetSamplesForDIO(fs, lens[u], dopt.frame_period);
    t[u].resize(fl[u]); f0[u].resize(fl[u]); f0r[u].resize(fl[u]); y[u].resize(lens[u]);
    sp[u].resize((size_t)fl[u] * bins); ap[u].resize((size_t)fl[u] * bins);
    sp_rows[u].resize(fl[u]); ap_rows[u].resize(fl[u]);
    for (int i = 0; i < fl[u]; ++i) { sp_rows[u][i] = &sp[u][(size_t)i * bins]; ap_rows[u][i] = &ap[u][(size_t)i * bins]; }
    tp[u] = t[u].data(); fp[u] = f0[u].data(); frp[u] = f0r[u].data(); yp[u] = y[u].data();
    spp[u] = sp_rows[u].data(); app[u] = ap_rows[u].data();
  }
  int rc = Dio(xs, lens, n_utts, fs, &dopt, tp, fp);
  if (!rc) rc = StoneMask(xs, lens, n_utts, fs, tp, fp, fl, frp);
  if (!rc) rc = CheapTrick(xs, lens, n_utts, fs, tp, frp, fl, &copt, spp);
  if (!rc) rc = D4C(xs, lens, n_utts, fs, tp, frp, fl, copt.fft_size, &aopt, app);
  if (!rc) rc = Synthesis(frp, fl, n_utts, spp, app, copt.fft_size, dopt.frame_period, fs, lens, yp);
  if (rc) { std::printf("FAIL: batched overload returned %d\n", rc); return 1; }

  // the same through the reference's own single-utterance API
  int bad = 0, voiced = 0;
  for (int u = 0; u < n_utts; ++u) {
    std::vector<double> t1(fl[u]), f1(fl[u]), r1(fl[u]), y1(lens[u]);
    std::vector<double> s1((size_t)fl[u] * bins), a1((size_t)fl[u] * bins);
    std::vector<double *> s1r(fl[u]), a1r(fl[u]);
    for (int i = 0; i < fl[u]; ++i) { s1r[i] = &s1[(size_t)i * bins]; a1r[i] = &a1[(size_t)i * bins]; }
    Dio(xs[u], lens[u], fs, &dopt, t1.data(), f1.data());
    StoneMask(xs[u], lens[u], fs, t1.data(), f1.data(), fl[u], r1.data());
    CheapTrick(xs[u], lens[u], fs, t1.data(), r1.data(), fl[u], &copt, s1r.data());
    D4C(xs[u], lens[u], fs, t1.data(), r1.data(), fl[u], copt.fft_size, &aopt, a1r.data());
    Synthesis(r1.data(), fl[u], s1r.data(), a1r.data(), copt.fft_size, dopt.frame_period, fs, lens[u], y1.data());
    for (int i = 0; i < fl[u]; ++i) {
      bad += (t1[i] != t[u][i]) + (f1[i] != f0[u][i]) + (r1[i] != f0r[u][i]);
      voiced += r1[i] > 0;
    }
    for (size_t i = 0; i < s1.size(); ++i) bad += (s1[i] != sp[u][i]) + (a1[i] != ap[u][i]);
    for (int i = 0; i < lens[u]; ++i) bad += (y1[i] != y[u][i]);
  }
  // Harvest: batched overload against the single-utterance entry point
  {
    HarvestOption hopt; InitializeHarvestOption(&hopt);
    std::vector<std::vector<double>> th(n_utts), fh(n_utts);
    double *thp[3], *fhp[3];
    for (int u = 0; u < n_utts; ++u) { th[u].resize(fl[u]); fh[u].resize(fl[u]); thp[u] = th[u].data(); fhp[u] = fh[u].data(); }
    rc = Harvest(xs, lens, n_utts, fs, &hopt, thp, fhp);
    if (rc) { std::printf("FAIL: batched Harvest returned %d\n", rc); return 1; }
    for (int u = 0; u < n_utts; ++u) {
      std::vector<double> t1(fl[u]), f1(fl[u]);
      Harvest(xs[u], lens[u], fs, &hopt, t1.data(), f1.data());
      for (int i = 0; i < fl[u]; ++i) bad += (t1[i] != th[u][i]) + (f1[i] != fh[u][i]);
    }
  }
  if (bad || voiced == 0) { std::printf("FAIL: %d mismatching values, %d voiced frames\n", bad, voiced); return 1; }
  std::printf("OK: batched overloads == single-utterance API on %d utterances (%d voiced frames)\n", n_utts, voiced);
  return 0;
}
