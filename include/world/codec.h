/* world/codec.h -- coder / decoder for the spectral envelope (mel-cepstrum by DCT) and the
 * aperiodicity (band values in dB), legacy entry points (reference: src/world/codec.h:20-92,
 * SURVEY.md 8 row f2).  Each call runs the batched kernels of world_b200.h with n_utts = 1. */
#ifndef WORLD_CODEC_H_
#define WORLD_CODEC_H_
#include "world/macrodefinitions.h"
WORLD_BEGIN_C_DECLS

/* int(min(15000, fs / 2 - 3000) / 3000): coded aperiodicity values per frame (codec.cpp:216-219). */
WORLD_API int GetNumberOfAperiodicities(int fs);

/* aperiodicity: f0_length rows of fft_size/2+1 -> coded_aperiodicity: f0_length rows of
 * GetNumberOfAperiodicities(fs) (dB at 3 kHz, 6 kHz, ...). */
WORLD_API void CodeAperiodicity(const double *const *aperiodicity, int f0_length, int fs, int fft_size,
                                double **coded_aperiodicity);
WORLD_API void DecodeAperiodicity(const double *const *coded_aperiodicity, int f0_length, int fs,
                                  int fft_size, double **aperiodicity);

/* spectrogram: f0_length rows of fft_size/2+1 -> coded_spectral_envelope: f0_length rows of
 * number_of_dimensions mel-cepstral coefficients (number_of_dimensions <= fft_size/4 + 1). */
WORLD_API void CodeSpectralEnvelope(const double *const *spectrogram, int f0_length, int fs, int fft_size,
                                    int number_of_dimensions, double **coded_spectral_envelope);
WORLD_API void DecodeSpectralEnvelope(const double *const *coded_spectral_envelope, int f0_length, int fs,
                                      int fft_size, int number_of_dimensions, double **spectrogram);

WORLD_END_C_DECLS
#endif /* WORLD_CODEC_H_ */
