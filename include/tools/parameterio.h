/* tools/parameterio.h -- F0 / spectral envelope / aperiodicity parameter files in the reference's binary
 * layout (reference: tools/parameterio.h:17-120, tools/parameterio.cpp:59-243), so that
 * the reference's examples (examples/analysis_synthesis) read what this library computed and vice versa.
 *   F0  file: "F0  " "NOF " int32 "FP  " float64, then NOF float64 (text_flag = 1: "%.5f %.5f\r\n" lines)
 *   SPEC / AP: tag, "NOF " int32, "FP  " float64, "FFT " int32, "NOD " int32, "FS  " int32, then NOF rows of
 *              (NOD == 0 ? FFT / 2 + 1 : NOD) float64
 * Host-only file glue. */
#ifndef WORLD_PARAMETERIO_H_
#define WORLD_PARAMETERIO_H_
#ifdef __cplusplus
extern "C" {
#endif

void WriteF0(const char *filename, int f0_length, double frame_period, const double *temporal_positions,
             const double *f0, int text_flag);
int ReadF0(const char *filename, double *temporal_positions, double *f0);
/* parameter: "NOF ", "FP  ", "FFT ", "NOD " or "FS  " */
double GetHeaderInformation(const char *filename, const char *parameter);
void WriteSpectralEnvelope(const char *filename, int fs, int f0_length, double frame_period, int fft_size,
                           int number_of_dimensions, const double *const *spectrogram);
int ReadSpectralEnvelope(const char *filename, double **spectrogram);
void WriteAperiodicity(const char *filename, int fs, int f0_length, double frame_period, int fft_size,
                       int number_of_dimensions, const double *const *aperiodicity);
int ReadAperiodicity(const char *filename, double **aperiodicity);

/* The same files straight from / into the flat row-major arrays of the batched ABI (row u of a batch starts
 * at rows + u * f0_stride * width): kind is "SPEC" or "AP  "; returns 0 or WORLD_B200_EINVAL. */
int world_b200_write_rows(const char *filename, const char *kind, int fs, int f0_length, double frame_period,
                          int fft_size, int number_of_dimensions, const double *rows);
int world_b200_read_rows(const char *filename, const char *kind, double *rows, int max_frames);

#ifdef __cplusplus
}
#endif
#endif /* WORLD_PARAMETERIO_H_ */
