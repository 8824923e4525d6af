// wb_fileio.cu -- host-only file glue around the analysis path: the reference's WAV reader / writer
// (tools/audioio.cpp) and parameter files (tools/parameterio.cpp), same function names, same bytes on disk
// (SURVEY.md 8 rows f3 / f4).  Nothing here touches the device; batches of files go through
// world_b200_wav_parse() + world_b200_pcm_to_double_batch() (wb_codec.cu).
#include "../../include/world_b200.h"
#include "../../include/tools/audioio.h"
#include "../../include/tools/parameterio.h"
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <vector>

namespace {

bool slurp(const char *filename, std::vector<unsigned char> *out) {
  FILE *fp = fopen(filename, "rb");
  if (!fp) return false;
  fseek(fp, 0, SEEK_END);
  const long n = ftell(fp);
  fseek(fp, 0, SEEK_SET);
  out->resize(n > 0 ? (size_t)n : 0);
  const size_t got = out->empty() ? 0 : fread(out->data(), 1, out->size(), fp);
  fclose(fp);
  return got == out->size();
}

void put_tag(FILE *fp, const char *tag) { fwrite(tag, 1, 4, fp); }
void put_i32(FILE *fp, const char *tag, int v) { put_tag(fp, tag); fwrite(&v, 4, 1, fp); }
void put_f64(FILE *fp, const char *tag, double v) { put_tag(fp, tag); fwrite(&v, 8, 1, fp); }

// header of a SPEC / AP file (parameterio.cpp:20-39, 148-166): returns the row width, 0 on mismatch
struct RowHeader { int frames, fft_size, width; };
bool read_row_header(FILE *fp, const char *kind, RowHeader *h) {
  char tag[4];
  int nod = 0;
  if (fread(tag, 1, 4, fp) != 4 || memcmp(tag, kind, 4) != 0) return false;
  if (fread(tag, 1, 4, fp) != 4 || fread(&h->frames, 4, 1, fp) != 1) return false;   // NOF
  if (fseek(fp, 12, SEEK_CUR) != 0) return false;                                     // FP + double
  if (fread(tag, 1, 4, fp) != 4 || fread(&h->fft_size, 4, 1, fp) != 1) return false;  // FFT
  if (fread(tag, 1, 4, fp) != 4 || fread(&nod, 4, 1, fp) != 1) return false;          // NOD
  if (fseek(fp, 8, SEEK_CUR) != 0) return false;                                      // FS + int
  h->width = nod == 0 ? h->fft_size / 2 + 1 : nod;
  return h->frames >= 0 && h->width > 0;
}

FILE *open_rows_for_write(const char *filename, const char *kind, int fs, int f0_length, double frame_period,
                          int fft_size, int nod) {
  FILE *fp = fopen(filename, "wb");
  if (!fp) { printf("File cannot be opened.\n"); return nullptr; }
  put_tag(fp, kind);
  put_i32(fp, "NOF ", f0_length);
  put_f64(fp, "FP  ", frame_period);
  put_i32(fp, "FFT ", fft_size);
  put_i32(fp, "NOD ", nod);
  put_i32(fp, "FS  ", fs);
  return fp;
}

void write_row_pointers(const char *filename, const char *kind, int fs, int f0_length, double frame_period,
                        int fft_size, int nod, const double *const *rows) {
  FILE *fp = open_rows_for_write(filename, kind, fs, f0_length, frame_period, fft_size, nod);
  if (!fp) return;
  const int width = nod == 0 ? fft_size / 2 + 1 : nod;
  for (int i = 0; i < f0_length; ++i) fwrite(rows[i], 8, width, fp);
  fclose(fp);
}

int read_row_pointers(const char *filename, const char *kind, double **rows) {
  FILE *fp = fopen(filename, "rb");
  if (!fp) { printf("File cannot be opened.\n"); return 0; }
  RowHeader h;
  if (!read_row_header(fp, kind, &h)) { printf("Header error.\n"); fclose(fp); return 0; }
  for (int i = 0; i < h.frames; ++i)
    if (fread(rows[i], 8, h.width, fp) != (size_t)h.width) break;
  fclose(fp);
  return 1;
}

}  // namespace

extern "C" {

// ---------------------------------------------------------------- tools/audioio.cpp
void wavwrite(const double *x, int x_length, int fs, int /*nbit*/, const char *filename) {   // :110-170
  FILE *fp = fopen(filename, "wb");
  if (!fp) { printf("File cannot be opened.\n"); return; }
  const uint32_t data_bytes = (uint32_t)x_length * 2u, riff = 36u + data_bytes, fmt_len = 16u;
  const uint32_t rate = (uint32_t)fs, byte_rate = (uint32_t)fs * 2u;
  const int16_t pcm = 1, channels = 1, block = 2, bits = 16;
  fwrite("RIFF", 1, 4, fp); fwrite(&riff, 4, 1, fp);
  fwrite("WAVE", 1, 4, fp); fwrite("fmt ", 1, 4, fp); fwrite(&fmt_len, 4, 1, fp);
  fwrite(&pcm, 2, 1, fp); fwrite(&channels, 2, 1, fp); fwrite(&rate, 4, 1, fp); fwrite(&byte_rate, 4, 1, fp);
  fwrite(&block, 2, 1, fp); fwrite(&bits, 2, 1, fp);
  fwrite("data", 1, 4, fp); fwrite(&data_bytes, 4, 1, fp);
  std::vector<int16_t> q((size_t)(x_length > 0 ? x_length : 0));
  for (int i = 0; i < x_length; ++i) {
    int v = static_cast<int>(x[i] * 32767);     // truncation toward zero, then clamp (:163-165)
    v = v < -32768 ? -32768 : (v > 32767 ? 32767 : v);
    q[i] = static_cast<int16_t>(v);
  }
  if (!q.empty()) fwrite(q.data(), 2, q.size(), fp);
  fclose(fp);
}

int GetAudioLength(const char *filename) {                                                   // :172-215
  std::vector<unsigned char> img;
  FILE *probe = fopen(filename, "rb");
  if (!probe) return 0;
  fclose(probe);
  if (!slurp(filename, &img)) return 0;
  int fs, nbit, n;
  unsigned long long off;
  if (world_b200_wav_parse(img.data(), img.size(), &fs, &nbit, &n, &off) != 0) return -1;
  return n;
}

void wavread(const char *filename, int *fs, int *nbit, double *x) {                          // :217-252
  std::vector<unsigned char> img;
  if (!slurp(filename, &img)) { printf("File not found.\n"); return; }
  int n;
  unsigned long long off;
  if (world_b200_wav_parse(img.data(), img.size(), fs, nbit, &n, &off) != 0) return;
  const int nb = *nbit / 8;
  const double zero_line = pow(2.0, *nbit - 1);
  for (int i = 0; i < n; ++i) {
    const unsigned char *s = img.data() + off + (size_t)i * nb;
    int64_t v = 0;
    for (int j = nb - 1; j >= 0; --j) v = v * 256 + s[j];
    if (s[nb - 1] >= 128) v -= (int64_t)1 << *nbit;    // two's complement: what :238-249 computes
    x[i] = static_cast<double>(v) / zero_line;
  }
}

// ---------------------------------------------------------------- tools/parameterio.cpp
void WriteF0(const char *filename, int f0_length, double frame_period, const double *temporal_positions,
             const double *f0, int text_flag) {                                              // :59-87
  FILE *fp = fopen(filename, text_flag == 1 ? "w" : "wb");
  if (!fp) { printf("File cannot be opened.\n"); return; }
  if (text_flag == 1) {
    for (int i = 0; i < f0_length; ++i) fprintf(fp, "%.5f %.5f\r\n", temporal_positions[i], f0[i]);
  } else {
    put_tag(fp, "F0  ");
    put_i32(fp, "NOF ", f0_length);
    put_f64(fp, "FP  ", frame_period);
    fwrite(f0, 8, f0_length, fp);
  }
  fclose(fp);
}

int ReadF0(const char *filename, double *temporal_positions, double *f0) {                   // :89-117
  FILE *fp = fopen(filename, "rb");
  if (!fp) { printf("File cannot be opened.\n"); return 0; }
  char tag[4];
  int frames = 0;
  double frame_period = 0.0;
  if (fread(tag, 1, 4, fp) != 4 || memcmp(tag, "F0  ", 4) != 0) { printf("Header error.\n"); fclose(fp); return 0; }
  if (fread(tag, 1, 4, fp) != 4 || fread(&frames, 4, 1, fp) != 1 || fread(tag, 1, 4, fp) != 4 ||
      fread(&frame_period, 8, 1, fp) != 1 || frames < 0) { fclose(fp); return 0; }
  const size_t got = fread(f0, 8, frames, fp);
  fclose(fp);
  for (int i = 0; i < frames; ++i) temporal_positions[i] = i / 1000.0 * frame_period;
  return got == (size_t)frames ? 1 : 0;
}

double GetHeaderInformation(const char *filename, const char *parameter) {                   // :119-144
  FILE *fp = fopen(filename, "rb");
  if (!fp) { printf("File cannot be opened.\n"); return 0; }
  // the reference scans 4-byte words; fields are 4-aligned, so walking the header word by word is the same
  char word[4];
  double answer = 0.0;
  for (int i = 0; i < 13 && fread(word, 1, 4, fp) == 4; ++i) {
    if (memcmp(word, parameter, 4) != 0) continue;
    if (memcmp(parameter, "FP  ", 4) == 0) { double v = 0; if (fread(&v, 8, 1, fp) == 1) answer = v; }
    else { int v = 0; if (fread(&v, 4, 1, fp) == 1) answer = static_cast<double>(v); }
    break;
  }
  fclose(fp);
  return answer;
}

void WriteSpectralEnvelope(const char *filename, int fs, int f0_length, double frame_period, int fft_size,
                           int number_of_dimensions, const double *const *spectrogram) {     // :146-172
  write_row_pointers(filename, "SPEC", fs, f0_length, frame_period, fft_size, number_of_dimensions, spectrogram);
}
int ReadSpectralEnvelope(const char *filename, double **spectrogram) {                       // :174-194
  return read_row_pointers(filename, "SPEC", spectrogram);
}
void WriteAperiodicity(const char *filename, int fs, int f0_length, double frame_period, int fft_size,
                       int number_of_dimensions, const double *const *aperiodicity) {        // :196-221
  write_row_pointers(filename, "AP  ", fs, f0_length, frame_period, fft_size, number_of_dimensions, aperiodicity);
}
int ReadAperiodicity(const char *filename, double **aperiodicity) {                          // :223-243
  return read_row_pointers(filename, "AP  ", aperiodicity);
}

int world_b200_write_rows(const char *filename, const char *kind, int fs, int f0_length, double frame_period,
                          int fft_size, int number_of_dimensions, const double *rows) {
  if (!filename || !kind || !rows || f0_length < 0 || fft_size < 2 || number_of_dimensions < 0 ||
      (memcmp(kind, "SPEC", 4) != 0 && memcmp(kind, "AP  ", 4) != 0))
    return WORLD_B200_EINVAL;
  FILE *fp = open_rows_for_write(filename, kind, fs, f0_length, frame_period, fft_size, number_of_dimensions);
  if (!fp) return WORLD_B200_EINVAL;
  const size_t width = number_of_dimensions == 0 ? fft_size / 2 + 1 : number_of_dimensions;
  const size_t wrote = fwrite(rows, 8, width * f0_length, fp);   // rows are contiguous: one write
  fclose(fp);
  return wrote == width * f0_length ? 0 : WORLD_B200_EINVAL;
}

int world_b200_read_rows(const char *filename, const char *kind, double *rows, int max_frames) {
  if (!filename || !kind || !rows) return WORLD_B200_EINVAL;
  FILE *fp = fopen(filename, "rb");
  if (!fp) return WORLD_B200_EINVAL;
  RowHeader h;
  int rc = WORLD_B200_EINVAL;
  if (read_row_header(fp, kind, &h) && h.frames <= max_frames)
    rc = fread(rows, 8, (size_t)h.frames * h.width, fp) == (size_t)h.frames * h.width ? 0 : WORLD_B200_EINVAL;
  fclose(fp);
  return rc;
}

}  // extern "C"
