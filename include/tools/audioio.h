/* tools/audioio.h -- mono PCM WAV reading / writing with the reference's interface and byte layout
 * (reference: tools/audioio.h:17-47, tools/audioio.cpp).  Host-only file glue; the batched ingest of many
 * files is world_b200_wav_parse() + world_b200_pcm_to_double_batch() in world_b200.h. */
#ifndef WORLD_AUDIOIO_H_
#define WORLD_AUDIOIO_H_
#ifdef __cplusplus
extern "C" {
#endif

/* Always writes 16-bit mono PCM (nbit is ignored, as in the reference): int(x * 32767) clamped to int16. */
void wavwrite(const double *x, int x_length, int fs, int nbit, const char *filename);
/* Number of samples; 0 when the file cannot be opened, -1 when it is not a mono PCM WAV. */
int GetAudioLength(const char *filename);
/* x must hold GetAudioLength() doubles; samples are value / 2^(nbit-1). */
void wavread(const char *filename, int *fs, int *nbit, double *x);

#ifdef __cplusplus
}
#endif
#endif /* WORLD_AUDIOIO_H_ */
