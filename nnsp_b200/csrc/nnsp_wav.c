/* nnsp_wav.c -- the audio front door: RIFF/WAVE files -> the [stream][frame][160] int16 PCM layout the batched
 * entry points take. Host code, plain C.
 *
 * In the reference the path is fed by the AUDADC interrupt (evb/src/main_nnsp.cc:46-74: one 160-sample, 16 kHz,
 * int16 frame per 10 ms) and, off the device, by the Python tools reading python/test_wavs/\*.wav with soundfile
 * (python/test_s2i.py, test_vad.py, test_kws.py: 16 kHz mono int16, the 44-byte canonical header). This reader takes
 * the same files -- and any other PCM WAVE file at 16 kHz / 16 bit (chunks in any order, WAVE_FORMAT_EXTENSIBLE,
 * several channels: one is picked). There is no resampler: the reference has none either, a different rate is an error. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "nnsp_b200.h"
#include "nnsp_model.h"     /* nnsp_set_error */

typedef struct {
    FILE *f;
    int rate, channels, bits;
    long long data_off, n_samples;      /* byte offset of the sample data; samples per channel */
} wav_file;

static unsigned rd_u16(const unsigned char *p) { return (unsigned)p[0] | ((unsigned)p[1] << 8); }
static unsigned long rd_u32(const unsigned char *p) { return (unsigned long)p[0] | ((unsigned long)p[1] << 8) | ((unsigned long)p[2] << 16) | ((unsigned long)p[3] << 24); }

static int wav_open(const char *path, wav_file *w)
{
    unsigned char h[12], ck[8], fmt[40];
    int have_fmt = 0;
    memset(w, 0, sizeof *w);
    if (!path || !(w->f = fopen(path, "rb"))) { nnsp_set_error("wav: cannot open %s", path ? path : "(null)"); return NNSP_B200_ERR_ARG; }
    if (fread(h, 1, 12, w->f) != 12 || memcmp(h, "RIFF", 4) != 0 || memcmp(h + 8, "WAVE", 4) != 0) {
        nnsp_set_error("wav: %s is not a RIFF/WAVE file", path);
        goto bad;
    }
    for (;;) {
        if (fread(ck, 1, 8, w->f) != 8) { nnsp_set_error("wav: %s has no data chunk", path); goto bad; }
        const unsigned long n = rd_u32(ck + 4);
        if (memcmp(ck, "fmt ", 4) == 0) {
            const size_t take = n < sizeof fmt ? n : sizeof fmt;
            if (n < 16 || fread(fmt, 1, take, w->f) != take) { nnsp_set_error("wav: %s: short fmt chunk", path); goto bad; }
            unsigned tag = rd_u16(fmt);
            w->channels = (int)rd_u16(fmt + 2);
            w->rate = (int)rd_u32(fmt + 4);
            w->bits = (int)rd_u16(fmt + 14);
            if (tag == 0xFFFE && take >= 26) tag = rd_u16(fmt + 24);      /* WAVE_FORMAT_EXTENSIBLE: first word of the sub-format GUID */
            if (tag != 1) { nnsp_set_error("wav: %s: format tag %u is not integer PCM", path, tag); goto bad; }
            if (fseek(w->f, (long)(n - take + (n & 1)), SEEK_CUR) != 0) goto bad;
            have_fmt = 1;
        } else if (memcmp(ck, "data", 4) == 0) {
            if (!have_fmt) { nnsp_set_error("wav: %s: data chunk before fmt chunk", path); goto bad; }
            w->data_off = ftell(w->f);
            if (w->channels < 1 || w->bits != 16) { nnsp_set_error("wav: %s: %d-bit samples (the path takes 16-bit PCM)", path, w->bits); goto bad; }
            w->n_samples = (long long)(n / (2u * (unsigned)w->channels));
            return NNSP_B200_OK;
        } else {
            if (fseek(w->f, (long)(n + (n & 1)), SEEK_CUR) != 0) { nnsp_set_error("wav: %s: truncated chunk", path); goto bad; }
        }
    }
bad:
    fclose(w->f);
    w->f = NULL;
    return NNSP_B200_ERR_ARG;
}

int nnsp_b200_wav_info(const char *path, int *sample_rate, int *channels, int *bits, long long *n_samples)
{
    wav_file w;
    const int rc = wav_open(path, &w);
    if (rc) return rc;
    fclose(w.f);
    if (sample_rate) *sample_rate = w.rate;
    if (channels) *channels = w.channels;
    if (bits) *bits = w.bits;
    if (n_samples) *n_samples = w.n_samples;
    return NNSP_B200_OK;
}

/* frames first_frame .. first_frame + n_frames - 1 (160 samples each) of one channel into pcm[n_frames * 160];
 * what lies past the end of the file is digital silence */
static int wav_read(wav_file *w, const char *path, int channel, long long first_frame, int n_frames, int16_t *pcm, int *frames_read)
{
    if (w->rate != 16000) { nnsp_set_error("wav: %s is sampled at %d Hz; the path runs at 16000 Hz and has no resampler", path, w->rate); return NNSP_B200_ERR_UNSUPPORTED; }
    if (channel < 0 || channel >= w->channels) { nnsp_set_error("wav: %s has %d channel(s), channel %d asked for", path, w->channels, channel); return NNSP_B200_ERR_ARG; }
    const long long want = (long long)n_frames * NNSP_B200_FRAME, first = first_frame * NNSP_B200_FRAME;
    long long have = w->n_samples - first;
    if (have < 0) have = 0;
    if (have > want) have = want;
    memset(pcm, 0, (size_t)want * sizeof(int16_t));
    if (have > 0) {
        if (fseek(w->f, (long)(w->data_off + first * 2 * w->channels), SEEK_SET) != 0) { nnsp_set_error("wav: %s: seek failed", path); return NNSP_B200_ERR_ARG; }
        if (w->channels == 1) {
            unsigned char *raw = (unsigned char *)pcm;                       /* decoded in place, little endian on disk */
            const size_t got = fread(raw, 2, (size_t)have, w->f);
            for (size_t i = got; i-- > 0;) pcm[i] = (int16_t)rd_u16(raw + 2 * i);
            have = (long long)got;
        } else {
            const size_t fb = 2u * (size_t)w->channels;
            unsigned char *raw = (unsigned char *)malloc(fb * 4096);
            if (!raw) return NNSP_B200_ERR_NOMEM;
            long long done = 0;
            while (done < have) {
                const size_t n = (size_t)((have - done) < 4096 ? (have - done) : 4096);
                const size_t got = fread(raw, fb, n, w->f);
                for (size_t i = 0; i < got; i++) pcm[done + (long long)i] = (int16_t)rd_u16(raw + i * fb + 2u * (size_t)channel);
                done += (long long)got;
                if (got < n) break;
            }
            free(raw);
            have = done;
        }
    }
    if (frames_read) *frames_read = (int)((have + NNSP_B200_FRAME - 1) / NNSP_B200_FRAME);
    return NNSP_B200_OK;
}

int nnsp_b200_wav_read_frames(const char *path, int channel, long long first_frame, int n_frames, int16_t *pcm, int *frames_read)
{
    if (!pcm || n_frames <= 0 || first_frame < 0) return NNSP_B200_ERR_ARG;
    wav_file w;
    int rc = wav_open(path, &w);
    if (rc) return rc;
    rc = wav_read(&w, path, channel, first_frame, n_frames, pcm, frames_read);
    fclose(w.f);
    return rc;
}

int nnsp_b200_wav_load_streams(const char *const *paths, int n_streams, int channel, long long first_frame, int n_frames,
                               int16_t *pcm, long long stream_stride, int *frames_read)
{
    if (!paths || !pcm || n_streams <= 0 || n_frames <= 0 || first_frame < 0 || stream_stride < (long long)n_frames * NNSP_B200_FRAME)
        return NNSP_B200_ERR_ARG;
    for (int s = 0; s < n_streams; s++) {
        wav_file w;
        int rc = wav_open(paths[s], &w);
        if (rc) return rc;
        rc = wav_read(&w, paths[s], channel, first_frame, n_frames, pcm + (long long)s * stream_stride, frames_read ? frames_read + s : NULL);
        fclose(w.f);
        if (rc) return rc;
    }
    return NNSP_B200_OK;
}
