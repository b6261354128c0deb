/* C restatement of the per-frame compressor loop -- TEST INFRASTRUCTURE ONLY.
 *
 * Follows pydub 0.25.x effects.compress_dynamic_range (third-party, not under
 * /root/reference; called from worker/audio_mastering_engine.py:207-209; spec in
 * SURVEY.md App. B.1) and the three CPython audioop primitives it uses
 * (Modules/audioop.c: rms = (unsigned)sqrt(sum_sq / n) with double accumulation,
 * mul = floor(fbound(x * f)), fbound clamps to [-2^15, 2^15-1]).  The Python-level
 * expressions are kept operation for operation so that, with the same libm, the
 * attenuation trajectory is bit-identical to the faithful loop in
 * oracle/thirdparty.py (tests/test_oracle.py checks that):
 *   thresh_rms = max_amp * pow(10, thr_db / 20)
 *   over       = 20 * (log(rms / thresh_rms) / log(10))      [math.log(x, 10)]
 *   factor     = pow(10, -att / 20)
 * The look-back window sum is carried as a sliding int64 (exact, like audioop's
 * double accumulation of int16 squares, which never leaves the 2^53 range).
 */
#include <math.h>
#include <stdint.h>

static inline int fbound16(double val)
{
    const double maxval = 32767.0, minval = -32768.0;
    if (val > maxval) val = maxval;
    else if (val < minval + 1.0) val = minval;
    return (int)floor(val);
}

/* pcm/out: interleaved int16, nframes * channels samples.  att_out / rms_out may be
 * NULL; when given they receive the per-frame attenuation (dB) and window RMS. */
void oracle_compress_dynamic_range(const int16_t *pcm, int16_t *out, int64_t nframes, int channels,
                                   int rate, double threshold_db, double ratio,
                                   double attack_ms, double release_ms,
                                   double *att_out, uint32_t *rms_out)
{
    const double max_amp = 32768.0; /* 2^(8*2) / 2 */
    const double thresh_rms = max_amp * pow(10.0, threshold_db / 20.0);
    const double attack_frames = attack_ms * (rate / 1000.0);
    const double release_frames = release_ms * (rate / 1000.0);
    const int64_t look = (int64_t)attack_frames;
    const double slope = 1 - (1.0 / ratio);
    const double log10_ = log(10.0);
    double att = 0.0;
    int64_t winsum = 0; /* sum of squares over frames [max(i-look,0), i) */

    for (int64_t i = 0; i < nframes; ++i) {
        int64_t start = i - look < 0 ? 0 : i - look;
        int64_t nsamp = (i - start) * channels;
        unsigned int rms = 0;
        if (nsamp > 0)
            rms = (unsigned int)sqrt((double)winsum / (double)nsamp);

        double over = 0.0;
        if (rms != 0) {
            double db = 20 * (log((double)rms / thresh_rms) / log10_);
            over = db > 0 ? db : 0.0;
        }
        double max_att = slope * over;
        double inc = max_att / attack_frames;
        double dec = max_att / release_frames;
        if ((double)rms > thresh_rms && att <= max_att) {
            att += inc;
            if (max_att < att) att = max_att;
        } else {
            att -= dec;
            if (0 > att) att = 0.0;
        }
        if (att_out) att_out[i] = att;
        if (rms_out) rms_out[i] = rms;

        const int16_t *f = pcm + i * channels;
        int16_t *o = out + i * channels;
        if (att != 0.0) {
            double factor = pow(10.0, (-att) / 20);
            for (int c = 0; c < channels; ++c)
                o[c] = (int16_t)fbound16((double)f[c] * factor);
        } else {
            for (int c = 0; c < channels; ++c) o[c] = f[c];
        }

        /* slide the window: frame i enters, frame i-look leaves */
        for (int c = 0; c < channels; ++c) winsum += (int64_t)f[c] * f[c];
        if (i - look >= 0) {
            const int16_t *g = pcm + (i - look) * channels;
            for (int c = 0; c < channels; ++c) winsum -= (int64_t)g[c] * g[c];
        }
    }
}
