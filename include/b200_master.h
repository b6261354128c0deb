/* b200_master.h -- C-ABI of libb200master.so, the B200-native mastering hot path.
 *
 * Drop-in boundary for the DSP chain of the reference engine
 * (/root/reference/worker/audio_mastering_engine.py, "ENG"): every entry point below
 * replaces the arithmetic of one reference function; the Python host module
 * python-audio-mastering_b200/audio_mastering_engine.py binds them with ctypes and
 * keeps the reference's function names and `settings` dict.
 *
 * Conventions
 *  - extern "C", plain pointers and sizes, no exceptions across the boundary.
 *  - Every function returns 0 on success or a B200M_ERR_* code;
 *    b200m_last_error(handle) returns the message of the last failure.
 *  - `*_dev` pointers are CUDA device pointers on the handle's device, `*_host`
 *    pointers are host pointers (pinned memory makes the copies asynchronous).
 *  - The caller owns every buffer; the handle owns only its scratch workspace.
 *  - All work is issued on the handle's stream (b200m_set_stream); functions that
 *    return host values synchronise that stream before returning.
 *  - A handle is not re-entrant; use one handle per host thread (the reference is
 *    called from one thread at a time: mastering_gui.py:204, worker/Dockerfile:15).
 *  - There is NO CPU fallback: without a CUDA device b200m_create fails.
 */
#ifndef B200_MASTER_H
#define B200_MASTER_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200M_ABI_VERSION 2

enum {
    B200M_OK = 0,
    B200M_ERR_INVALID = 1,   /* bad argument (-> ValueError in the Python host) */
    B200M_ERR_CUDA = 2,      /* CUDA runtime failure (-> RuntimeError)           */
    B200M_ERR_TOO_SHORT = 3, /* loudness asked for < 400 ms of audio: pyloudnorm's
                                "Audio must have length greater than the block size" */
    B200M_ERR_NOMEM = 4
};

/* PCM sample formats of the staging kernels (interleaved frames). */
enum {
    B200M_FMT_S16 = 0,  /* int16 little endian: the only width the reference handles (ENG:125) */
    B200M_FMT_S24 = 1,  /* packed 3-byte little endian; declared extension: b200m_stage_pcm    */
    B200M_FMT_F32 = 2   /* float32 in [-1, 1]; declared extension: b200m_stage_pcm             */
};

typedef struct b200m_handle b200m_handle;

/* One normalised second-order section: y = b0 x + z0; z0 = b1 x - a1 y + z1; z1 = b2 x - a2 y
 * (scipy sosfilt / lfilter DF2T with a0 == 1). */
typedef struct { double b0, b1, b2, a1, a2; } b200m_biquad;

/* The reference's `settings` dict frozen into a POD (keys: ENG:58-73,84-86,147-150;
 * SURVEY.md App. C). */
typedef struct {
    double saturation;      /* percent, 0 = bypass (ENG:129)                 */
    double bass_boost;      /* dB, low shelf "250 Hz" (ENG:154)              */
    double mid_cut;         /* dB, applied as -mid_cut at "1 kHz" (ENG:156)  */
    double presence_boost;  /* dB, peak "4 kHz" (ENG:158)                    */
    double treble_boost;    /* dB, high shelf "8 kHz" (ENG:160)              */
    double width;           /* 1.0 = bypass (ENG:60)                         */
    int32_t multiband;      /* truthiness of settings["multiband"] (ENG:65)  */
    int32_t has_lufs;       /* 0 <=> settings["lufs"] is None (ENG:84)       */
    double low_thresh, low_ratio, mid_thresh, mid_ratio, high_thresh, high_ratio; /* ENG:67-72 */
    double lufs;            /* target, used when has_lufs                    */
} b200m_settings;

/* One compressor band (pydub compress_dynamic_range arguments, ENG:207-209). */
typedef struct {
    double thresh_rms;      /* 32768 * 10^(threshold_dB/20)                     */
    double attack_frames;   /* attack_ms * (rate/1000.0)  (float, pydub)        */
    double release_frames;  /* release_ms * (rate/1000.0)                       */
    double slope;           /* 1 - 1/ratio                                      */
    int32_t look_frames;    /* int(attack_frames): RMS window [i-look, i)       */
    int32_t reserved;
} b200m_band;

/* Everything the kernels need for one distinct `settings` at one sample rate.
 * Filter DESIGN is O(1) per job and stays on the host: either
 * b200m_plan_from_settings() below (C, libm) or the Python host, which evaluates
 * the reference's own numpy/scipy expressions (ENG:172-182,187-193,197-198). */
typedef struct {
    int32_t sample_rate;
    int32_t channels;        /* 1 or 2 */
    int32_t sat_on;          /* ENG:129 */
    float sat_clean;         /* float32(1 - mix)      ENG:134 */
    float sat_mix;           /* float32(mix)          ENG:131 */
    float sat_drive;         /* float32(1 + 4*mix)    ENG:133 */
    int32_t n_eq;            /* active EQ sections, in application order (ENG:154-161) */
    int32_t width_on;        /* ENG:60 */
    b200m_biquad eq[4];
    double width;
    int32_t multiband;
    int32_t has_lufs;
    b200m_biquad lp[2];      /* butter(4, 250, lowpass) sos   ENG:197 */
    b200m_biquad hp[2];      /* butter(4, 4000, highpass) sos ENG:198 */
    b200m_band band[3];      /* low, mid, high */
    b200m_biquad kw[2];      /* K-weighting: high shelf, high pass (pyloudnorm) */
    double lufs;
    /* ENG:128-134 only ever sees x = int16 / 32768: the exciter is a pure function of the 65536 possible
     * samples.  sat_lut (HOST pointer, 65536 float32, entry [(uint16_t)s] for the int16 sample s) holds
     * (1 - mix) x + mix tanh(x (1 + 4 mix)) as the HOST's numpy evaluates it in float32 -- numpy's float32 tanh
     * is a SIMD polynomial that no device routine reproduces bit for bit, so the Python host tabulates it
     * with numpy itself (b200master/plan.py) and the kernels gather.  NULL: the library tabulates the same
     * expression with its own libm tanhf (<= 1 ulp from numpy's; non-Python hosts).  sat_lut_key names the
     * table's CONTENT for the plan cache (same key = same 65536 values); 0 = the library hashes the table. */
    const float *sat_lut;
    uint64_t sat_lut_key;
} b200m_plan;

/* ---- lifetime ------------------------------------------------------------------ */
int b200m_abi_version(void);
int b200m_create(int device, b200m_handle **out);
void b200m_destroy(b200m_handle *h);
const char *b200m_last_error(const b200m_handle *h); /* h may be NULL: creation error */
int b200m_set_stream(b200m_handle *h, void *cuda_stream);  /* cudaStream_t; NULL = default */
int b200m_synchronize(b200m_handle *h);
/* Upper bound (bytes) the handle may allocate for scratch; default 64 GiB. */
int b200m_set_workspace_limit(b200m_handle *h, int64_t bytes);
/* Number of kernel launches issued by this handle so far (bench.py's gpu_launches). */
int64_t b200m_launch_count(const b200m_handle *h);
/* Average device time (ms) of the named internal kernel over the launches since the
 * last reset (CUDA events on the handle's stream; enable with b200m_set_profiling). */
int b200m_set_profiling(b200m_handle *h, int on);
int b200m_kernel_time_ms(b200m_handle *h, const char *kernel, double *total_ms, int64_t *launches);
int b200m_reset_profile(b200m_handle *h);
/* Compressor recurrence tiling: each (chunk, band) attenuation chain is cut into time tiles that
 * run in parallel after a warm-up over the preceding `warm_frames` (k_comp), wrong guesses are
 * repaired, and an exact sequential pass (k_comp_fix) is the backstop.
 * tile_frames = 0: automatic tile length; warm_frames = 0: automatic: 16384 or 1.75 release times of the band, whichever is more (counted in frames of
 * 32-frame blocks with any activity; silent stretches carry the state unchanged); rounds < 0: automatic repair -- the true state is carried
 * through the wrong stretches by the recurrence alone (k_comp_sprint) and the 1024-frame pieces whose samples change are recomputed in one pass;
 * rounds >= 0: that many Jacobi rounds instead (each a list-building launch plus a re-run of the tiles whose start state was wrong: a round carries
 * the truth one tile further).  Results never depend on any of the three. */
int b200m_set_recur_tiling(b200m_handle *h, int tile_frames, int warm_frames, int rounds);
/* Time segmentation of the filter kernels: k_chain (2048-frame tiles) and k_kweight (4096-sample
 * tiles) cut every stream / track into segments of this many tiles, one CTA each, joined by
 * overlap-discard with a warm-up derived from the pole radii (fp64-exact decay).  0 = automatic,
 * negative = off (one CTA walks the whole stream).  Results do not depend on it. */
int b200m_set_segment_tiles(b200m_handle *h, int chain_tiles, int kweight_tiles);
/* Which kernel runs ENG:117-206: 1 = k_chain (one CTA of eight warps per time segment), 2 = k_chainw (every
 * warp on its own segment: no CTA barriers, shuffle-only exchanges; needs eight times as many segments, so
 * its warm-up share is only small on large batches), 0 = automatic (k_chainw when every warp still gets
 * a run of >= 4 warm-up lengths; environment variable B200M_CHAIN_KERNEL presets it).  The K-weighting ahead of the
 * loudness measurement (ENG:214-218) has the same two shapes, k_kweight and k_kweightw, and follows the same setting
 * (B200M_KWEIGHT_KERNEL presets it alone).  Results do not depend on it. */
int b200m_set_chain_kernel(b200m_handle *h, int mode);
/* Host-buffer pipeline of b200m_master_batch: with host PCM the batch is cut into groups of
 * tracks and the H2D copy of group g+1 / D2H copy of group g-1 run on side streams while the
 * kernels of group g run (three workspace slots).  on = 0 runs copy -> kernels -> copy in sequence.
 * Default on.  Results do not depend on it. */
int b200m_set_pipeline(b200m_handle *h, int on);
/* Shape of that pipeline: the batch is cut into about `groups` groups of whole tracks (0 = default 8)
 * and their kernels run on `compute_streams` streams (1 or 2; 0 = default 1; with 2, neighbouring
 * groups alternate, which measured slower on B200: 68.7 vs 65.0 ms per 64-track step).  Results do
 * not depend on it. */
int b200m_set_pipeline_shape(b200m_handle *h, int groups, int compute_streams);
/* Verification counters since the last reset: tiles repaired by the sequential pass, frames it
 * re-ran, and tiles repaired in the parallel rounds. */
int b200m_recur_stats(b200m_handle *h, int64_t *wrong_tiles, int64_t *rerun_frames, int64_t *round_repairs, int reset);

/* ---- host-side design (replaces ENG:172-182, 187-193, 197-198 + pydub/pyloudnorm
 *      parameter set-up); pure C, no device work -------------------------------- */
int b200m_plan_from_settings(const b200m_settings *s, int sample_rate, int channels, b200m_plan *out);

/* ---- whole path: replaces the chunk loop + loudness + limiter, ENG:46-89 ---------
 * Masters n_tracks independent tracks in one batch.  All tracks share the sample
 * rate / channel count of their plans.
 *   pcm_in        interleaved PCM of all tracks (format `fmt`), device or host.  B200M_FMT_S16 is the reference's
 *                 domain; B200M_FMT_S24 / B200M_FMT_F32 are staged to it first exactly like b200m_stage_pcm
 *                 (declared extension: ENG:125 handles 16-bit PCM only; the output stays int16)
 *   in_offsets    [n_tracks] first frame of each track inside pcm_in          (host)
 *   in_frames     [n_tracks] frames available for each track                  (host)
 *   out_frames    [n_tracks] frames to produce (pydub's ms framing, ENG:51-54: may be
 *                 a few frames less than in_frames, or more = zero padded)    (host)
 *   plans/n_plans distinct plans; plan_index [n_tracks] selects one per track (host)
 *   pcm_out       interleaved int16 (ENG:125 always emits int16), tracks packed back to
 *                 back in out_frames order, device or host
 *   loudness_out  [n_tracks] integrated loudness (LUFS) of the pre-gain signal, NaN
 *                 when the track's plan has no target; may be NULL             (host)
 *   gain_out      [n_tracks] linear gain applied; may be NULL                 (host)
 * Errors: B200M_ERR_TOO_SHORT if a track with a loudness target is < 400 ms.     */
int b200m_master_batch(b200m_handle *h,
                       const void *pcm_in, int in_on_device, int fmt,
                       int n_tracks, const int64_t *in_offsets, const int64_t *in_frames,
                       const int64_t *out_frames,
                       const b200m_plan *plans, int n_plans, const int32_t *plan_index,
                       void *pcm_out, int out_on_device,
                       double *loudness_out, double *gain_out);

/* The same batch for SEVERAL loudness targets at once (a preset x loudness sweep shares
 * everything ahead of ENG:84): ENG:46-82 and the loudness measurement (ENG:212-218) run once per track,
 * gain, limiter and final cast (ENG:219-227, ENG:89) once per target.  Bit-identical to n_targets calls of
 * b200m_master_batch with plan.lufs = targets[k].
 *   targets      [n_targets] LUFS (1..64); every track's plan must have has_lufs, its own lufs is ignored  (host)
 *   pcm_out      n_targets copies of the packed batch output, target-major: copy k starts
 *                k * sum(out_frames) frames after copy 0; device or host
 *   loudness_out [n_tracks], gain_out [n_targets * n_tracks] (gain_out[k * n_tracks + t]); may be NULL     (host) */
int b200m_master_batch_targets(b200m_handle *h,
                               const void *pcm_in, int in_on_device, int fmt,
                               int n_tracks, const int64_t *in_offsets, const int64_t *in_frames,
                               const int64_t *out_frames,
                               const b200m_plan *plans, int n_plans, const int32_t *plan_index,
                               const double *targets, int n_targets,
                               void *pcm_out, int out_on_device,
                               double *loudness_out, double *gain_out);

/* ENG:96-99 `mastered.export(buffer, format="wav")` folded into the batch: every track's output
 * is a complete RIFF/WAVE file image -- the 44-byte header pydub's writer (the stdlib wave module: PCM, 16 bit)
 * emits, written by the GPU immediately ahead of the samples -- so the host writes or uploads the span
 * [44 bytes before the track's first sample, its last sample] as it is, without another sweep over the PCM.
 *   out_offsets [n_tracks] first FRAME of each track's samples inside `out` (host); ascending, each track
 *               preceded by >= 44 free bytes (11 stereo / 22 mono frames) after the end of the previous one and
 *               starting at a multiple of 4 bytes (mono: an even frame);
 *               samples that start at multiples of 16 bytes take the fast store path of the final kernel.
 * Everything else as b200m_master_batch.  b200m_wav_header writes the same 44 bytes on the host. */
int b200m_master_batch_wav(b200m_handle *h,
                           const void *pcm_in, int in_on_device, int fmt,
                           int n_tracks, const int64_t *in_offsets, const int64_t *in_frames,
                           const int64_t *out_frames,
                           const b200m_plan *plans, int n_plans, const int32_t *plan_index,
                           const int64_t *out_offsets, void *out, int out_on_device,
                           double *loudness_out, double *gain_out);
int b200m_wav_header(int sample_rate, int channels, int64_t frames, unsigned char *out44);

/* ---- one long track split along time over several GPUs (BASELINE config 4) ---------
 * Slices are cut at 30-s chunk boundaries (ENG:48-54: every filter and compressor restarts
 * there), so ENG:48-80 is local to a slice.  Only the loudness measurement (ENG:82-86,
 * 212-222) couples slices; the host exchanges two halos of processed samples with its
 * neighbours and SUM-all-reduces the block energies (b200master/longtrack.py, NCCL).
 * All PCM / z buffers below are DEVICE pointers; work is issued on the handle's stream. */

/* Declared extension (ENG:125 handles 16-bit only): reduce packed little-endian 24-bit PCM to
 * the reference's 16-bit domain like pydub's set_sample_width(2) = audioop.lin2lin (keep the
 * high-order 16 bits); float32 goes through the reference's quantiser (ENG:123-126);
 * B200M_FMT_S16 is a copy.  n_samples = frames * channels. */
int b200m_stage_pcm(b200m_handle *h, const void *pcm_dev, int fmt, int64_t n_samples, int16_t *out_dev);
/* Frames of processed audio a slice starting at absolute frame abs_offset needs from the
 * slice before it (K-weighting warm-up, aligned so the buffer starts at a filter tile
 * boundary; 0 for the first slice) and after it (one 400 ms block; clip to the track end). */
int b200m_slice_halo(const b200m_plan *plan, int64_t abs_offset, int64_t *halo_before, int64_t *halo_after);
/* ENG:48-80 for one slice (a whole number of 30-s chunks, or the track's tail): exciter, EQ,
 * width, quantise, multiband -> proc_dev (out_frames frames of interleaved int16).  in_frames /
 * out_frames as in b200m_master_batch (pydub's ms framing applies to the last slice only). */
int b200m_slice_chain(b200m_handle *h, const int16_t *pcm_dev, int64_t in_frames, int64_t out_frames,
                      const b200m_plan *plan, int16_t *proc_dev);
/* K-weighting over [halo | slice | halo] and the 400 ms block energies z_j of every block of
 * the TRACK whose first frame lies in the slice, written to z_dev[j] (j = track block index;
 * other entries are left untouched: start from zeros and SUM-all-reduce).  proc_ext_dev holds
 * ext_frames frames, the slice starts at frame halo_before of it = absolute frame abs_offset. */
int b200m_slice_energies(b200m_handle *h, const int16_t *proc_ext_dev, int64_t ext_frames, int64_t halo_before,
                         int64_t local_frames, int64_t abs_offset, int64_t track_frames, const b200m_plan *plan,
                         double *z_dev, int32_t *first_block_out, int32_t *n_blocks_out);
/* pyloudnorm numBlocks for a track of track_frames frames (0 if shorter than 400 ms). */
int b200m_track_blocks(int64_t track_frames, int rate);
/* Gating over the assembled block energies: integrated loudness and the gain of ENG:219-220. */
int b200m_gate(b200m_handle *h, const double *z_dev, int32_t n_blocks, const b200m_plan *plan, double *loudness_out, double *gain_out);
/* ENG:82-89 for one slice: re-float, gain (has_gain != 0: float64 path; else the float32
 * limiter of a job without loudness target), soft limiter, final quantise. */
int b200m_slice_final(b200m_handle *h, const int16_t *proc_dev, int64_t frames, const b200m_plan *plan,
                      int has_gain, double gain, int16_t *out_dev);

/* ---- stage-level entry points (back the reference's helper functions and the
 *      stage parity tests).  All arrays are HOST pointers; n = frames. ----------- */

/* ENG:117-121 audio_segment_to_float_array: int16 interleaved -> float32 / 2^15. */
int b200m_pcm16_to_float(b200m_handle *h, const int16_t *pcm, int64_t n_samples, float *out);
/* ENG:123-126 float_array_to_audio_segment: clip, *2^15, truncate, +FS wrap.
 * is_f64 selects the input dtype (float32 / float64). */
int b200m_float_to_pcm16(b200m_handle *h, const void *x, int is_f64, int64_t n_samples, int16_t *out);
/* ENG:128-134 apply_saturation on float32 samples (any layout, elementwise). */
int b200m_saturation(b200m_handle *h, const float *x, int64_t n_samples, double saturation_percent, float *out);
/* ENG:117-134 on int16 samples: audio_segment_to_float_array followed by apply_saturation, as ONE gather
 * from the 65536-entry table of b200m_plan::sat_lut (same conventions: host pointer, content key or 0) --
 * bit-exact with the host's numpy, unlike b200m_saturation's tanhf (<= 1 ulp) on arbitrary floats. */
int b200m_saturation_pcm(b200m_handle *h, const int16_t *pcm, int64_t n_samples, const float *sat_lut,
                         uint64_t sat_lut_key, float *out);
/* ENG:136-144 apply_stereo_width on interleaved (n, 2) samples, float32 or float64. */
int b200m_stereo_width(b200m_handle *h, const void *x, int is_f64, int64_t n_frames, double width, void *out);
/* ENG:183/194/200-201 scipy.signal.sosfilt from zero state: n_sections biquads applied in
 * order to each of `channels` interleaved channels; float32 or float64 in, float64 out. */
int b200m_sosfilt(b200m_handle *h, const b200m_biquad *sections, int n_sections,
                  const void *x, int is_f64, int64_t n_frames, int channels, double *out);
/* ENG:196-210 apply_multiband_compressor on one int16 chunk (zero state). */
int b200m_multiband(b200m_handle *h, const b200m_plan *plan, const int16_t *pcm, int64_t n_frames, int16_t *out);
/* pydub compress_dynamic_range on one int16 band (ENG:207-209).  att_out (n_frames
 * doubles, dB) and rms_out (n_frames uint16... stored as uint32) may be NULL. */
int b200m_compress_dynamic_range(b200m_handle *h, const int16_t *pcm, int64_t n_frames, int channels,
                                 const b200m_band *band, int16_t *out, double *att_out, uint32_t *rms_out);
/* pyloudnorm Meter(rate).integrated_loudness of a mono float32 signal (ENG:213-218). */
int b200m_integrated_loudness(b200m_handle *h, const b200m_biquad *kw, const float *mono, int64_t n, int rate, double *lufs_out);
/* ENG:212-222 normalize_to_lufs on float32 samples, interleaved (n_frames, channels) with
 * channels 1 or 2: loudness of the float32 channel mean (ENG:215), one gain (ENG:219-220),
 * out[i] = (double)x[i] * gain (float64, numpy >= 2 promotion).  loudness_out / gain_out
 * may be NULL. */
int b200m_normalize_to_lufs(b200m_handle *h, const b200m_biquad *kw, const float *x, int64_t n_frames,
                            int channels, int rate, double target_lufs, double *out,
                            double *loudness_out, double *gain_out);
/* ENG:224-227 soft_limiter (threshold 0.98, knee 0.02) on float32 or float64 samples. */
int b200m_soft_limiter(b200m_handle *h, const void *x, int is_f64, int64_t n_samples, double threshold, void *out);

#ifdef __cplusplus
}
#endif
#endif /* B200_MASTER_H */
