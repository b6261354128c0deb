// b200m_device.cuh -- device-side data structures and kernels of the mastering hot path.
//
// Reference being replaced: /root/reference/worker/audio_mastering_engine.py ("ENG").
// Layout of one batch in HBM (all buffers are flat over the batch; a frame = CH samples):
//   pcm_in   interleaved int16 of every track                    (caller's or staged)
//   proc     interleaved int16, the chunk-processed track (ENG:80 `processed_audio`)
//   band[3]  interleaved int16 crossover bands after quantisation (ENG:204-206)
//   rms[3]   uint16 window RMS per frame and band (audioop.rms inside pydub)
//   att[3]   fp64 attenuation trajectory per frame and band (pydub's `attenuation`)
//   kw       float32 K-weighted mono signal (pyloudnorm input_data after both stages)
//   z        fp64 400 ms block mean squares per track
//   out      interleaved int16 final output (ENG:89)
// A *stream* is one (track, 30-s chunk): every filter and compressor restarts from zero
// state there (ENG:48-54), so streams are the unit of parallel work ahead of loudness.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200m {

constexpr int SEG = 16;            // samples per thread segment in the blocked IIR scan
constexpr int NSEG = 128;          // segments per channel per tile (k_chain)
constexpr int TILE = SEG * NSEG;   // 2048 frames per k_chain tile
constexpr int TILE_PAD = TILE + NSEG;  // one pad word per segment: conflict-free smem
constexpr unsigned FULL = 0xffffffffu;

// One biquad section plus everything its blocked parallel-prefix evaluation needs.
// State-space form of the DF2T section: s' = A s + B x, y = s[0] + b0 x with
//   A = [[-a1, 1], [-a2, 0]],  B = [b1 - a1 b0, b2 - a2 b0].
struct SecTab {
    double b0, b1, b2, a1, a2;
    double pad_[3];
    double g[SEG][2];   // g[n] = A^(SEG-1-n) B : end state of a segment from zero state
    double P[5][4];     // A^(SEG * 2^k), k = 0..4 : warp-shuffle scan steps (row major 2x2)
    double PW[4];       // A^(SEG * 32)            : warp-to-warp carry
    double AH[4];       // A^(SEG / 2)             : state in the middle of a lane's segment (section_round_w)
    double pad2_[4];
    double Q[32][4];    // A^(SEG * lane)          : carry-in to each lane's segment start
};
static_assert(sizeof(SecTab) == 200 * 8, "SecTab layout");

struct BandDev {
    double thresh_rms, attack_frames, release_frames, slope;
    double r_attack, r_release;   // correctly rounded 1/attack_frames, 1/release_frames
    int32_t look;                 // int(attack_frames)
    int32_t div_trick;            // 1: M/A and M/R via div_const are exact for this band's curve
    int32_t hold_max;             // curve[r] == 0 exactly for r <= hold_max (-1: no such prefix; then nothing is ever "held")
    int32_t att_bounded;          // 1: 0 <= attenuation <= 5000 dB whatever the input (curve finite, >= +0, max <= 5000): 10^(-att/20) needs no underflow path
};

// Compressor static curve: max attenuation M for each of the 32769 possible integer RMS
// values (0.0 when rms <= threshold, which also encodes "hold").
constexpr int CURVE_N = 32769;

struct PlanDev {
    int32_t rate, channels, sat_on, n_eq, width_on, multiband, has_lufs, pad0_;
    float sat_clean, sat_mix, sat_drive, pad1_;
    double width, lufs;
    BandDev band[3];
    const double *curve[3];       // device pointers, CURVE_N doubles each (NULL if !multiband)
    const int32_t *ptree;         // numpy pairwise-sum tree of a full 400 ms block (k_blocks), or NULL
    const int32_t *htree;         // the same for one 100 ms hop when the block tree is four hop trees (k_hops), or NULL
    int32_t hop;                  // hop length in samples (0: no hop sharing at this rate)
    int32_t sat_sym;              // sat_lut is odd: entry(-s) == -entry(s) bit for bit (k_chainw may index a half table by |s|)
    const float *sat_lut;         // sat_on: 2^15 * exciter(s / 2^15) for the 65536 int16 samples s, indexed by (uint16_t)s (ENG:128-134)
    SecTab eq[4], lp[2], hp[2], kw[2];
};

struct StreamDesc {     // one (track, chunk)
    int64_t in_off;     // first frame in pcm_in
    int64_t out_off;    // first frame in the flat workspace buffers
    int32_t in_frames;  // frames available in pcm_in (the rest of out_frames reads as silence)
    int32_t out_frames; // frames to produce
    int32_t plan;
    int32_t track;
    int32_t blk_off;    // first 32-frame block of this stream in the per-band hold-flag arrays
    int32_t pad_;
};

// A time segment of a stream (k_chain) or of a track (k_kweight): one CTA each.  Segments are
// joined by overlap-discard: the CTA starts `warm` frames early from zero state and only
// stores [begin, end).  `warm` is chosen at plan time from the pole radii so that the
// homogeneous response has decayed below 2^-64 (i.e. to nothing in fp64) by `begin`; segments
// that start at frame 0 of their owner are exact by construction (zero state, ENG:48-54).
struct SegDesc {
    int64_t begin, end; // frames relative to the owner's first frame; begin is a tile multiple
    int32_t owner;      // stream index (k_chain) or track index (k_kweight)
    int32_t warm;       // tile multiple
};

struct TrackDesc {
    int64_t off;        // first frame in the flat workspace buffers (a multiple of 32 frames)
    int64_t dst_off;    // first frame in the group's packed output (tracks back to back)
    int64_t frames;     // out frames
    int64_t zoff;       // first block in z
    int32_t nblocks;    // pyloudnorm numBlocks (of this buffer's share when the track is split along time)
    int32_t plan;
    // A time slice of a longer track (b200m_slice_*): buffer frame 0 is absolute frame abs0 of a
    // track of total_frames frames, and the first loudness block computed here is block j0.
    // Whole tracks: abs0 = 0, total_frames = frames, j0 = 0.
    int64_t abs0, total_frames;
    int32_t j0, pad_;
};

// ------------------------------------------------------------------------------------
// Blocked parallel-prefix evaluation of ONE biquad section over one tile.
// Each thread owns SEG consecutive samples x[] of one channel (32*NW threads per channel).
//   pass 1  end state of the segment from zero state = sum_n g[n] x[n]        (2 FMA/sample)
//   scan    Kogge-Stone over the warp with 2x2 transfer-matrix powers P[k], then the
//           NW warp totals are chained with PW and the tile carry-in            (shuffles)
//   pass 2  the true DF2T recurrence from the now-known start state          (5 FMA/sample)
// `carry` (smem, 2 doubles) holds the section state at the tile start and is advanced to
// the tile end.  Contains exactly one __syncthreads(): call it uniformly across the CTA.
// The lane-independent part of a SecTab.  Passed BY VALUE as a kernel parameter it lives in the constant
// bank: table entries reach the DFMAs without going through the shared-memory / shuffle data pipe, the
// busiest unit of the filter kernels (65 % of peak in k_chainw with the tables in shared memory).  Only
// Q[lane] (lane-dependent) still comes from the shared-memory copy of the SecTab.
struct SecTabC {
    double b0, b1, b2, a1, a2;
    double g[SEG][2];
    double P[5][4];
    double PW[4];
    double AH[4];
};

// U: the lane-independent tables (a SecTab in shared memory or a SecTabC in the constant bank), Q: A^(SEG lane).
template <int NW, typename TU>
__device__ __forceinline__ void section_round(double (&x)[SEG], const TU &U, const double (*__restrict__ Q)[4],
                                              double *carry, double *wtot, int lane, int wid,
                                              bool writer)
{
    const TU *T = &U;
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int n = 0; n < SEG; ++n) {
        s0 = fma(T->g[n][0], x[n], s0);
        s1 = fma(T->g[n][1], x[n], s1);
    }
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        double t0 = __shfl_up_sync(FULL, s0, 1 << k);
        double t1 = __shfl_up_sync(FULL, s1, 1 << k);
        if (lane >= (1 << k)) {
            s0 = fma(T->P[k][0], t0, fma(T->P[k][1], t1, s0));
            s1 = fma(T->P[k][2], t0, fma(T->P[k][3], t1, s1));
        }
    }
    double e0 = __shfl_up_sync(FULL, s0, 1);
    double e1 = __shfl_up_sync(FULL, s1, 1);
    if (lane == 0) { e0 = 0.0; e1 = 0.0; }
    if (lane == 31) { wtot[2 * wid] = s0; wtot[2 * wid + 1] = s1; }
    double w0 = carry[0], w1 = carry[1];
    __syncthreads();
    double m0 = w0, m1 = w1;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
        if (w == wid) { m0 = w0; m1 = w1; }
        double n0 = fma(T->PW[0], w0, fma(T->PW[1], w1, wtot[2 * w]));
        double n1 = fma(T->PW[2], w0, fma(T->PW[3], w1, wtot[2 * w + 1]));
        w0 = n0; w1 = n1;
    }
    if (writer) { carry[0] = w0; carry[1] = w1; }
    double z0 = fma(Q[lane][0], m0, fma(Q[lane][1], m1, e0));
    double z1 = fma(Q[lane][2], m0, fma(Q[lane][3], m1, e1));
    const double b0 = T->b0, b1 = T->b1, b2 = T->b2, na1 = -T->a1, na2 = -T->a2;
    // scipy _sosfilt (DF2T): y = b0 x + z0; z0 = b1 x - a1 y + z1; z1 = b2 x - a2 y.  The terms that do
    // not depend on y are formed first, so the dependent chain is two FMAs per sample (y -> z0 -> y).
#pragma unroll
    for (int n = 0; n < SEG; ++n) {
        const double xn = x[n];
        const double t0 = fma(b1, xn, z1), t1 = b2 * xn;
        const double y = fma(b0, xn, z0);
        z0 = fma(na1, y, t0);
        z1 = fma(na2, y, t1);
        x[n] = y;
    }
}

// ENG:123-126: clip to [-1,1], * 2^15, astype(int16) = truncate toward zero, and the
// +1.0 -> 32768 -> -32768 wrap.  y * 2^15 is exact and F2I.F64.TRUNC truncates and saturates, so
// the clip is done on the integer (ALU pipe, not the fp64 pipe).  NaN must cast to 0 (x86
// cvttsd2si gives 0x80000000, whose low 16 bits are 0); F2I.F64 has no NaN-to-zero mode, so NaN is
// detected on the bit pattern.
template <bool NANCHK = true>
__device__ __forceinline__ int quant16(double y)
{
    const int raw = __double2int_rz(y * 32768.0);
    int iv = max(-32768, min(32768, raw));
    // F2I yields INT_MIN for NaN.  Where NaN cannot occur (k_chain with stable filters: finite input
    // stays finite) the test is compiled out.
    if (NANCHK && ((unsigned long long)__double_as_longlong(y) << 1) > 0xffe0000000000000ull) iv = 0;
    return (int)(short)iv;
}

// The same quantiser for a value that is already scaled by 2^15 (k_chainw runs the crossover on
// 2^15 u: the filters are linear and scaling by a power of two commutes with every rounding, so
// fl(2^15 y) is what y * 32768.0 would have produced).
template <bool NANCHK = true>
__device__ __forceinline__ int quant16s(double ys)
{
    const int raw = __double2int_rz(ys);
    int iv = max(-32768, min(32768, raw));
    if (NANCHK && ((unsigned long long)__double_as_longlong(ys) << 1) > 0xffe0000000000000ull) iv = 0;
    return (int)(short)iv;
}

// ENG:128-134 in float32 with every product / sum separately rounded (numpy semantics), for the stand-alone
// apply_saturation helper on arbitrary floats (tanhf: <= 1 ulp from numpy's SIMD tanh).  The chain kernels
// never call it: their inputs are int16 / 2^15, so they gather from PlanDev::sat_lut, which the host tabulated.
__device__ __forceinline__ float exciter(float x, float clean, float mix, float drive)
{
    float t = tanhf(__fmul_rn(x, drive));
    return __fadd_rn(__fmul_rn(clean, x), __fmul_rn(mix, t));
}

__device__ __forceinline__ int pidx(int f) { return f + (f / SEG); }

// prmt.b32 in its generic form: selector nibbles with bit 3 set replicate the SIGN of the byte they name, so
// 0x9910 / 0xbb32 sign-extend the low / high int16 of a word in one instruction.  (__byte_perm masks that bit.)
__device__ __forceinline__ int prmt_sx(unsigned x, unsigned sel)
{
    int r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(0u), "r"(sel));
    return r;
}

}  // namespace b200m
