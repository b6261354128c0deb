// b200m_kernels.cuh -- the sm_100a kernels of the mastering chain (see b200m_device.cuh).
#pragma once
#include "b200m_device.cuh"

namespace b200m {

struct BandPtrs {
    int16_t *band[3];    // in: quantised crossover bands (ENG:204-206), interleaved
    uint16_t *rms[3];    // per frame: integer window RMS (audioop.rms) -- k_detect -> k_comp
    uint32_t *hold[3];   // bit per 32-frame block of a stream: 1 = rms <= threshold in the whole block (state held)
    double *bend[3];     // per 32-frame block: pydub's `attenuation` after the block's last frame (merge test of the repair passes)
    double *att[3];      // NULL, or per frame: the attenuation trajectory (debug output of b200m_compress_dynamic_range)
};

// =====================================================================================
// k_chain: per (track, chunk) stream, sequential over 2048-frame tiles:
//   int16 -> float32 (ENG:117-121) -> exciter (ENG:128-134) -> up to 4 EQ biquads in
//   fp64 (ENG:146-194) -> M/S width (ENG:136-144) -> quantise #1 (ENG:123-126)
//   and, when the plan is multiband, straight on with ENG:197-206:
//   LP4 / HP4 crossover in fp64, mid = x - low - high, quantise #2 per band.
// One CTA per stream; 128 threads per channel, each owning SEG consecutive samples of a
// tile; every biquad is a section_round (blocked parallel prefix).  Filter state is
// carried tile to tile in shared memory and starts at zero (chunk-boundary semantics).
// =====================================================================================
// Quantised samples are staged planar in shared memory, 16 per thread packed into two 128-bit
// stores; one pad of 8 int16 per 16-sample segment keeps them 16-byte aligned and conflict free.
constexpr int QSEG = 24;                 // int16 slots per segment (16 used)
constexpr int QW = QSEG / 2;             // the same in 32-bit words

__device__ __forceinline__ void stage_q16(unsigned *dst32, const int (&q)[SEG])
{
    uint4 a, b;
    a.x = (unsigned)(q[0] & 0xffff) | ((unsigned)q[1] << 16);   a.y = (unsigned)(q[2] & 0xffff) | ((unsigned)q[3] << 16);
    a.z = (unsigned)(q[4] & 0xffff) | ((unsigned)q[5] << 16);   a.w = (unsigned)(q[6] & 0xffff) | ((unsigned)q[7] << 16);
    b.x = (unsigned)(q[8] & 0xffff) | ((unsigned)q[9] << 16);   b.y = (unsigned)(q[10] & 0xffff) | ((unsigned)q[11] << 16);
    b.z = (unsigned)(q[12] & 0xffff) | ((unsigned)q[13] << 16); b.w = (unsigned)(q[14] & 0xffff) | ((unsigned)q[15] << 16);
    reinterpret_cast<uint4 *>(dst32)[0] = a;
    reinterpret_cast<uint4 *>(dst32)[1] = b;
}

// planar staged tile -> interleaved int16 in global memory, two frames per thread and step
template <int CH>
__device__ __forceinline__ void store_q16_tile(int16_t *__restrict__ dst, const unsigned *src32, int nvalid, int tid)
{
    constexpr int NT = NSEG * CH;
    if (CH == 2) {
        const bool al8 = (reinterpret_cast<unsigned long long>(dst) & 7ull) == 0;
        for (int p = tid; 2 * p < nvalid; p += NT) {
            const int w = (p >> 3) * QW + (p & 7);
            const unsigned L = src32[w], R = src32[NSEG * QW + w];
            const unsigned o0 = __byte_perm(L, R, 0x5410), o1 = __byte_perm(L, R, 0x7632);
            unsigned *g = reinterpret_cast<unsigned *>(dst) + 2 * p;
            if (al8 && 2 * p + 1 < nvalid) *reinterpret_cast<uint2 *>(g) = make_uint2(o0, o1);
            else { g[0] = o0; if (2 * p + 1 < nvalid) g[1] = o1; }
        }
    } else {
        const bool al4 = (reinterpret_cast<unsigned long long>(dst) & 3ull) == 0;
        for (int p = tid; 2 * p < nvalid; p += NT) {
            const unsigned v = src32[(p >> 3) * QW + (p & 7)];
            if (al4 && 2 * p + 1 < nvalid) reinterpret_cast<unsigned *>(dst)[p] = v;
            else { dst[2 * p] = (int16_t)(v & 0xffff); if (2 * p + 1 < nvalid) dst[2 * p + 1] = (int16_t)(v >> 16); }
        }
    }
}

#ifndef B200M_CHAIN_OCC
#define B200M_CHAIN_OCC 2
#endif
// NANCHK = false: every filter of every plan in the launch is stable, so no NaN / Inf can reach a
// quantiser and their NaN -> 0 handling (5 of ~11 instructions, 5 quantisations per sample) is omitted.
template <int CH, bool NANCHK>
__global__ void __launch_bounds__(NSEG * CH, (CH == 2 ? B200M_CHAIN_OCC : 2 * B200M_CHAIN_OCC))
k_chain(const int16_t *__restrict__ pcm_in, const StreamDesc *__restrict__ streams, const SegDesc *__restrict__ segs,
        const PlanDev *__restrict__ plans, int16_t *__restrict__ proc, BandPtrs bp)
{
    constexpr int NT = NSEG * CH;
    constexpr int FPT = TILE / NT;                                       // frames staged per thread and tile
    constexpr size_t SY_BYTES = (size_t)CH * TILE_PAD * 8 > (size_t)3 * CH * NSEG * QSEG * 2
                                    ? (size_t)CH * TILE_PAD * 8 : (size_t)3 * CH * NSEG * QSEG * 2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SecTab *tabs = reinterpret_cast<SecTab *>(smem_raw);                 // eq[4] lp[2] hp[2]
    double *sy = reinterpret_cast<double *>(smem_raw + 8 * sizeof(SecTab));  // [CH][TILE_PAD] (width exchange)
    float *sx = reinterpret_cast<float *>(smem_raw + 8 * sizeof(SecTab) + SY_BYTES);   // [CH][TILE_PAD]
    double *carry = reinterpret_cast<double *>(sx + CH * TILE_PAD);      // [8][CH][2]
    double *wtot = carry + 8 * CH * 2;                                   // [2][CH][4][2]
    unsigned *stq = reinterpret_cast<unsigned *>(sy);                    // aliases sy: [3][CH][NSEG * QW]
    int *sraw = reinterpret_cast<int *>(wtot + 2 * CH * 4 * 2);          // [TILE] raw frames of the next tile

    const SegDesc sg = segs[blockIdx.x];
    const StreamDesc sd = streams[sg.owner];
    const PlanDev *__restrict__ pl = plans + sd.plan;
    const int tid = threadIdx.x;
    const int c = tid / NSEG, j = tid % NSEG, lane = tid & 31, wid = j >> 5;

    {   // stage this plan's section tables, zero the carries
        const double *src = reinterpret_cast<const double *>(pl->eq);
        double *dst = reinterpret_cast<double *>(tabs);
        for (int i = tid; i < 8 * (int)(sizeof(SecTab) / 8); i += NT) dst[i] = src[i];
        for (int i = tid; i < 8 * CH * 2; i += NT) carry[i] = 0.0;
    }
    const int sat_on = pl->sat_on, n_eq = pl->n_eq, width_on = pl->width_on, multiband = pl->multiband;
    const float *__restrict__ lut = pl->sat_lut;
    const double width = pl->width;
    const float widthf = (float)width;

    const int16_t *__restrict__ in = pcm_in + sd.in_off * CH;
    float *myx = sx + c * TILE_PAD + j * (SEG + 1);
    unsigned *myq = stq + (c * NSEG + j) * QW;
    unsigned round = 0;

    // Raw PCM of the tile AFTER the one being filtered is brought in with cp.async (LDGSTS): no
    // registers held, the copies are in flight during the filtering.  Every thread converts
    // exactly the frames it fetched, so cp.async.wait_group is the only synchronisation needed.
    // (Mono frames are 2 bytes, below cp.async's 4-byte minimum: plain loads.)
    auto fetch = [&](int t0) {
#pragma unroll
        for (int k = 0; k < FPT; ++k) {
            const int f = tid + k * NT, gf = t0 + f;
            if (CH == 2) {
                if (gf < sd.in_frames) {
                    const unsigned sa = (unsigned)__cvta_generic_to_shared(sraw + f);
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(reinterpret_cast<const int *>(in) + gf) : "memory");
                } else {
                    sraw[f] = 0;
                }
            } else {
                sraw[f] = gf < sd.in_frames ? (int)__ldg(in + gf) : 0;
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    const int seg_begin = (int)sg.begin, seg_end = (int)sg.end;
    const int t_first = max(0, seg_begin - sg.warm);
    fetch(t_first);
    __syncthreads();
    for (int t0 = t_first; t0 < seg_end; t0 += TILE) {
        const bool store = t0 >= seg_begin;          // warm-up tiles only advance the filter states
        const int nvalid = store ? min(TILE, seg_end - t0) : 0;
        // ---- stage: interleaved int16 -> planar float32 (+ exciter), ENG:117-134 ---------
        // Everything up to quantise #1 runs on 2^15 x: the filters and the widener are linear and scaling by
        // a power of two commutes with every rounding, so the quantiser sees fl(2^15 y) without a multiply.
        // The exciter is a pure function of the int16 sample: one gather from the host-tabulated table.
        asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
        for (int k = 0; k < FPT; ++k) {
            const int f = tid + k * NT;
            const int rw = sraw[f];
            float v[CH];
            if (sat_on) {
                v[0] = __ldg(lut + (rw & 0xffff));
                if (CH == 2) v[CH - 1] = __ldg(lut + ((unsigned)rw >> 16));
            } else {
                v[0] = (float)(short)(rw & 0xffff);
                if (CH == 2) v[CH - 1] = (float)(rw >> 16);
            }
#pragma unroll
            for (int q = 0; q < CH; ++q) sx[q * TILE_PAD + pidx(f)] = v[q];
        }
        if (t0 + TILE < seg_end) fetch(t0 + TILE);
        __syncthreads();

        double x[SEG];
#pragma unroll
        for (int n = 0; n < SEG; ++n) x[n] = (double)myx[n];

        // ---- EQ: one section_round per active biquad (bypassed sections were dropped
        //      at plan time, exactly like ENG:171,186) ------------------------------------
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            if (s < n_eq) {
                section_round<4>(x, tabs[s], tabs[s].Q, carry + (s * CH + c) * 2,
                                 wtot + (((round & 1) * CH + c) * 4) * 2, lane, wid, j == 0);
                ++round;
            }
        }

        // ---- M/S width (needs the other channel: exchange through smem) ----------------
        if (CH == 2 && width_on) {
            double *myy = sy + c * TILE_PAD + j * (SEG + 1);
#pragma unroll
            for (int n = 0; n < SEG; ++n) myy[n] = x[n];
            __syncthreads();
            const double *oy = sy + (1 - c) * TILE_PAD + j * (SEG + 1);
            // L' = mid + side, R' = mid - side with side = (L - R) / 2 * w.  Seen from the R thread,
            // (R - L) / 2 * w is exactly -side (IEEE sign symmetry), so both channels evaluate
            // (me + other) / 2 + (me - other) / 2 * w with identical roundings to the reference.
            if (n_eq > 0) {     // float64 arithmetic (EQ output is float64)
#pragma unroll
                for (int n = 0; n < SEG; ++n) {
                    const double o = oy[n];
                    const double mid = __dmul_rn(__dadd_rn(x[n], o), 0.5);
                    const double side = __dmul_rn(__dmul_rn(__dsub_rn(x[n], o), 0.5), width);
                    x[n] = __dadd_rn(mid, side);
                }
            } else {            // EQ fully bypassed: the reference stays in float32
#pragma unroll
                for (int n = 0; n < SEG; ++n) {
                    const float o = (float)oy[n], me = (float)x[n];
                    const float mid = __fmul_rn(__fadd_rn(me, o), 0.5f);
                    const float side = __fmul_rn(__fmul_rn(__fsub_rn(me, o), 0.5f), widthf);
                    x[n] = (double)__fadd_rn(mid, side);
                }
            }
            __syncthreads();    // sy is reused as the int16 staging area below
        }

        int q[SEG];
        if (!multiband) {
            // ---- quantise #1 -> proc --------------------------------------------------
#pragma unroll
            for (int n = 0; n < SEG; ++n) q[n] = quant16s<NANCHK>(x[n]);
            stage_q16(myq, q);
            __syncthreads();
            store_q16_tile<CH>(proc + (sd.out_off + t0) * CH, stq, nvalid, tid);
            __syncthreads();
        } else {
            // ---- quantise #1, re-float (ENG:199), crossover ------------------------------
#pragma unroll
            for (int n = 0; n < SEG; ++n) {
                const float u = (float)quant16s<NANCHK>(x[n]);     // 2^15 u of ENG:199: the crossover stays in the scaled domain
                myx[n] = u;
                x[n] = (double)u;
            }
            section_round<4>(x, tabs[4], tabs[4].Q, carry + (4 * CH + c) * 2,
                             wtot + (((round & 1) * CH + c) * 4) * 2, lane, wid, j == 0); ++round;
            section_round<4>(x, tabs[5], tabs[5].Q, carry + (5 * CH + c) * 2,
                             wtot + (((round & 1) * CH + c) * 4) * 2, lane, wid, j == 0); ++round;
            // ---- low band staged now; mid = x - low - high (ENG:202) keeps x - low ---------
            double rest[SEG];
#pragma unroll
            for (int n = 0; n < SEG; ++n) {
                const double u = (double)myx[n];
                q[n] = quant16s<NANCHK>(x[n]);
                rest[n] = __dsub_rn(u, x[n]);
                x[n] = u;
            }
            stage_q16(myq + 0 * CH * NSEG * QW, q);
            section_round<4>(x, tabs[6], tabs[6].Q, carry + (6 * CH + c) * 2,
                             wtot + (((round & 1) * CH + c) * 4) * 2, lane, wid, j == 0); ++round;
            section_round<4>(x, tabs[7], tabs[7].Q, carry + (7 * CH + c) * 2,
                             wtot + (((round & 1) * CH + c) * 4) * 2, lane, wid, j == 0); ++round;
#pragma unroll
            for (int n = 0; n < SEG; ++n) q[n] = quant16s<NANCHK>(__dsub_rn(rest[n], x[n]));
            stage_q16(myq + 1 * CH * NSEG * QW, q);
#pragma unroll
            for (int n = 0; n < SEG; ++n) q[n] = quant16s<NANCHK>(x[n]);
            stage_q16(myq + 2 * CH * NSEG * QW, q);
            __syncthreads();
#pragma unroll
            for (int b = 0; b < 3; ++b)
                store_q16_tile<CH>(bp.band[b] + (sd.out_off + t0) * CH, stq + b * CH * NSEG * QW, nvalid, tid);
            __syncthreads();
        }
    }
}

template <int CH>
constexpr size_t chain_smem_bytes()
{
    return 8 * sizeof(SecTab) +
           ((size_t)CH * TILE_PAD * 8 > (size_t)3 * CH * NSEG * QSEG * 2 ? (size_t)CH * TILE_PAD * 8 : (size_t)3 * CH * NSEG * QSEG * 2) +
           (size_t)CH * TILE_PAD * 4 + 8 * CH * 2 * 8 + 2 * CH * 4 * 2 * 8 + (size_t)TILE * 4;
}

// =====================================================================================
// k_chainw: the same chain (ENG:117-206) with every WARP on its own: a warp walks one time segment
// of a stream in tiles of 256 frames (stereo: lanes 0-15 own 16 consecutive samples of L, lanes 16-31
// the same frames of R) or 512 frames (mono), so
//   * the blocked scan of a biquad is a 16- (32-) lane shuffle scan and the tile-to-tile carry is the
//     DF2T state the last lane ends its segment with -- no cross-warp chain, no __syncthreads;
//   * the M/S width exchange is one shuffle (xor 16) per sample instead of a trip through shared memory;
//   * raw PCM comes in as 16-byte cp.async pieces straight into the layout the lanes read back as
//     128-bit words, and the quantised bands leave as 16-byte stores after a shuffle interleave.
// Warps never wait for each other, so their fp64-heavy and integer-heavy phases overlap freely.  Segments are joined by overlap-discard exactly like
// k_chain's; since eight times as many independent segments are needed, the host picks this kernel
// for batches large enough to keep the warm-up share small (b200m_set_chain_kernel).
// One warp per CTA, or sixteen around a shared-memory copy of the exciter table (template parameter CW, below).
// =====================================================================================
// Warps per CTA of k_chainw (template parameter CW).  ONE (the default shape): everything a warp does then
// depends on blockIdx and kernel parameters only, the compiler can prove its control flow uniform, and the
// shuffles lose their divergence guards (the code shrinks from 6400 to 3300 instructions, no spills); with the
// tables in the constant bank a CTA needs 12 KB of shared memory, so 16 of them fit an SM like two CTAs of eight
// warps did.  Measured: 10.5 -> 9.5 ms on 64 tracks, 1.49 -> 1.38 ms on 8.
// SIXTEEN (SLUT: the launch has one plan and its exciter is on): the CTA fills an SM and keeps the exciter
// table in shared memory.  ENG:128-134 is a gather per sample; 32 lanes gathering from a 256 KB table in global
// memory cost the L1 one sector per clock (+3.5 ms per 64-track step, more than the tanhf it replaced), while
// shared memory serves the same gather in ~3 conflict cycles.  The table is odd (numpy's tanh is; checked at
// upload), so |s| <= 32768 indexes 128 KB.  Control flow stays provably uniform: the 16 warps of a CTA walk
// their own segments, but all for the CTA's common number of tiles (cta_iters[blockIdx.x], the host pads with
// empty segments), a warp past its segment's end filtering zeros into nothing.
constexpr int CW_SLUT = 16;
constexpr size_t SLUT_BYTES = 131200;                       // 32769 floats, rounded up to a multiple of 64
template <int CH> struct ChainW {
    static constexpr int NL = 32 / CH;                      // lanes per channel
    static constexpr int WT = NL * SEG;                     // frames per warp tile
    static constexpr int RSTRIDE = CH == 2 ? 20 : 12;       // 32-bit words per 16-frame raw segment (16 / 8 used): conflict-free LDS.128
    static constexpr int RAW_WORDS = NL * RSTRIDE;
    static constexpr int USTRIDE = 20;                      // floats per lane in the re-floated stash
    static constexpr size_t WARP_BYTES = (size_t)RAW_WORDS * 4 + 32 * USTRIDE * 4 + 8 * CH * 2 * 8;
    static constexpr size_t QTAB_BYTES = 8 * 32 * 4 * 8;  // Q[lane] of the eight sections: all the shared memory the tables need when the rest travels as a parameter
    static constexpr size_t SMEM = 8 * sizeof(SecTab) + WARP_BYTES;               // CW = 1, tables in shared memory (several plans in the launch)
    static constexpr size_t SMEM_PT = QTAB_BYTES + WARP_BYTES;                    // CW = 1, tables in the constant bank
    static constexpr size_t SMEM_SLUT = SLUT_BYTES + QTAB_BYTES + CW_SLUT * WARP_BYTES;   // CW = 16, exciter table + Q tables + 16 warps
    static constexpr size_t SMEM_SLUT_MP = SLUT_BYTES + 8 * sizeof(SecTab) + CW_SLUT * WARP_BYTES;   // ... several plans sharing one exciter table: the CTA's plan's tables in shared memory
};

// The lane-independent tables of the launch's single plan (SecTabC, b200m_device.cuh), by value.
struct ChainTabsC { SecTabC sec[8]; };       // eq[4] lp[2] hp[2]: 4160 bytes (kernel parameters may take 32 KB since CUDA 12.1)
struct KwTabsC { SecTabC sec[2]; };          // K-weighting shelf and high-pass (a function of the rate alone)
static_assert(sizeof(ChainTabsC) <= 4224, "ChainTabsC grew: check the constant-bank budget of the chain kernels");

// One biquad over the warp's tile.  j = lane within the channel group of NL lanes; carry (shared
// memory, 2 doubles per section and channel) = the section state at the tile start, replaced by the
// state after the tile's last sample.
// U: the lane-independent tables (a SecTab in shared memory or a SecTabC in the constant bank), Q: A^(SEG j).
template <int NL, typename TU>
__device__ __forceinline__ void section_round_w(double (&x)[SEG], const TU &U, const double (*__restrict__ Q)[4], double *carry, int j)
{
    const TU *T = &U;
    constexpr int H = SEG / 2;
    // zero-state end states of the two HALVES of the lane's segment (g[n + H] = A^(H-1-n) B weighs sample n of a half):
    // f = state after samples 0 .. H-1, s = what samples H .. SEG-1 add; the segment's own is A^H f + s.
    double f0 = 0.0, f1 = 0.0, s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int n = 0; n < H; ++n) {
        f0 = fma(T->g[n + H][0], x[n], f0);
        f1 = fma(T->g[n + H][1], x[n], f1);
        s0 = fma(T->g[n + H][0], x[n + H], s0);
        s1 = fma(T->g[n + H][1], x[n + H], s1);
    }
    s0 = fma(T->AH[0], f0, fma(T->AH[1], f1, s0));
    s1 = fma(T->AH[2], f0, fma(T->AH[3], f1, s1));
#pragma unroll
    for (int k = 0; (1 << k) < NL; ++k) {
        const double t0 = __shfl_up_sync(FULL, s0, 1 << k, NL);
        const double t1 = __shfl_up_sync(FULL, s1, 1 << k, NL);
        if (j >= (1 << k)) {
            s0 = fma(T->P[k][0], t0, fma(T->P[k][1], t1, s0));
            s1 = fma(T->P[k][2], t0, fma(T->P[k][3], t1, s1));
        }
    }
    double e0 = __shfl_up_sync(FULL, s0, 1, NL);
    double e1 = __shfl_up_sync(FULL, s1, 1, NL);
    if (j == 0) { e0 = 0.0; e1 = 0.0; }
    const double c0 = carry[0], c1 = carry[1];
    double z0 = fma(Q[j][0], c0, fma(Q[j][1], c1, e0));
    double z1 = fma(Q[j][2], c0, fma(Q[j][3], c1, e1));
    // the state in the middle of the segment, so that the two halves run as two independent recurrences (the
    // dependent chain of the DF2T step is two FMAs per sample: one chain leaves the fp64 pipe waiting)
    double y0 = fma(T->AH[0], z0, fma(T->AH[1], z1, f0));
    double y1 = fma(T->AH[2], z0, fma(T->AH[3], z1, f1));
    const double b0 = T->b0, b1 = T->b1, b2 = T->b2, na1 = -T->a1, na2 = -T->a2;
#pragma unroll
    for (int n = 0; n < H; ++n) {
        const double xa = x[n], xb = x[n + H];
        const double ta0 = fma(b1, xa, z1), ta1 = b2 * xa;
        const double tb0 = fma(b1, xb, y1), tb1 = b2 * xb;
        const double ya = fma(b0, xa, z0), yb = fma(b0, xb, y0);
        z0 = fma(na1, ya, ta0);
        y0 = fma(na1, yb, tb0);
        z1 = fma(na2, ya, ta1);
        y1 = fma(na2, yb, tb1);
        x[n] = ya;
        x[n + H] = yb;
    }
    __syncwarp();                                   // every lane has read the carry
    if (j == NL - 1) { carry[0] = y0; carry[1] = y1; }
}

// Two INDEPENDENT biquads over the same tile side by side (the low-pass and the high-pass branch of
// the crossover, ENG:200-201): the same arithmetic as two section_round_w calls, written so that the
// two dependent chains interleave (twice the instruction-level parallelism of one).
template <int NL, typename TU>
__device__ __forceinline__ void section_round_w2(double (&xa)[SEG], double (&xb)[SEG], const TU &Ua, const TU &Ub,
                                                 const double (*__restrict__ Qa)[4], const double (*__restrict__ Qb)[4],
                                                 double *carry_a, double *carry_b, int j)
{
    const TU *Ta = &Ua, *Tb = &Ub;
    double a0 = 0.0, a1 = 0.0, b0s = 0.0, b1s = 0.0;
#pragma unroll
    for (int n = 0; n < SEG; ++n) {
        a0 = fma(Ta->g[n][0], xa[n], a0);
        a1 = fma(Ta->g[n][1], xa[n], a1);
        b0s = fma(Tb->g[n][0], xb[n], b0s);
        b1s = fma(Tb->g[n][1], xb[n], b1s);
    }
#pragma unroll
    for (int k = 0; (1 << k) < NL; ++k) {
        const double ta0 = __shfl_up_sync(FULL, a0, 1 << k, NL), ta1 = __shfl_up_sync(FULL, a1, 1 << k, NL);
        const double tb0 = __shfl_up_sync(FULL, b0s, 1 << k, NL), tb1 = __shfl_up_sync(FULL, b1s, 1 << k, NL);
        if (j >= (1 << k)) {
            a0 = fma(Ta->P[k][0], ta0, fma(Ta->P[k][1], ta1, a0));
            a1 = fma(Ta->P[k][2], ta0, fma(Ta->P[k][3], ta1, a1));
            b0s = fma(Tb->P[k][0], tb0, fma(Tb->P[k][1], tb1, b0s));
            b1s = fma(Tb->P[k][2], tb0, fma(Tb->P[k][3], tb1, b1s));
        }
    }
    double ea0 = __shfl_up_sync(FULL, a0, 1, NL), ea1 = __shfl_up_sync(FULL, a1, 1, NL);
    double eb0 = __shfl_up_sync(FULL, b0s, 1, NL), eb1 = __shfl_up_sync(FULL, b1s, 1, NL);
    if (j == 0) { ea0 = 0.0; ea1 = 0.0; eb0 = 0.0; eb1 = 0.0; }
    const double ca0 = carry_a[0], ca1 = carry_a[1], cb0 = carry_b[0], cb1 = carry_b[1];
    double za0 = fma(Qa[j][0], ca0, fma(Qa[j][1], ca1, ea0));
    double za1 = fma(Qa[j][2], ca0, fma(Qa[j][3], ca1, ea1));
    double zb0 = fma(Qb[j][0], cb0, fma(Qb[j][1], cb1, eb0));
    double zb1 = fma(Qb[j][2], cb0, fma(Qb[j][3], cb1, eb1));
    const double ab0 = Ta->b0, ab1 = Ta->b1, ab2 = Ta->b2, ana1 = -Ta->a1, ana2 = -Ta->a2;
    const double bb0 = Tb->b0, bb1 = Tb->b1, bb2 = Tb->b2, bna1 = -Tb->a1, bna2 = -Tb->a2;
#pragma unroll
    for (int n = 0; n < SEG; ++n) {
        const double xan = xa[n], xbn = xb[n];
        const double ta0 = fma(ab1, xan, za1), ta1 = ab2 * xan;
        const double tb0 = fma(bb1, xbn, zb1), tb1 = bb2 * xbn;
        const double ya = fma(ab0, xan, za0), yb = fma(bb0, xbn, zb0);
        za0 = fma(ana1, ya, ta0);
        zb0 = fma(bna1, yb, tb0);
        za1 = fma(ana2, ya, ta1);
        zb1 = fma(bna2, yb, tb1);
        xa[n] = ya;
        xb[n] = yb;
    }
    __syncwarp();
    if (j == NL - 1) { carry_a[0] = za0; carry_a[1] = za1; carry_b[0] = zb0; carry_b[1] = zb1; }
}

// 16 quantised samples per lane -> interleaved int16 in global memory.  Stereo: lane j (L) and lane
// j + 16 (R) hold the two channels of frames 16j .. 16j+15; they swap halves by shuffle, the L lane
// writes frames 0..7 and the R lane frames 8..15 of the segment (32 contiguous bytes each).
template <int CH>
__device__ __forceinline__ void store_q16_w(int16_t *__restrict__ dst, const int (&q)[SEG], int j, int c, int nvalid, bool al16)
{
    unsigned pk[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) pk[i] = (unsigned)(q[2 * i] & 0xffff) | ((unsigned)q[2 * i + 1] << 16);
    if (CH == 2) {
        unsigned o[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const unsigned snd = c ? pk[i] : pk[4 + i];
            const unsigned rcv = __shfl_xor_sync(FULL, snd, 16);
            const unsigned Lw = c ? rcv : pk[i], Rw = c ? pk[4 + i] : rcv;
            o[2 * i] = __byte_perm(Lw, Rw, 0x5410);
            o[2 * i + 1] = __byte_perm(Lw, Rw, 0x7632);
        }
        const int f0 = 16 * j + 8 * c;
        unsigned *g = reinterpret_cast<unsigned *>(dst) + f0;
        if (al16 && f0 + 8 <= nvalid) {
            reinterpret_cast<uint4 *>(g)[0] = make_uint4(o[0], o[1], o[2], o[3]);
            reinterpret_cast<uint4 *>(g)[1] = make_uint4(o[4], o[5], o[6], o[7]);
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) if (f0 + k < nvalid) g[k] = o[k];
        }
    } else {
        const int f0 = 16 * j;
        if (al16 && f0 + 16 <= nvalid) {
            reinterpret_cast<uint4 *>(dst + f0)[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            reinterpret_cast<uint4 *>(dst + f0)[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        } else {
#pragma unroll
            for (int k = 0; k < 16; ++k) if (f0 + k < nvalid) dst[f0 + k] = (int16_t)q[k];
        }
    }
}

#ifndef B200M_CHAINW_OCC
#define B200M_CHAINW_OCC 2
#endif
// PT: the launch uses ONE plan and its lane-independent tables arrive in `ct` (constant bank).
// CW / SLUT: see above (SLUT implies CW == CW_SLUT; cta_iters is only read when CW > 1; without PT the segments of a CTA
// belong to one plan -- the host pads per plan -- and every plan of the launch has the same exciter table).
template <int CH, bool NANCHK, bool PT, int CW = 1, bool SLUT = false>
__global__ void __launch_bounds__(32 * CW, CW == 1 ? B200M_CHAINW_OCC * 8 : 1)
k_chainw(const int16_t *__restrict__ pcm_in, const StreamDesc *__restrict__ streams, const SegDesc *__restrict__ segs, int n_segs,
         const PlanDev *__restrict__ plans, int16_t *__restrict__ proc, BandPtrs bp, const __grid_constant__ ChainTabsC ct,
         const int32_t *__restrict__ cta_iters)
{
    using W = ChainW<CH>;
    constexpr int NL = W::NL, WT = W::WT;
    static_assert(!SLUT || CW == CW_SLUT, "the shared-memory exciter table needs the 16-warp shape");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *slut = reinterpret_cast<float *>(smem_raw);                   // SLUT: 2^15 * exciter(m / 2^15), m = 0 .. 32768
    unsigned char *smem_tab = smem_raw + (SLUT ? SLUT_BYTES : 0);
    SecTab *tabs = reinterpret_cast<SecTab *>(smem_tab);                 // !PT: eq[4] lp[2] hp[2] of the CTA's plan
    double (*sQ)[32][4] = reinterpret_cast<double (*)[32][4]>(smem_tab); //  PT: only their Q[lane] tables
    const int warp = CW == 1 ? 0 : (int)(threadIdx.x >> 5), lane = threadIdx.x & 31;    // one warp per CTA: everything below is CTA-uniform
    unsigned char *wbase = smem_tab + (PT ? W::QTAB_BYTES : 8 * sizeof(SecTab)) + (size_t)warp * W::WARP_BYTES;
    unsigned *raw = reinterpret_cast<unsigned *>(wbase);                 // [NL][RSTRIDE] raw PCM of the tile being fetched
    float *su = reinterpret_cast<float *>(wbase + (size_t)W::RAW_WORDS * 4) + lane * W::USTRIDE;   // this lane's 16 re-floated samples
    double *carry = reinterpret_cast<double *>(wbase + (size_t)W::RAW_WORDS * 4 + 32 * W::USTRIDE * 4);   // [8][CH][2]

    const int sidx = blockIdx.x * CW + warp;
    const SegDesc sg0 = segs[blockIdx.x * CW];                           // the CTA's first segment names the plan (CW > 1: the launch has one)
    const PlanDev *__restrict__ pl = plans + streams[sg0.owner].plan;
    if (PT) {
        const SecTab *src = pl->eq;                                      // eq[4] lp[2] hp[2] are contiguous in PlanDev
        double *dst = reinterpret_cast<double *>(sQ);
        for (int i = threadIdx.x; i < 8 * 128; i += 32 * CW) dst[i] = reinterpret_cast<const double *>(src[i >> 7].Q)[i & 127];
    } else {
        const double *src = reinterpret_cast<const double *>(pl->eq);
        double *dst = reinterpret_cast<double *>(tabs);
        for (int i = threadIdx.x; i < 8 * (int)(sizeof(SecTab) / 8); i += 32 * CW) dst[i] = src[i];
    }
    if (SLUT) {
        // table entry for the int16 sample s lives at (uint16_t)s: [0, 32768) are s >= 0, entry 32768 is s = -32768
        const float *__restrict__ g = pl->sat_lut;
        for (int i = threadIdx.x; i < 32768; i += 32 * CW) slut[i] = __ldg(g + i);
        if (threadIdx.x == 0) slut[32768] = -__ldg(g + 32768);
    }
    if (lane < 8 * CH * 2) carry[lane] = 0.0;
    __syncthreads();                                                     // the only CTA-wide barrier
    if (CW == 1) {
        if (sidx >= n_segs) return;
    }
    const SegDesc sg = segs[sidx];                                       // CW > 1: the host pads the list to whole CTAs with empty segments
    if (CW == 1) {
        if (sg.begin >= sg.end) return;
    }
    const StreamDesc sd = streams[sg.owner];
    const int sat_on = pl->sat_on, n_eq = pl->n_eq, width_on = pl->width_on, multiband = pl->multiband;
    const float *__restrict__ lut = pl->sat_lut;
    const double width = pl->width;
    const double half_width = 0.5 * width;                              // exact
    const float widthf = (float)width;
    const int c = CH == 2 ? lane >> 4 : 0, j = CH == 2 ? lane & 15 : lane;
    const unsigned csel = c ? 0xbb32u : 0x9910u;                        // __byte_perm selector: this lane's channel of a packed frame, sign-extended
    const int in_frames = sg.begin < sg.end ? sd.in_frames : 0;         // an empty (padding) segment reads nothing

    const int16_t *__restrict__ in = pcm_in + sd.in_off * CH;
    const bool in16 = (reinterpret_cast<unsigned long long>(in) & 15ull) == 0;
    // every output buffer of the stream starts at the same frame offset: one alignment test serves all
    const bool out16 = ((reinterpret_cast<unsigned long long>(proc + sd.out_off * CH) | reinterpret_cast<unsigned long long>(bp.band[0] + sd.out_off * CH) |
                         reinterpret_cast<unsigned long long>(bp.band[1] + sd.out_off * CH) | reinterpret_cast<unsigned long long>(bp.band[2] + sd.out_off * CH)) & 15ull) == 0;

    // raw PCM of tile t0 -> raw[]: 64 pieces of 16 bytes, two per lane (zeros past the end of the input)
    auto fetch = [&](int t0) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int p = lane + 32 * h;
            if (CH == 2) {
                const int gf = t0 + 4 * p;
                unsigned *d = raw + W::RSTRIDE * (p >> 2) + 4 * (p & 3);
                if (in16 && gf + 4 <= in_frames) {
                    const unsigned sa = (unsigned)__cvta_generic_to_shared(d);
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(reinterpret_cast<const unsigned *>(in) + gf) : "memory");
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k) d[k] = gf + k < in_frames ? reinterpret_cast<const unsigned *>(in)[gf + k] : 0u;
                }
            } else {
                const int gf = t0 + 8 * p;
                unsigned *d = raw + W::RSTRIDE * (p >> 1) + 4 * (p & 1);
                if (in16 && gf + 8 <= in_frames) {
                    const unsigned sa = (unsigned)__cvta_generic_to_shared(d);
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(in + gf) : "memory");
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const unsigned lo = gf + 2 * k < in_frames ? (unsigned)(unsigned short)in[gf + 2 * k] : 0u;
                        const unsigned hi = gf + 2 * k + 1 < in_frames ? (unsigned)(unsigned short)in[gf + 2 * k + 1] : 0u;
                        d[k] = lo | (hi << 16);
                    }
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    const int seg_begin = (int)sg.begin, seg_end = (int)sg.end;
    const int t_first = max(0, seg_begin - sg.warm);
    // CW == 1: the warp's own tile count (CTA-uniform: it depends on blockIdx alone); CW > 1: the CTA's common count
    const int n_iter = CW == 1 ? (seg_end - t_first + WT - 1) / WT : cta_iters[blockIdx.x];
    fetch(t_first);
    for (int it = 0; it < n_iter; ++it) {
        const int t0 = t_first + it * WT;
        const bool store = t0 >= seg_begin && t0 < seg_end;      // warm-up tiles only advance the filter states
        const int nvalid = store ? min(WT, seg_end - t0) : 0;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        // ---- int16 -> float32 (+ exciter), ENG:117-134: this lane's 16 samples of its channel ----------
        // Everything up to quantise #1 runs on 2^15 x (linear stages, power-of-two scale: every rounding is the
        // reference's, and the quantiser needs no multiply).  The exciter is a pure function of the int16
        // sample: one gather from the table the host tabulated with numpy (PlanDev::sat_lut, pre-scaled).
        double x[SEG];
        {
            int sx[SEG];                             // the lane's 16 samples, sign-extended
            if (CH == 2) {
                const uint4 *rp = reinterpret_cast<const uint4 *>(raw + W::RSTRIDE * j);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint4 w = rp[i];
                    const unsigned ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) sx[4 * i + k] = prmt_sx(ww[k], csel);
                }
            } else {
                const uint4 *rp = reinterpret_cast<const uint4 *>(raw + W::RSTRIDE * j);
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const uint4 w = rp[i];
                    const unsigned ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        sx[8 * i + 2 * k] = prmt_sx(ww[k], 0x9910u);
                        sx[8 * i + 2 * k + 1] = prmt_sx(ww[k], 0xbb32u);
                    }
                }
            }
            if (SLUT) {
                // the table is odd: look up |s| and put the sign back (sat_on is certain: the host chose this shape)
                float v[SEG];
#pragma unroll
                for (int n = 0; n < SEG; ++n) v[n] = __uint_as_float(__float_as_uint(slut[abs(sx[n])]) ^ ((unsigned)sx[n] & 0x80000000u));
#pragma unroll
                for (int n = 0; n < SEG; ++n) x[n] = (double)v[n];
            } else if (sat_on) {
                float v[SEG];
#pragma unroll
                for (int n = 0; n < SEG; ++n) v[n] = __ldg(lut + ((unsigned)sx[n] & 0xffffu));      // all 16 gathers in flight together
#pragma unroll
                for (int n = 0; n < SEG; ++n) x[n] = (double)v[n];
            } else {
#pragma unroll
                for (int n = 0; n < SEG; ++n) x[n] = (double)sx[n];
            }
        }
        __syncwarp();                                // raw[] is free again
        if (CW > 1 || t0 + WT < seg_end) fetch(t0 + WT);        // CW > 1: uniformly (past the input it is a zero fill)

        // ---- EQ (bypassed sections were dropped at plan time, ENG:171,186) ---------------------------
#pragma unroll
        for (int s = 0; s < 4; ++s)
            if (s < n_eq) {
                if (PT) section_round_w<NL>(x, ct.sec[s], sQ[s], carry + (s * CH + c) * 2, j);
                else    section_round_w<NL>(x, tabs[s], tabs[s].Q, carry + (s * CH + c) * 2, j);
            }

        // ---- M/S width: the other channel of the same frames lives 16 lanes away ----------------------
        if (CH == 2 && width_on) {
            if (n_eq > 0) {     // float64 arithmetic (EQ output is float64)
#pragma unroll
                for (int n = 0; n < SEG; ++n) {
                    // mid = (me + o) / 2, side = (me - o) / 2 * w, me' = mid + side (seen from R, (R - L) / 2 * w is exactly
                    // -side).  Halving is exact, so fl(fl(me - o) * 0.5 * w) == fl(fl(me - o) * (0.5 w)) and
                    // fl(fl(me + o) * 0.5 + side) is one FMA: four fp64 instructions instead of six, the same roundings.
                    const double o = __shfl_xor_sync(FULL, x[n], 16);
                    const double side = __dmul_rn(__dsub_rn(x[n], o), half_width);
                    x[n] = fma(__dadd_rn(x[n], o), 0.5, side);
                }
            } else {            // EQ fully bypassed: the reference stays in float32
#pragma unroll
                for (int n = 0; n < SEG; ++n) {
                    const float me = (float)x[n], o = __shfl_xor_sync(FULL, me, 16);
                    const float mid = __fmul_rn(__fadd_rn(me, o), 0.5f);
                    const float side = __fmul_rn(__fmul_rn(__fsub_rn(me, o), 0.5f), widthf);
                    x[n] = (double)__fadd_rn(mid, side);
                }
            }
        }

        int q[SEG];
        const int64_t o0 = (sd.out_off + t0) * CH;
        if (!multiband) {
#pragma unroll
            for (int n = 0; n < SEG; ++n) q[n] = quant16s<NANCHK>(x[n]);
            store_q16_w<CH>(proc + o0, q, j, c, nvalid, out16);
        } else {
            // ---- quantise #1, re-float (ENG:199), crossover: run on 2^15 u, i.e. on the integer itself ----
            // (linear filters, power-of-two scale: every rounding is that of the reference's u = q / 2^15
            // path, and the three band quantisers lose their multiply)
            int *sq = reinterpret_cast<int *>(su);
            double xh[SEG];
#pragma unroll
            for (int n = 0; n < SEG; ++n) {
                const int q1 = quant16s<NANCHK>(x[n]);
                sq[n] = q1;
                x[n] = (double)q1;
                xh[n] = x[n];
            }
            // low-pass and high-pass branches side by side (both start from u)
            if (PT) {
                section_round_w2<NL>(x, xh, ct.sec[4], ct.sec[6], sQ[4], sQ[6], carry + (4 * CH + c) * 2, carry + (6 * CH + c) * 2, j);
                section_round_w2<NL>(x, xh, ct.sec[5], ct.sec[7], sQ[5], sQ[7], carry + (5 * CH + c) * 2, carry + (7 * CH + c) * 2, j);
            } else {
                section_round_w2<NL>(x, xh, tabs[4], tabs[6], tabs[4].Q, tabs[6].Q, carry + (4 * CH + c) * 2, carry + (6 * CH + c) * 2, j);
                section_round_w2<NL>(x, xh, tabs[5], tabs[7], tabs[5].Q, tabs[7].Q, carry + (5 * CH + c) * 2, carry + (7 * CH + c) * 2, j);
            }
            // ---- low band; mid = x - low - high (ENG:202), same order of subtractions -----------------
#pragma unroll
            for (int n = 0; n < SEG; ++n) q[n] = quant16s<NANCHK>(x[n]);
            store_q16_w<CH>(bp.band[0] + o0, q, j, c, nvalid, out16);
#pragma unroll
            for (int n = 0; n < SEG; ++n) q[n] = quant16s<NANCHK>(__dsub_rn(__dsub_rn((double)sq[n], x[n]), xh[n]));
            store_q16_w<CH>(bp.band[1] + o0, q, j, c, nvalid, out16);
#pragma unroll
            for (int n = 0; n < SEG; ++n) q[n] = quant16s<NANCHK>(xh[n]);
            store_q16_w<CH>(bp.band[2] + o0, q, j, c, nvalid, out16);
        }
    }
}

// =====================================================================================
// k_detect: audioop.rms over the look-back window [i-look, i) of both channels, for every
// frame of every band (pydub rms_at, called from compress_dynamic_range; ENG:207-209).
// Exact integer arithmetic: per-tile prefix sums of frame energies in uint64, window sum
// by difference, rms = (unsigned)sqrt(sum/n) reproduced as an integer square root.
// grid = (tiles, streams, bands).
// =====================================================================================
#ifndef B200M_DT
#define B200M_DT 4096
#endif
#ifndef B200M_DNT
#define B200M_DNT 384
#endif
constexpr int DT = B200M_DT;    // frames per tile
constexpr int DNT = B200M_DNT;

__device__ __forceinline__ unsigned window_rms_rn(unsigned long long S, unsigned n, float rn);
// rcp.approx(n), scaled down by 2^-20 (see window_rms_rn)
__device__ __forceinline__ float window_rcp(unsigned n)
{
    float rn;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rn) : "f"((float)n));
    return __fmul_rn(rn, 1.0f - 0x1p-20f);
}
__device__ __forceinline__ unsigned window_rms(unsigned long long S, unsigned n)
{
    if (n == 0) return 0u;
    return window_rms_rn(S, n, window_rcp(n));
}

// largest r with r*r*n <= S  ==  (unsigned)sqrt((double)S / n): the quotient is a multiple of 1/n, so it
// can never sit within double rounding of a perfect square without being one.  n > 0, rn = window_rcp(n)
// (n is the same for every frame past the look-back).
// MUFU-grade estimate from BELOW: (float)S, rcp.approx, the two products and sqrt.approx carry at most
// 2^-24 + 2^-23 + 2^-24 + 2^-24 of relative error on the radicand (half of it survives the root) and 2^-23 on
// the root, < 2.8e-7 together; the factor 1 - 2^-20 on the radicand lowers the root by 4.8e-7.  So the estimate
// is below the true root by between 2.0e-7 and 7.6e-7 of it, < 0.025 absolute at r <= 32768: its floor is r or
// r - 1, never more (an exact square lands on r - 1), and ONE exact compare settles it: S >= (r + 1)^2 n.
// The floor itself is an add (round down) of 2^23: F2I goes through the quarter-rate conversion unit.
// (n <= 2 * 8192 frames of look-back; (r + 1)^2 <= 2^30; S <= n 2^30 < 2^45.)
__device__ __forceinline__ unsigned window_rms_rn(unsigned long long S, unsigned n, float rn)
{
    float q;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(q) : "f"(__fmul_rn((float)S, rn)));
    unsigned r = __float_as_uint(__fadd_rd(q, 8388608.0f)) & 0x7fffffu;
    const unsigned r1 = r + 1u;
    r += (unsigned)(S >= (unsigned long long)(r1 * r1) * n);
    return r;
}

__device__ __forceinline__ unsigned energy2(unsigned w)      // one stereo frame packed as two int16
{
    const int l = (int)(short)(w & 0xffffu), r = (int)w >> 16;
    return (unsigned)(l * l) + (unsigned)(r * r);
}

// elements per thread in the prefix scan of the extended tile [t0 - HP, t0 + DT): a multiple of 4
__host__ __device__ inline int detect_run(int H)
{
    const int HP = (H + 15) & ~15;
    return (((HP + DT + DNT - 1) / DNT) + 3) & ~3;
}

template <int CH>
__global__ void __launch_bounds__(DNT)
k_detect(const StreamDesc *__restrict__ streams, const PlanDev *__restrict__ plans, BandPtrs bp,
         int band_base)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const StreamDesc sd = streams[blockIdx.y];
    const PlanDev *__restrict__ pl = plans + sd.plan;
    const int band = band_base + blockIdx.z;
    const int t0 = blockIdx.x * DT;
    if (!pl->multiband || t0 >= sd.out_frames) return;
    const int H = pl->band[band].look;
    const int HP = (H + 15) & ~15;                  // history kept ahead of the tile (16-byte aligned loads)
    const int ET = HP + DT;                         // extended tile: element k <-> frame t0 - HP + k
    const int R = detect_run(H);                    // R * DNT >= ET
    unsigned long long *P = reinterpret_cast<unsigned long long *>(smem_raw);   // [R * DNT + 4] exclusive prefix sums
    unsigned *e = reinterpret_cast<unsigned *>(P + R * DNT + 4);               // [R * DNT] frame energies
    __shared__ unsigned long long wsum[DNT / 32];
    __shared__ unsigned sbits[DT / 1024];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

    // ---- phase 1: frame energies, four frames per load (coalesced 16 / 8-byte pieces) -----------
    const int16_t *__restrict__ src = bp.band[band] + sd.out_off * CH;
    const bool vec = (reinterpret_cast<unsigned long long>(src) & (CH == 2 ? 15ull : 7ull)) == 0;   // chunk starts at odd rates may not be
    if (vec) {
        // four groups per thread and round: all four loads are issued before the first energy is formed
        for (int g0 = tid; g0 * 4 < R * DNT; g0 += 4 * DNT) {
            uint4 q[4];
            bool ok[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int g = g0 + u * DNT, k = g * 4, f = t0 - HP + k;       // f is a multiple of 4: a group never straddles frame 0
                ok[u] = k < R * DNT && k < ET && f >= 0 && f < sd.out_frames;
                q[u] = make_uint4(0u, 0u, 0u, 0u);
                if (ok[u]) {
                    if (CH == 2) q[u] = __ldg(reinterpret_cast<const uint4 *>(src + (int64_t)f * 2));
                    else { const uint2 h2 = __ldg(reinterpret_cast<const uint2 *>(src + f)); q[u].x = h2.x; q[u].y = h2.y; }
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int g = g0 + u * DNT, k = g * 4, f = t0 - HP + k;
                if (k >= R * DNT) break;
                uint4 ev = make_uint4(0u, 0u, 0u, 0u);
                if (ok[u]) {
                    if (CH == 2) {
                        ev = make_uint4(energy2(q[u].x), energy2(q[u].y), energy2(q[u].z), energy2(q[u].w));
                    } else {
                        const int s0 = (int)(short)(q[u].x & 0xffffu), s1 = (int)q[u].x >> 16, s2 = (int)(short)(q[u].y & 0xffffu), s3 = (int)q[u].y >> 16;
                        ev = make_uint4((unsigned)(s0 * s0), (unsigned)(s1 * s1), (unsigned)(s2 * s2), (unsigned)(s3 * s3));
                    }
                    if (f + 3 >= sd.out_frames) {           // the stream ends inside this group
                        if (f + 1 >= sd.out_frames) ev.y = 0u;
                        if (f + 2 >= sd.out_frames) ev.z = 0u;
                        ev.w = 0u;
                    }
                }
                reinterpret_cast<uint4 *>(e)[g] = ev;
            }
        }
    } else {
        for (int g = tid; g * 4 < R * DNT; g += DNT) {
            const int k = g * 4, f = t0 - HP + k;
            unsigned t[4] = {0u, 0u, 0u, 0u};
            if (k < ET && f >= 0)
                for (int i = 0; i < 4 && f + i < sd.out_frames; ++i) {
                    if (CH == 2) t[i] = energy2(*reinterpret_cast<const unsigned *>(src + (int64_t)(f + i) * 2));
                    else { const int v = src[f + i]; t[i] = (unsigned)(v * v); }
                }
            reinterpret_cast<uint4 *>(e)[g] = make_uint4(t[0], t[1], t[2], t[3]);
        }
    }
    if (tid < DT / 1024) sbits[tid] = 0xffffffffu;                 // blocks past the end count as held
    __syncthreads();

    // ---- phase 2: exclusive prefix sums (uint64: a 960-frame window of full-scale stereo is 2^41) --
    const uint4 *my = reinterpret_cast<const uint4 *>(e + tid * R);
    unsigned long long run = 0;
    for (int i = 0; i < R / 4; ++i) {
        const uint4 v = my[i];
        run += (unsigned long long)v.x + v.y + ((unsigned long long)v.z + v.w);
    }
    unsigned long long inc = run;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long t = __shfl_up_sync(FULL, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) wsum[wid] = inc;
    __syncthreads();
    unsigned long long ex = inc - run;
    for (int w = 0; w < wid; ++w) ex += wsum[w];
    unsigned long long *myP = P + tid * R;
    for (int i = 0; i < R / 4; ++i) {
        const uint4 v = my[i];
        ulonglong2 a, b2;
        a.x = ex; a.y = ex + v.x; ex = a.y + v.y;
        b2.x = ex; b2.y = ex + v.z; ex = b2.y + v.w;
        reinterpret_cast<ulonglong2 *>(myP)[2 * i] = a;
        reinterpret_cast<ulonglong2 *>(myP)[2 * i + 1] = b2;
    }
    if (tid == DNT - 1) P[R * DNT] = ex;
    __syncthreads();

    // ---- phase 3: window sum by difference, integer RMS, hold flags ------------------------------
    // The static curve is a pure function of the integer RMS (32769 values, tabulated at plan
    // time); it is applied inside the recurrence kernel.  Here: the RMS itself (2 bytes per frame)
    // and one flag per 32-frame block saying that rms <= threshold throughout (M == 0: state held).
    // Four frames per thread, 32 apart (lane i of a warp takes frames i, i + 32, i + 64, i + 96 of its
    // 128-frame span): neighbouring lanes read neighbouring 8-byte prefix sums, so every shared-memory load
    // is conflict-free -- this kernel runs at the speed of the shared-memory data pipe (ncu: 88 % of peak
    // when each thread took four CONSECUTIVE frames, whose 32-byte lane stride cost 2- and 4-way bank
    // conflicts) -- the RMS stores are coalesced 64-byte rows, and the four ballots of a span are its four
    // hold bits.
    uint16_t *__restrict__ dst = bp.rms[band] + sd.out_off;
    const int hold_max = pl->band[band].hold_max;                  // curve[r] == 0  <=>  r <= hold_max
    const int nvalid = min(DT, sd.out_frames - t0);
    const unsigned nH = (unsigned)CH * (unsigned)H;
    float rnH = 0.0f;
    if (nH) rnH = window_rcp(nH);
    for (int c0 = wid * 128; c0 < ((nvalid + 127) & ~127); c0 += (DNT / 32) * 128) {      // whole warps
        unsigned m4 = 0u;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int i = c0 + lane + 32 * k, f = t0 + i;
            bool act = false;
            if (i < nvalid) {
                unsigned r;
                if (f >= H) {                       // the whole look-back lies inside the stream: n = CH * H
                    r = nH ? window_rms_rn(P[HP + i] - P[HP + i - H], nH, rnH) : 0u;
                } else {
                    r = window_rms(P[HP + i] - P[HP + i - f], (unsigned)CH * (unsigned)f);
                }
                dst[f] = (uint16_t)r;
                act = (int)r > hold_max;
            }
            if (__ballot_sync(FULL, act) != 0u) m4 |= 1u << k;
        }
        if (lane == 0 && m4 != 0u) atomicAnd(&sbits[c0 >> 10], ~(m4 << ((c0 >> 5) & 31)));
    }
    __syncthreads();
    if (tid < DT / 1024 && tid * 1024 < nvalid) bp.hold[band][sd.blk_off + (t0 >> 10) + tid] = sbits[tid];
}

// =====================================================================================
// k_detectw: the same window RMS (audioop.rms over [i-look, i) of both channels) with every WARP walking a
// contiguous span of one (stream, band) in steps of 256 frames, eight per lane.  The window sum slides:
// S(i+1) = S(i) + e(i) - e(i-look), so a step needs only the LOCAL prefix of d = e_lead - e_trail (eight adds per
// lane and one 64-bit warp scan) on top of the sum carried from the step before -- no far reads of prefix sums,
// no CTA barriers; the trail energies are the lead energies of `look` frames ago, kept in a per-warp ring in shared
// memory.  ~35 instructions per frame and band against k_detect's 82 (whose CTA-wide prefix scan moves 40 bytes of
// shared memory per frame); k_detect stays for look-backs beyond the ring (sample rates above ~170 kHz).
// grid = (spans / DW_WARPS, streams, bands).
// =====================================================================================
constexpr int DW_WARPS = 4;         // warps per CTA
constexpr int DW_STEP = 256;        // frames per warp step (eight per lane)
constexpr int DW_SPAN = 8192;       // frames per warp: a multiple of 1024 (whole hold words)
constexpr int DW_RING = 2048;       // frame energies kept per warp: look + DW_STEP must fit
constexpr int DW_MAX_LOOK = DW_RING - DW_STEP;

template <int CH>
__global__ void __launch_bounds__(32 * DW_WARPS)
k_detectw(const StreamDesc *__restrict__ streams, const PlanDev *__restrict__ plans, BandPtrs bp, int band_base)
{
    __shared__ __align__(16) unsigned ring_all[DW_WARPS][DW_RING];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const StreamDesc sd = streams[blockIdx.y];
    const PlanDev *__restrict__ pl = plans + sd.plan;
    const int band = band_base + blockIdx.z;
    const int t0 = (blockIdx.x * DW_WARPS + warp) * DW_SPAN;
    if (!pl->multiband || t0 >= sd.out_frames) return;             // warps never synchronise with each other
    const int t1 = min(t0 + DW_SPAN, sd.out_frames);
    const int H = pl->band[band].look, hold_max = pl->band[band].hold_max;
    const unsigned nH = (unsigned)CH * (unsigned)H;
    float rnH = 0.0f;
    if (nH) rnH = window_rcp(nH);
    const int16_t *__restrict__ src = bp.band[band] + sd.out_off * CH;
    uint16_t *__restrict__ dst = bp.rms[band] + sd.out_off;
    uint32_t *__restrict__ hold = bp.hold[band] + sd.blk_off;
    unsigned *ring = ring_all[warp];
    for (int i = lane; i < DW_RING / 4; i += 32) reinterpret_cast<uint4 *>(ring)[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncwarp();
    // 16-byte accesses where the stream's buffers allow them (chunk starts at odd rates may not be aligned)
    const bool src16 = (reinterpret_cast<unsigned long long>(src) & 15ull) == 0;
    const bool dst16 = (reinterpret_cast<unsigned long long>(dst) & 15ull) == 0;
    const bool h4 = (H & 3) == 0;                                   // the trail group of a lane is then two aligned 16-byte ring spans
    // The walk starts `look` frames (rounded up to whole steps) ahead of the span with S = 0 and an empty ring: frames
    // the ring has not seen read as zero energy, so by frame t0 the sum is exactly that of [t0 - look, t0).
    long long S_base = 0;
    unsigned hword = 0xffffffffu;                                   // hold bits of the current 1024-frame word (blocks past the end count as held)
    const int i_start = t0 == 0 ? 0 : t0 - ((H + DW_STEP - 1) / DW_STEP) * DW_STEP;
    // the lane's eight frames of a step, as packed words (stereo: one frame per word; mono: two): issued one step
    // ahead, so the loads of step i + 1 are in flight while step i is being worked on
    auto load_lead = [&](int fl, unsigned (&w)[8]) {
        const bool inside = fl >= 0 && fl + 8 <= sd.out_frames;
        if (inside && src16) {
            if (CH == 2) {
                const uint4 a = __ldg(reinterpret_cast<const uint4 *>(src + (int64_t)fl * 2)), b = __ldg(reinterpret_cast<const uint4 *>(src + (int64_t)fl * 2) + 1);
                w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
            } else {
                const uint4 a = __ldg(reinterpret_cast<const uint4 *>(src + fl));
                w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w;
            }
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const bool ok = fl + k >= 0 && fl + k < sd.out_frames;
                if (CH == 2) w[k] = ok ? *reinterpret_cast<const unsigned *>(src + (int64_t)(fl + k) * 2) : 0u;
                else {
                    const unsigned v = ok ? (unsigned)(unsigned short)src[fl + k] : 0u;
                    if (k & 1) w[k >> 1] |= v << 16; else w[k >> 1] = v;
                }
            }
        }
    };
    unsigned wn[8];
    load_lead(i_start + 8 * lane, wn);
    for (int i0 = i_start; i0 < t1; i0 += DW_STEP) {
        const int f = i0 + 8 * lane;
        // ---- lead: this lane's eight frames -> energies -> ring --------------------------------------------------
        unsigned el[8];
        {
            unsigned w[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) w[k] = wn[k];
            if (i0 + DW_STEP < t1) load_lead(f + DW_STEP, wn);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (CH == 2) {
                    const int l = prmt_sx(w[k], 0x9910u), r = prmt_sx(w[k], 0xbb32u);
                    el[k] = (unsigned)(l * l) + (unsigned)(r * r);
                } else {
                    const int v = prmt_sx(w[k >> 1], (k & 1) ? 0xbb32u : 0x9910u);
                    el[k] = (unsigned)(v * v);
                }
            }
            uint4 *rp = reinterpret_cast<uint4 *>(ring + (f & (DW_RING - 1)));     // f is a multiple of 8: one aligned span, no wrap
            rp[0] = make_uint4(el[0], el[1], el[2], el[3]);
            rp[1] = make_uint4(el[4], el[5], el[6], el[7]);
        }
        __syncwarp();
        // ---- trail: the energies of `look` frames ago, from the ring ------------------------------------------------
        unsigned et[8];
        if (h4) {
            const uint4 a = *reinterpret_cast<const uint4 *>(ring + ((f - H) & (DW_RING - 1)));
            const uint4 b = *reinterpret_cast<const uint4 *>(ring + ((f - H + 4) & (DW_RING - 1)));
            et[0] = a.x; et[1] = a.y; et[2] = a.z; et[3] = a.w; et[4] = b.x; et[5] = b.y; et[6] = b.z; et[7] = b.w;
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) et[k] = ring[(f - H + k) & (DW_RING - 1)];
        }
        __syncwarp();                                               // the next step's lead writes may land on this step's trail slots
        // ---- window sums: S(f + k) = S_base + (exclusive prefix of d over the step) ------------------------------
        long long p[8];
        long long run = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) { p[k] = run; run += (long long)el[k] - (long long)et[k]; }
        long long inc = run;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const long long t = __shfl_up_sync(FULL, inc, d);
            if (lane >= d) inc += t;
        }
        const long long mine = S_base + (inc - run);
        S_base += __shfl_sync(FULL, inc, 31);
        if (i0 < t0) continue;                                      // warm-up of the sum: nothing to emit (warp-uniform)
        // ---- integer RMS, stores, hold bits ---------------------------------------------------------------------
        unsigned r[8];
        bool act = false;
        if (i0 >= H && nH != 0u) {                                  // the whole look-back lies inside the stream: n = CH * look (warp-uniform)
            // no test inside the loop: the eight roots are independent chains (conversion, MUFU, multiply, compare) that
            // must interleave; a per-element "nH ?" compiles to a branch around each of them
#pragma unroll
            for (int k = 0; k < 8; ++k) r[k] = window_rms_rn((unsigned long long)(mine + p[k]), nH, rnH);
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) r[k] = f + k >= H ? (nH ? window_rms_rn((unsigned long long)(mine + p[k]), nH, rnH) : 0u)
                                                          : window_rms((unsigned long long)(mine + p[k]), (unsigned)CH * (unsigned)(f + k));
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) act = act || (f + k < sd.out_frames && (int)r[k] > hold_max);
        if (dst16 && f + 8 <= sd.out_frames) {
            *reinterpret_cast<uint4 *>(dst + f) = make_uint4(r[0] | (r[1] << 16), r[2] | (r[3] << 16), r[4] | (r[5] << 16), r[6] | (r[7] << 16));
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) if (f + k < sd.out_frames) dst[f + k] = (uint16_t)r[k];
        }
        // a 32-frame block = four lanes; the step's eight blocks are bits (i0 / 32) % 32 .. + 7 of hold word i0 / 1024
        const unsigned bal = __ballot_sync(FULL, act);
        unsigned busy = 0u;
#pragma unroll
        for (int b = 0; b < 8; ++b) busy |= ((bal >> (4 * b)) & 0xfu) ? (1u << b) : 0u;
        hword &= ~(busy << ((i0 >> 5) & 31));
        if (((i0 + DW_STEP) & 1023) == 0 || i0 + DW_STEP >= t1) {
            if (lane == 0) hold[i0 >> 10] = hword;
            hword = 0xffffffffu;
        }
    }
}

// =====================================================================================
// k_comp / k_comp_fix: pydub compress_dynamic_range after the level detector (ENG:207-209 run per
// chunk, state reset to 0 at every chunk) for ALL bands of a stream, fused with the gain
// application and the overlay (ENG:210): static curve, attenuation recurrence, 10^(-att/20),
// audioop.mul, two saturating adds -- the attenuation trajectory never goes to HBM.  With
// M_i = curve[rms_i] the per-frame maximum attenuation (M_i != 0  <=>  rms_i > threshold):
//   if M_i != 0 and att <= M_i:  att = min(att + M_i/A, M_i)   else  att = max(att - M_i/R, 0)
//   frame_i *= 10^(-att/20) via audioop.mul (floor of the clamped product) when att != 0
// Every decision of the recurrence is an integer compare on bit patterns (att >= +0 always), both
// candidate sums are formed in parallel, and M/A, M/R are exact constant divisions (3
// instructions, off the dependent chain).
//
// The recurrence is not a linear scan, but two trajectories coincide for ever
// once they are equal, and clamps (att = M, att = 0) make them equal.  So each stream is cut into
// time tiles; a CTA takes 32 tiles, one lane each, and gives every band its own warp:
//   mode 0  speculate: per band, warm up over the preceding `warm` ACTIVE frames from att = 0
//           (no stores; held stretches are skipped via the hold words), then run the tile with
//           the CTA's bands in lockstep, recording per band the assumed start state, the reached
//           end state and the state at the end of every 32-frame block (bp.bend);
//   mode 1  repair round (Jacobi): every tile in which some band's assumed start differs from its
//           predecessor's current end state is re-run from those states (samples included) until
//           all bands meet their stored trajectory at the end of a block;
// and k_comp_fix finally walks each stream's tiles in order and repairs, sequentially, whatever
// is still inconsistent -- so the result is exact for any input and any tile length, and the
// sequential path only runs in the worst case.
//
// Data movement per 32-frame block of a warp (32 lanes = 32 different tiles of one band): the 64-byte RMS
// rows and the 128-byte sample rows of the 32 tiles are brought into shared memory with cp.async one
// block ahead; phase B expands RMS -> M row-wise (one lane per frame: coalesced table gathers); then
// every lane walks its own row: recurrence step, and -- off the dependent chain, so neighbouring frames
// interleave -- 10^(-att/20) and the floor-multiply, leaving the compressed frames in shared memory;
// after a CTA barrier the three bands' rows are overlaid row-wise (two saturating adds in band order)
// and leave as coalesced 128-byte stores of `proc`.  HBM traffic per frame: 6 B of RMS and 12 B of band
// samples in (plus warm-up re-reads of the RMS), 4 B out, 0.75 B of block-end states -- the fp64
// attenuation rows (24 B written + 24 B read per frame when k_apply was a kernel of its own) never exist.
// =====================================================================================
#ifndef B200M_WARM_RELEASES
#define B200M_WARM_RELEASES 1.75
#endif
#ifndef B200M_COMP_MINB
#define B200M_COMP_MINB 4           // resident CTAs per SM k_comp is built for (5 fit since the RMS rows lost their own storage, but on a 64-track group the shorter tiles -- 23 k frames behind a 16 k warm-up -- cost more than the extra warps bring: 10.7 vs 8.7 ms)
#endif
#ifndef B200M_COMP_UNROLL
#define B200M_COMP_UNROLL 2         // 128-bit sample words (4 stereo frames each) per iteration of the lane-serial walk
#endif
constexpr int COMP_UNROLL = B200M_COMP_UNROLL;
#ifndef B200M_RECUR_GB
#define B200M_RECUR_GB 32           // rows of curve gathers in flight per batch (phase B): 32 beats 16 by 1.4 ms per 64-track step
#endif

constexpr int SROW = 36;            // 32-bit words per sample row (32 used; mono: 16): lane-serial 128-bit accesses stay conflict free
struct RecurWarpSmem {              // one per warp = per band of the CTA's 32 (stream, tile) lanes
    double m[32][34];               // M rows in (phase B, row-wise), read back lane-wise (128-bit) by the recurrence, which leaves
                                    // the block's compressed frames in the first 128 bytes of each row
    unsigned smp[32][SROW];         // band samples of the block being worked on: one packed frame per word (mono: two)
                                    // RMS rows (64 bytes each) have no storage of their own: during the warm-up they are staged
                                    // in the (idle) sample rows, afterwards in bytes 128..191 of the M rows, which the walk has
                                    // consumed by the time the next block's RMS values are asked for (rms_row below)
    unsigned long long base_curve[32];
    double etab[32];                // 2^(j/32) for the gains (exp10_gain)
    ulonglong2 row[32];             // per lane: {first workspace frame of its current block, frames to produce (0 = none)}:
                                    // one 16-byte broadcast load per row in the row-wise overlay
};
constexpr size_t recur_smem_bytes(int nb) { return (size_t)nb * (sizeof(RecurWarpSmem) + 32); }

struct RecurParams {
    int tile_len, warm, tiles, nbands, band_base, n_streams, mode;
};

// correctly rounded m / c from the correctly rounded reciprocal rc (Markstein); the host
// verifies it against true division for every value of the band's curve at plan time.
__device__ __forceinline__ double div_const(double m, double c, double rc, bool exact)
{
    if (!exact) return __ddiv_rn(m, c);
    const double q = __dmul_rn(m, rc);
    const double r = fma(-q, c, m);
    return fma(r, rc, q);
}

__device__ __forceinline__ double recur_step(double a, double M, double inc, double dec)
{
    const long long ab = __double_as_longlong(a), Mb = __double_as_longlong(M);
    const bool p = (Mb != 0) && (ab <= Mb);                  // rms > thr and att <= M
    const bool q2 = ab < __double_as_longlong(dec);          // att - dec < 0  -> max() picks 0
    const double u = __dadd_rn(a, inc), d = __dsub_rn(a, dec);
    const bool q1 = __double_as_longlong(u) >= Mb;           // att + inc >= M -> min() picks M (u, M >= 0 here)
    const double vu = q1 ? M : u, vd = q2 ? 0.0 : d;
    return p ? vu : vd;
}

// The same step for curves whose values are all finite and >= +0 (checked at plan time; true for
// every ratio >= 1): "M != 0" is implied -- with M == 0 the attack branch gives min(a + 0, 0) = 0
// only when a == 0, which is what the release branch gives too -- and max(a - dec, 0) is a sign
// test on the difference.  Six integer compares become three and a half.
__device__ __forceinline__ double recur_step_pos(double a, double M, double inc, double dec)
{
    const long long ab = __double_as_longlong(a), Mb = __double_as_longlong(M);
    const bool p = ab <= Mb;                                 // (as fp64 compares -- one instruction each, on the fp64 pipe -- the kernel gains 1 %)
    const double u = __dadd_rn(a, inc), d = __dsub_rn(a, dec);
    const bool q1 = __double_as_longlong(u) >= Mb;
    const bool q2 = __double2hiint(d) < 0;                   // a - dec < 0 (a - dec == -0 cannot occur: a, dec >= +0 and RN)
    const double vu = q1 ? M : u, vd = q2 ? 0.0 : d;
    return p ? vu : vd;
}

// audioop.mul: floor(fbound(sample * factor)), fbound clipping to [-32768, 32767].  The factor here is a
// gain in [0, 1] -- the attenuation is never negative (the recurrence clamps at +0) and exp10_gain(x <= 0)
// <= 1 -- so |v g| <= |v| <= 32768 and the floor is in range without the clip.
// The floor is an add, rounded down, of 1.5 * 2^52 (the integer then sits in the low word, two's complement):
// F2I.F64 occupies the fp64 pipe for 6.5 cycles per warp against DADD's 2.5 (scripts/ubench/cvt_rate.cu).
__device__ __forceinline__ int mul_floor16_le1(int v, double g)
{
    return __double2loint(__dadd_rd(__dmul_rn((double)v, g), 6755399441055744.0));
}

// 10^x for the gain of an attenuation: x = -att / 20 <= 0 and far from underflow (att is at most
// slope * 20 log10(32768 / thresh_rms) dB; the caller falls back to the library routine beyond
// 10^-300).  32 x log2(10) is split as m + f with the 1.5 * 2^52 trick, m = 32 n + j;
// r = x - m log10(2) / 32 (two-piece constant, |r| <= 0.0047); 10^x = 2^n * T[j] * 10^r with
// T[j] = 2^(j/32) from a 32-entry table (`tab`: shared memory in k_comp, lanes gather) and
// 10^r - 1 = r q(r), q the degree-5 Taylor quotient (next term < 4e-18), so the result is ONE
// fused multiply-add T + T * (r q): the table entry's rounding and the final rounding are the only
// errors of size.  2^n goes straight into the exponent field.  Coefficients are constant-bank
// operands.  Eleven fp64 operations (the degree-13 polynomial without a table took seventeen).
// Max error against 60-digit arithmetic: 0.99 ulp (scripts/exp10_check.py simulates this
// sequence with exact FMA semantics), the class of the library's exp10 -- far inside what
// floor(sample * gain) can see.
__constant__ double c_exp10[11] = {
    0x1.0000000000000p+0, 0x1.26bb1bbb55516p+1, 0x1.53524c73cea69p+1, 0x1.0470591de2ca4p+1,
    0x1.2bd7609fd98c4p+0, 0x1.1429ffd1d4d76p-1, 0x1.a7ed70847c8b6p-3,
    0x1.a934f0979a371p+6,       // [7] 32 log2(10)
    -0x1.34413509f79ffp-7,      // [8] -log10(2) / 32, high part
    0x1.9dc1da994fd21p-64,      // [9] -log10(2) / 32, low part
    6755399441055744.0};        // [10] 1.5 * 2^52
__device__ const double g_exp10_tab[32] = {
    0x1.0000000000000p+0, 0x1.059b0d3158574p+0, 0x1.0b5586cf9890fp+0, 0x1.11301d0125b51p+0, 0x1.172b83c7d517bp+0, 0x1.1d4873168b9aap+0,
    0x1.2387a6e756238p+0, 0x1.29e9df51fdee1p+0, 0x1.306fe0a31b715p+0, 0x1.371a7373aa9cbp+0, 0x1.3dea64c123422p+0, 0x1.44e086061892dp+0,
    0x1.4bfdad5362a27p+0, 0x1.5342b569d4f82p+0, 0x1.5ab07dd485429p+0, 0x1.6247eb03a5585p+0, 0x1.6a09e667f3bcdp+0, 0x1.71f75e8ec5f74p+0,
    0x1.7a11473eb0187p+0, 0x1.82589994cce13p+0, 0x1.8ace5422aa0dbp+0, 0x1.93737b0cdc5e5p+0, 0x1.9c49182a3f090p+0, 0x1.a5503b23e255dp+0,
    0x1.ae89f995ad3adp+0, 0x1.b7f76f2fb5e47p+0, 0x1.c199bdd85529cp+0, 0x1.cb720dcef9069p+0, 0x1.d5818dcfba487p+0, 0x1.dfc97337b9b5fp+0,
    0x1.ea4afa2a490dap+0, 0x1.f50765b6e4540p+0};
__device__ __noinline__ double exp10_far(double x) { return exp10(x); }
// FAR = false: the caller knows x > -300 (BandDev::att_bounded) -- no branch, so that several gains of one thread
// form independent chains the scheduler can interleave.
template <bool FAR>
__device__ __forceinline__ double exp10_gain(double x, const double *__restrict__ tab)
{
    if (FAR && !(x > -300.0)) return exp10_far(x);
    const double t = fma(x, c_exp10[7], c_exp10[10]);
    const double mf = __dsub_rn(t, c_exp10[10]);
    double r = fma(mf, c_exp10[8], x);
    r = fma(mf, c_exp10[9], r);
    double p = c_exp10[6];
#pragma unroll
    for (int k = 5; k >= 1; --k) p = fma(p, r, c_exp10[k]);
    const double sr = __dmul_rn(p, r);
    const int m = __double2loint(t);
    const double T = tab[m & 31];
    const double v = fma(T, sr, T);
    // n << 20 with n = m >> 5, written as (m & ~31) * 2^15 so that mask, shift and add are two instructions (|m| < 2^15: no overflow)
    return __hiloint2double(__double2hiint(v) + (m & ~31) * 32768, __double2loint(v));
}

// pydub: frame * db_to_float(-attenuation), db_to_float(db) = 10 ** (db / 20).  The quotient att / 20 is
// rounded like CPython's true division (one Markstein correction of att * fl(1/20): exact for every att,
// checked over 4e8 values on the host), so the exponent handed to 10^x is pydub's, bit for bit.  pydub
// multiplies only `if attenuation != 0.0`; 10^-0 is exactly 1.0 here (the polynomial at r = 0 is its
// constant term) and floor(v * 1.0) == v, so the test needs no branch.
template <bool FAR>
__device__ __forceinline__ double gain_of_att(double a, const double *__restrict__ tab)
{
    const double q = __dmul_rn(a, 0.05);
    const double r = fma(-q, 20.0, a);
    return exp10_gain<FAR>(-fma(r, 0.05, q), tab);
}

// one frame (packed: CH == 2: L | R << 16, CH == 1: the low 16 bits) through audioop.mul
template <int CH>
__device__ __forceinline__ unsigned mul_frame(unsigned smp, double g)
{
    const int v0 = mul_floor16_le1((int)(short)(smp & 0xffffu), g);
    if (CH == 2) {
        const int v1 = mul_floor16_le1((int)smp >> 16, g);
        return __byte_perm((unsigned)v0, (unsigned)v1, 0x5410);
    }
    return (unsigned)v0 & 0xffffu;
}

// audioop.add on a packed frame: per-sample saturating int16 add (pydub overlay, ENG:210).  The SIMD-in-a-word
// intrinsic is six instructions for both samples (VIADD.16x2 + fix-up); min / max on unpacked halves cost 25.
template <int CH>
__device__ __forceinline__ unsigned add_frame_sat(unsigned x, unsigned y)
{
    return __vaddss2(x, y);          // mono: the upper halves are zero and stay zero
}

template <int CH>
__device__ __forceinline__ unsigned load_frame(const int16_t *__restrict__ p, int64_t f)
{
    if (CH == 2) return __ldg(reinterpret_cast<const unsigned *>(p) + f);
    return (unsigned)(unsigned short)__ldg(p + f);
}

template <int CH>
__device__ __forceinline__ void store_frame(int16_t *__restrict__ p, int64_t f, unsigned v)
{
    if (CH == 2) reinterpret_cast<unsigned *>(p)[f] = v;
    else p[f] = (int16_t)v;
}

// NB = 3: the crossover's bands 0..2, one WARP of the CTA per band (band_base == 0); NB = 1: the single-band helper
// entry point (band `band_base`), one warp per CTA.  Lane l of every warp of CTA c works on (stream, tile) number 32 c + l.
template <int CH, int NB, bool DBG>
__global__ void __launch_bounds__(32 * NB, B200M_COMP_MINB)
k_comp(const StreamDesc *__restrict__ streams, const PlanDev *__restrict__ plans, RecurParams P, BandPtrs bp,
       int16_t *__restrict__ proc, const double *__restrict__ ss_in, const double *__restrict__ se_in,
       double *__restrict__ ss_out, double *__restrict__ se_out,
       const unsigned *__restrict__ dirty_list, const unsigned *__restrict__ dirty_count)
{
    extern __shared__ __align__(16) unsigned char recur_smem[];
    const int wid = NB == 1 ? 0 : (int)(threadIdx.x >> 5), lane = threadIdx.x & 31;
    // mode 1 (repair round): the lanes take the (stream, tile) pairs k_comp_dirty listed, densely packed
    unsigned n_dirty = 0u;
    if (P.mode >= 1) {
        n_dirty = *dirty_count;
        if (blockIdx.x * 32u >= n_dirty) return;                         // CTA-uniform
    }
    RecurWarpSmem *WS = reinterpret_cast<RecurWarpSmem *>(recur_smem);
    RecurWarpSmem &W = WS[wid];
    unsigned char (*s_same)[32] = reinterpret_cast<unsigned char (*)[32]>(recur_smem + NB * sizeof(RecurWarpSmem));   // [NB][32]
    const int band = NB == 3 ? wid : P.band_base;
    const unsigned li = blockIdx.x * 32u + lane;
    const int gl = P.mode >= 1 ? (li < n_dirty ? (int)dirty_list[li] : P.n_streams * P.tiles) : (int)li;
    const int s = gl / P.tiles, tile = gl % P.tiles;
    bool live = s < P.n_streams;
    int start = 0, end = 0;
    int64_t out_off = 0;
    const PlanDev *__restrict__ pl = plans;
    const uint32_t *hold_b = nullptr;
    const uint16_t *rms_b = nullptr;
    double *bend_b = nullptr;
    double a = 0.0;
    W.base_curve[lane] = 0ull;
    W.etab[lane] = g_exp10_tab[lane];
    bool okfast = true;
    if (live) {
        const StreamDesc sd = streams[s];
        pl = plans + sd.plan;
        start = tile * P.tile_len;
        live = pl->multiband && start < sd.out_frames;
        if (live) {
            end = min(start + P.tile_len, sd.out_frames);
            out_off = sd.out_off;
            hold_b = bp.hold[band] + sd.blk_off;
            rms_b = bp.rms[band] + sd.out_off;
            bend_b = bp.bend[band] + (size_t)sd.blk_off * 32;
            W.base_curve[lane] = (unsigned long long)pl->curve[band];
            okfast = pl->band[band].div_trick != 0 && pl->band[band].att_bounded != 0;
        }
    }
    const size_t slot0 = ((size_t)s * NB) * P.tiles + tile;            // band b: slot0 + b * P.tiles
    const size_t slot = slot0 + (size_t)(NB == 3 ? wid : 0) * P.tiles;
    double a_in = 0.0;
    if (P.mode == 1 && live) {
        // repair round: some band's assumed start state was not its predecessor's current end state (k_comp_dirty).
        // Every band of the tile is re-run from its predecessor's end; one that was consistent reproduces its
        // stored trajectory.
        a = tile == 0 ? 0.0 : se_in[slot - 1];
        a_in = a;
    }
    // mode 2: recompute a piece whose true start state k_comp_sprint left in the block-end states (no speculation,
    // no merging: the whole piece, from the state at the end of the block before it)
    if (P.mode == 2 && live && start > 0) a = bend_b[(start >> 5) - 1];
    // warp-uniform: this band passed the plan-time checks in every plan of the warp (exact constant division, curve
    // finite and >= +0, attenuation bounded): the branch-free step and gain
    const bool fast = __all_sync(FULL, okfast || !live);
    const int sb = start >> 5, eb = (end + 31) >> 5;
    const BandDev &bd = pl->band[band];
    const double A = bd.attack_frames, R = bd.release_frames, rA = bd.r_attack, rR = bd.r_release;
    const bool ex = bd.div_trick != 0;
    __syncwarp();

    // the hold word of blocks 32w .. 32w+31, kept while the cursor stays inside it
    int hw_idx = -1;
    unsigned hw_val = 0u;
    auto hold_word = [&](int w) { if (w != hw_idx) { hw_idx = w; hw_val = hold_b[w]; } return hw_val; };

    // The RMS rows of every lane's block `cbn` -> W.rms: 64 bytes per row, moved as 16-byte cp.async
    // pieces, eight rows per instruction (lane l: row q0 + l / 4, piece l % 4, so every 32-byte sector is
    // asked for once and the shared-memory side is one contiguous 512-byte span).  The source address
    // of a row travels by shuffle from the lane that owns it.  Rows of held blocks, of lanes without
    // work and the ragged last block of a stream are written by their owner: zeros (r = 0 gives
    // M = 0, a hold, i.e. an identity step, so partial rows need no branches later) plus whatever
    // valid elements there are.  Returns whether this lane's block can change its state.
    bool warming = P.mode == 0;                 // where RMS rows are staged (see RecurWarpSmem)
    auto rms_row = [&](int q) -> uint16_t * {
        return warming ? reinterpret_cast<uint16_t *>(W.smp[q]) : reinterpret_cast<uint16_t *>(&W.m[q][16]);
    };
    auto issue_rms = [&](int cbn, bool valid) -> bool {
        const bool on_n = valid && cbn < eb;
        const int i0n = cbn << 5;
        const int cntn = on_n ? min(32, end - i0n) : 0;
        const bool heldn = on_n && ((hold_word(cbn >> 5) >> (cbn & 31)) & 1u);
        const bool work_n = on_n && !heldn;
        const bool rms16 = (reinterpret_cast<unsigned long long>(rms_b) & 15ull) == 0;    // chunk starts at odd rates may not be
        const uint16_t *srcp = (work_n && cntn == 32 && rms16) ? rms_b + i0n : nullptr;
        if (srcp == nullptr) {
            uint16_t *row = rms_row(lane);
#pragma unroll
            for (int j = 0; j < 4; ++j) reinterpret_cast<uint4 *>(row)[j] = make_uint4(0u, 0u, 0u, 0u);
            if (work_n) {
#pragma unroll 1
                for (int k = 0; k < cntn; ++k) row[k] = rms_b[i0n + k];
            }
        }
        const int sub = lane >> 2, piece = lane & 3;
#pragma unroll
        for (int q0 = 0; q0 < 32; q0 += 8) {
            const unsigned long long p = __shfl_sync(FULL, (unsigned long long)srcp, q0 + sub);
            if (p != 0ull) {
                const unsigned sa = (unsigned)__cvta_generic_to_shared(rms_row(q0 + sub) + piece * 8);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(p + 16ull * piece) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        return work_n;
    };
    // The band samples of every lane's block `cbn` -> W.smp (main part of the tile only), the same way: 64 * CH
    // bytes per row as 16-byte cp.async pieces; unaligned or ragged rows are copied by their owner (whole row
    // zeroed first).  A warp's time is the length of its serial path (all lanes of the grid are resident at once),
    // so the samples land in shared memory one block ahead and nothing below waits for DRAM.
    const int16_t *__restrict__ band_smp = bp.band[band];
    auto issue_smp = [&](int cbn, bool valid) {
        constexpr int PIECES = 4 * CH;                 // 16-byte pieces per row
        const bool on_n = valid && cbn < eb;
        const int i0n = cbn << 5;
        const int cntn = on_n ? min(32, end - i0n) : 0;
        const int16_t *base = band_smp + (out_off + i0n) * CH;
        const bool al = (reinterpret_cast<unsigned long long>(base) & 15ull) == 0;
        const int16_t *srcp = (on_n && cntn == 32 && al) ? base : nullptr;
        if (srcp == nullptr && on_n) {
#pragma unroll
            for (int j = 0; j < PIECES; ++j) reinterpret_cast<uint4 *>(W.smp[lane])[j] = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll 1
            for (int k = 0; k < cntn; ++k) {
                if (CH == 2) W.smp[lane][k] = reinterpret_cast<const unsigned *>(base)[k];
                else reinterpret_cast<uint16_t *>(W.smp[lane])[k] = (uint16_t)base[k];
            }
        }
        const int sub = lane / PIECES, piece = lane % PIECES;
#pragma unroll
        for (int q0 = 0; q0 < 32; q0 += 32 / PIECES) {
            const unsigned long long p = __shfl_sync(FULL, (unsigned long long)srcp, q0 + sub);
            if (p != 0ull) {
                const unsigned sa = (unsigned)__cvta_generic_to_shared(&W.smp[q0 + sub][piece * 4]);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(p + 16ull * piece) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // phase B: RMS rows -> M rows through the static curve (row-wise: consecutive frames have neighbouring
    // RMS values, so a row's gathers touch a few cache lines); B200M_RECUR_GB rows of gathers in flight together
    auto phase_b = [&]() {
#pragma unroll
        for (int q0 = 0; q0 < 32; q0 += B200M_RECUR_GB) {
            double v[B200M_RECUR_GB];
#pragma unroll
            for (int j = 0; j < B200M_RECUR_GB; ++j) {
                const ulonglong2 cp2 = *reinterpret_cast<const ulonglong2 *>(&W.base_curve[(q0 + j) & ~1]);   // one load serves two rows
                const double *curve = reinterpret_cast<const double *>((j & 1) ? cp2.y : cp2.x);
                const unsigned r = rms_row(q0 + j)[lane];
                v[j] = r != 0u ? __ldg(curve + r) : 0.0;    // a row without a curve (a lane without work) is all zeros
            }
#pragma unroll
            for (int j = 0; j < B200M_RECUR_GB; ++j) W.m[q0 + j][lane] = v[j];
        }
    };

    // ---- mode 0: warm up over the preceding frames until `warm` ACTIVE frames have been seen -------------------
    // Walk the hold words (32 blocks = 1024 frames each) backwards; a held stretch carries the state
    // unchanged, so it neither helps nor costs anything.  Every lane then walks its own cursor over the
    // blocks [wstart, sb), skipping held blocks outright, so the warp iterates max-over-lanes of the
    // blocks that need work, not the span.  Only the recurrence runs here: no gains, no samples, no stores.
    if (P.mode == 0) {
        int cb = sb;
        if (live) {
            // automatic: 16384 ACTIVE frames or 1.75 release times, whichever is more.  Two trajectories from different
            // pasts stay apart while the higher one is still releasing towards the static curve (about one release time
            // after a transient, longer when the curve value is small).  A tile that starts from a wrong guess costs a
            // repair pass whose length is that of its slowest lane -- up to a whole tile -- so a longer warm-up for
            // every lane is the cheaper side: on the benchmark programme (kicks every 0.5 s) 8192 frames leave 36 % of
            // the tiles dirty (main pass 7.8 ms + repairs 7.6 ms per 64 tracks), 16384 leave 2 % (8.8 + 0.9 ms).
            const int warm_frames = P.warm > 0 ? P.warm : max(16384, (int)(B200M_WARM_RELEASES * R));
            int need = (warm_frames + 31) >> 5, wb = sb;
            while (wb > 0 && need > 0) {
                const int w = (wb - 1) >> 5, lo = w << 5, nbits = wb - lo;
                const unsigned mask = nbits == 32 ? 0xffffffffu : ((1u << nbits) - 1u);
                need -= nbits - __popc(hold_b[w] & mask);
                wb = lo;
            }
            cb = wb;
        }
        auto skip_held = [&](int c0) {           // first block >= c0 that needs work, or sb
            if (live) {
                while (c0 < sb) {
                    const unsigned wv = hold_word(c0 >> 5) >> (c0 & 31);
                    if (!(wv & 1u)) break;
                    const int run = (~wv) ? __ffs(~wv) - 1 : 32;     // run of held blocks (shifted-in zeros end it)
                    c0 = min(c0 + run, sb);
                }
            }
            return c0;
        };
        cb = skip_held(cb);
        issue_rms(cb, live && cb < sb);
        for (;;) {
            const bool on = live && cb < sb;
            if (!__any_sync(FULL, on)) break;
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncwarp();
            phase_b();                                            // skip_held left only blocks with work
            __syncwarp();
            const int nb = on ? skip_held(cb + 1) : cb;
            issue_rms(nb, on && nb < sb);
            if (on) {
                const double2 *mrow = reinterpret_cast<const double2 *>(W.m[lane]);
                if (fast) {
#pragma unroll 4
                    for (int k2 = 0; k2 < 16; ++k2) {
                        const double2 M = mrow[k2];
                        a = recur_step_pos(a, M.x, div_const(M.x, A, rA, true), div_const(M.x, R, rR, true));
                        a = recur_step_pos(a, M.y, div_const(M.y, A, rA, true), div_const(M.y, R, rR, true));
                    }
                } else {
#pragma unroll 1
                    for (int k2 = 0; k2 < 16; ++k2) {
                        const double2 M = mrow[k2];
                        a = recur_step(a, M.x, div_const(M.x, A, rA, ex), div_const(M.x, R, rR, ex));
                        a = recur_step(a, M.y, div_const(M.y, A, rA, ex), div_const(M.y, R, rR, ex));
                    }
                }
            }
            cb = nb;
            __syncwarp();
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
    }
    warming = false;
    const double a_start = a;

    // ---- the tile itself: the CTA's bands in lockstep over its 32-frame blocks ---------------------------------
    // Per block: phase B (row-wise: one lane per frame) turns the RMS rows into M rows; then every lane walks ITS
    // OWN row: 32 recurrence steps, and right behind each step -- off the dependent chain, so the 10^x polynomials
    // of neighbouring frames interleave freely -- the gain and audioop.mul on the frame's samples; the compressed
    // frames go back into the bytes of the lane's M row that the walk has already consumed (16 bytes per four
    // frames against 32 bytes of M).  Then (CTA barrier) the bands' compressed rows are overlaid row-wise, one
    // lane per frame, every warp of the CTA taking a third of the rows: two saturating adds in band order
    // (ENG:210) and a coalesced 128-byte store of `proc`; a second barrier frees the rows for the next block.
    int cb = sb;
    bool merged = false;
    bool work = issue_rms(cb, live);
    issue_smp(cb, live);
    double *__restrict__ att_dbg = DBG ? bp.att[band] : nullptr;     // DBG: the helper entry point wants the trajectory itself
    for (;;) {
        const bool on = live && !merged && cb < eb;
        if (!__any_sync(FULL, on)) break;
        const int i0 = cb << 5;
        const int cnt = on ? min(32, end - i0) : 0;
        const double old_end = (P.mode == 1 && on) ? bend_b[cb] : 0.0;     // what this tile stored earlier at the end of this block
        W.row[lane] = make_ulonglong2((unsigned long long)(out_off + i0), (unsigned long long)cnt);
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        const bool any_work = __any_sync(FULL, on && work);       // else every lane's block is held: att stays put
        // nothing moves and nothing is attenuated: the band passes through (a band that stays under its threshold)
        const bool quiet = !__any_sync(FULL, on && (work || __double_as_longlong(a) != 0ll));
        if (any_work) phase_b();
        __syncwarp();
        {
            uint4 *crow = reinterpret_cast<uint4 *>(W.m[lane]);      // compressed frames: word w overwrites M[2w], M[2w+1] (consumed)
            const double2 *mrow = reinterpret_cast<const double2 *>(W.m[lane]);
            const uint4 *srow = reinterpret_cast<const uint4 *>(W.smp[lane]);
            constexpr int NW = 8 * CH / 2;                          // 128-bit words per row: 8 stereo (4 frames each), 4 mono (8 frames each)
            if (on) {
                if (quiet) {
#pragma unroll
                    for (int w = 0; w < NW; ++w) crow[w] = srow[w];
                } else {
                    if (!any_work) {
#pragma unroll
                        for (int k2 = 0; k2 < 16; ++k2) reinterpret_cast<double2 *>(W.m[lane])[k2] = make_double2(0.0, 0.0);   // held throughout: identity steps
                    }
                    const int64_t fdbg = out_off + i0;
                    if (fast) {
#pragma unroll COMP_UNROLL
                        for (int w = 0; w < NW; ++w) {
                            const uint4 sw = srow[w];
                            const unsigned sv[4] = {sw.x, sw.y, sw.z, sw.w};
                            unsigned v[4];
                            if (CH == 2) {
                                const double2 Ma = mrow[2 * w], Mb = mrow[2 * w + 1];
                                const double Mv[4] = {Ma.x, Ma.y, Mb.x, Mb.y};
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    a = recur_step_pos(a, Mv[i], div_const(Mv[i], A, rA, true), div_const(Mv[i], R, rR, true));
                                    if (DBG && att_dbg != nullptr && 4 * w + i < cnt) att_dbg[fdbg + 4 * w + i] = a;
                                    v[i] = mul_frame<2>(sv[i], gain_of_att<false>(a, W.etab));
                                }
                            } else {                                 // mono: two frames per 32-bit word
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    const double2 M = mrow[4 * w + i];
                                    a = recur_step_pos(a, M.x, div_const(M.x, A, rA, true), div_const(M.x, R, rR, true));
                                    if (DBG && att_dbg != nullptr && 8 * w + 2 * i < cnt) att_dbg[fdbg + 8 * w + 2 * i] = a;
                                    const unsigned lo = mul_frame<1>(sv[i] & 0xffffu, gain_of_att<false>(a, W.etab));
                                    a = recur_step_pos(a, M.y, div_const(M.y, A, rA, true), div_const(M.y, R, rR, true));
                                    if (DBG && att_dbg != nullptr && 8 * w + 2 * i + 1 < cnt) att_dbg[fdbg + 8 * w + 2 * i + 1] = a;
                                    const unsigned hi = mul_frame<1>(sv[i] >> 16, gain_of_att<false>(a, W.etab));
                                    v[i] = lo | (hi << 16);
                                }
                            }
                            crow[w] = make_uint4(v[0], v[1], v[2], v[3]);
                        }
                    } else {
#pragma unroll 1
                        for (int w = 0; w < NW; ++w) {
                            const uint4 sw = srow[w];
                            const unsigned sv[4] = {sw.x, sw.y, sw.z, sw.w};
                            unsigned v[4];
                            if (CH == 2) {
                                const double2 Ma = mrow[2 * w], Mb = mrow[2 * w + 1];
                                const double Mv[4] = {Ma.x, Ma.y, Mb.x, Mb.y};
#pragma unroll 1
                                for (int i = 0; i < 4; ++i) {
                                    a = recur_step(a, Mv[i], div_const(Mv[i], A, rA, ex), div_const(Mv[i], R, rR, ex));
                                    if (DBG && att_dbg != nullptr && 4 * w + i < cnt) att_dbg[fdbg + 4 * w + i] = a;
                                    v[i] = mul_frame<2>(sv[i], gain_of_att<true>(a, W.etab));
                                }
                            } else {
#pragma unroll 1
                                for (int i = 0; i < 4; ++i) {
                                    const double2 M = mrow[4 * w + i];
                                    a = recur_step(a, M.x, div_const(M.x, A, rA, ex), div_const(M.x, R, rR, ex));
                                    if (DBG && att_dbg != nullptr && 8 * w + 2 * i < cnt) att_dbg[fdbg + 8 * w + 2 * i] = a;
                                    const unsigned lo = mul_frame<1>(sv[i] & 0xffffu, gain_of_att<true>(a, W.etab));
                                    a = recur_step(a, M.y, div_const(M.y, A, rA, ex), div_const(M.y, R, rR, ex));
                                    if (DBG && att_dbg != nullptr && 8 * w + 2 * i + 1 < cnt) att_dbg[fdbg + 8 * w + 2 * i + 1] = a;
                                    const unsigned hi = mul_frame<1>(sv[i] >> 16, gain_of_att<true>(a, W.etab));
                                    v[i] = lo | (hi << 16);
                                }
                            }
                            crow[w] = make_uint4(v[0], v[1], v[2], v[3]);
                        }
                    }
                }
                if (DBG && quiet && att_dbg != nullptr)
                    for (int k = 0; k < cnt; ++k) att_dbg[out_off + i0 + k] = a;
                s_same[wid][lane] = (unsigned char)(__double_as_longlong(a) == __double_as_longlong(old_end));
                if (P.mode != 2) bend_b[cb] = a;                    // (mode 2 would store what is there already)
            }
        }
        __syncwarp();
        work = issue_rms(cb + 1, on);                               // the next block's RMS rows (into the consumed halves of the M rows)
        issue_smp(cb + 1, on);                                      // ... and samples (W.smp is consumed)
        if (NB > 1) __syncthreads();                                // every band's compressed rows are in place
        // ---- overlay + store, row-wise, rows dealt round-robin to the CTA's warps -----------------------------------
        if (CH == 2) {
            // stereo: four frames (one 128-bit word) per lane, eight lanes per row, four rows per instruction
#pragma unroll
            for (int it = wid; it < 8; it += NB) {
                const int q = 4 * it + (lane >> 3), part = lane & 7;
                const ulonglong2 d = W.row[q];
                const int left = (int)d.y - 4 * part;                    // frames of the row from this lane's first one on
                if (left > 0) {
                    uint4 v = reinterpret_cast<const uint4 *>(WS[0].m[q])[part];
#pragma unroll
                    for (int b = 1; b < NB; ++b) {
                        const uint4 o = reinterpret_cast<const uint4 *>(WS[b].m[q])[part];
                        v.x = add_frame_sat<2>(v.x, o.x); v.y = add_frame_sat<2>(v.y, o.y);
                        v.z = add_frame_sat<2>(v.z, o.z); v.w = add_frame_sat<2>(v.w, o.w);
                    }
                    unsigned *g = reinterpret_cast<unsigned *>(proc) + ((int64_t)d.x + 4 * part);
                    if (left >= 4 && (reinterpret_cast<unsigned long long>(g) & 15ull) == 0) *reinterpret_cast<uint4 *>(g) = v;
                    else {
                        g[0] = v.x;
                        if (left > 1) g[1] = v.y;
                        if (left > 2) g[2] = v.z;
                        if (left > 3) g[3] = v.w;
                    }
                }
            }
        } else {
#pragma unroll 2
            for (int q = wid; q < 32; q += NB) {
                const ulonglong2 d = W.row[q];
                if (lane < (int)d.y) {
                    const int64_t f = (int64_t)d.x + lane;
                    unsigned v = reinterpret_cast<const uint16_t *>(WS[0].m[q])[lane];
#pragma unroll
                    for (int b = 1; b < NB; ++b) v = add_frame_sat<1>(v, reinterpret_cast<const uint16_t *>(WS[b].m[q])[lane]);
                    proc[f] = (int16_t)v;
                }
            }
        }
        if (P.mode == 1 && on) {
            bool same = true;
#pragma unroll
            for (int b = 0; b < NB; ++b) same = same && s_same[b][lane] != 0;
            if (same) merged = true;
        }
        ++cb;
        if (NB > 1) __syncthreads();                                // rows and flags consumed: the next block may overwrite them
        else __syncwarp();
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (live) {
        if (P.mode == 0) {
            ss_out[slot] = a_start;
            se_out[slot] = a;
        } else if (P.mode == 1) {
            ss_out[slot] = a_in;
            se_out[slot] = merged ? se_in[slot] : a;
        }
    }
}

// k_comp_dirty: ahead of a repair round, one thread per (stream, tile): a tile in which every band's assumed start
// state is its predecessor's current end state (bit for bit) is carried over to the round's output unchanged; the
// others are appended to `list` (any order), so that k_comp's repair pass runs on dense warps of dirty tiles and a
// round with nothing to repair costs two empty launches.
template <int NB>
__global__ void __launch_bounds__(256)
k_comp_dirty(const StreamDesc *__restrict__ streams, const PlanDev *__restrict__ plans, RecurParams P,
             const double *__restrict__ ss_in, const double *__restrict__ se_in, double *__restrict__ ss_out, double *__restrict__ se_out,
             unsigned *__restrict__ list, unsigned *__restrict__ count, unsigned long long *__restrict__ counters)
{
    const int gl = blockIdx.x * blockDim.x + threadIdx.x;
    if (gl >= P.n_streams * P.tiles) return;
    const int s = gl / P.tiles, tile = gl % P.tiles;
    const StreamDesc sd = streams[s];
    if (!plans[sd.plan].multiband || tile * P.tile_len >= sd.out_frames) return;
    const size_t slot0 = ((size_t)s * NB) * P.tiles + tile;
    bool dirty = false;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        const size_t sl = slot0 + (size_t)b * P.tiles;
        dirty = dirty || (tile != 0 && __double_as_longlong(ss_in[sl]) != __double_as_longlong(se_in[sl - 1]));
    }
    if (dirty) {
        list[atomicAdd(count, 1u)] = (unsigned)gl;
        atomicAdd(&counters[2], 1ull);
    } else {
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const size_t sl = slot0 + (size_t)b * P.tiles;
            ss_out[sl] = ss_in[sl]; se_out[sl] = se_in[sl];
        }
    }
}

// k_comp_sprint: the TRUE state at every tile joint, ahead of the (single) repair round.  A Jacobi round carries
// the truth one tile further and costs a launch whose length is its slowest lane's -- with gains, samples and overlay,
// ~40 ns per frame; where trajectories from different pasts stay apart for long (pydub releases at a rate
// proportional to the CURRENT maximum attenuation: after a loud passage a quiet one just above the threshold lets the
// state drift for seconds without touching a clamp) that was one round per tile of the stretch.  The state alone
// is much cheaper to carry: one warp per (stream, band) walks the stream's tiles in order; a tile whose assumed start
// is the truth is skipped (its stored end is the truth), any other is walked -- recurrence only, 32 frames per
// step of the warp: every lane loads one frame's RMS, gathers its curve value and divides, the values go round by
// shuffle, the loads of the next block are in flight meanwhile -- until the state meets the stored block-end state of
// the speculative pass (then the stored end holds) or the tile ends.  `se_true` receives the end state of every
// tile; k_comp_dirty / k_comp (mode 1) with se_in = se_true then re-run exactly the tiles whose guess was wrong, each
// from its true state, all in ONE round.
template <int NB>
__global__ void __launch_bounds__(128)
k_comp_sprint(const StreamDesc *__restrict__ streams, const PlanDev *__restrict__ plans, RecurParams P, BandPtrs bp,
              const double *__restrict__ ss, const double *__restrict__ se, double *__restrict__ ss_true, double *__restrict__ se_true,
              unsigned *__restrict__ fine_list, unsigned *__restrict__ fine_count, unsigned *__restrict__ fine_bitmap, int fine_tiles,
              unsigned long long *__restrict__ counters)
{
    const int gw = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    const int s = gw / NB, b = gw % NB;
    if (s >= P.n_streams) return;
    const StreamDesc sd = streams[s];
    const PlanDev *__restrict__ pl = plans + sd.plan;
    if (!pl->multiband || sd.out_frames <= 0) return;
    const int band = NB == 3 ? b : P.band_base;
    const int ntiles = (sd.out_frames + P.tile_len - 1) / P.tile_len;
    const size_t base = ((size_t)s * NB + b) * P.tiles;
    ss += base; se += base; ss_true += base; se_true += base;
    bool broken = false;
    for (int t = lane; t < ntiles; t += 32) {
        ss_true[t] = ss[t];
        se_true[t] = se[t];
        broken |= t > 0 && __double_as_longlong(ss[t]) != __double_as_longlong(se[t - 1]);
    }
    if (!__any_sync(FULL, broken)) return;
    __syncwarp();
    const uint16_t *__restrict__ rms = bp.rms[band] + sd.out_off;
    const uint32_t *__restrict__ hold = bp.hold[band] + sd.blk_off;
    double *bend = bp.bend[band] + (size_t)sd.blk_off * 32;
    const double *__restrict__ curve = pl->curve[band];
    const BandDev bd = pl->band[band];
    const bool exact = bd.div_trick != 0;
    const double A = bd.attack_frames, R = bd.release_frames, rA = bd.r_attack, rR = bd.r_release;
    // one frame per lane of block `blk` of a tile ending at frame i1: its curve value and the two quotients, and
    // (every lane the same) the stored state at the end of the block
    auto fetch = [&](int blk, int i1, double &M, double &inc, double &dec, double &old_end) {
        const int f = (blk << 5) + lane;
        const bool held = (hold[blk >> 5] >> (blk & 31)) & 1u;
        const unsigned r = (!held && f < i1) ? rms[f] : 0u;
        old_end = bend[blk];
        M = r ? __ldg(curve + r) : 0.0;
        inc = div_const(M, A, rA, exact);
        dec = div_const(M, R, rR, exact);
    };
    __shared__ double s_step[4][3][32];           // the block's M, M / A, M / R, read back as broadcasts by the serial steps
    double (*sw)[32] = s_step[threadIdx.x >> 5];
    double truth = se[0];                         // tile 0 starts from att = 0 exactly
    for (int t = 1; t < ntiles; ++t) {
        if (__double_as_longlong(ss[t]) == __double_as_longlong(truth)) { truth = se[t]; continue; }   // warp-uniform
        const int i0 = t * P.tile_len, i1 = min(i0 + P.tile_len, sd.out_frames);
        const int b0 = i0 >> 5, b1 = (i1 + 31) >> 5;
        double a = truth;
        bool merged = false;
        double M, inc, dec, old_end;
        fetch(b0, i1, M, inc, dec, old_end);
        int last = b0;                                // the last block whose trajectory differs from the stored one
        for (int blk = b0; blk < b1; ++blk) {
            last = blk;
            const bool any = __any_sync(FULL, __double_as_longlong(M) != 0ll);      // an all-zero block (held) is 32 identity steps
            if (any) { sw[0][lane] = M; sw[1][lane] = inc; sw[2][lane] = dec; }
            const double cmp = old_end;
            __syncwarp();
            if (blk + 1 < b1) fetch(blk + 1, i1, M, inc, dec, old_end);             // in flight during the steps
            if (any) {
                if (exact) {
#pragma unroll 8
                    for (int k = 0; k < 32; ++k) a = recur_step_pos(a, sw[0][k], sw[1][k], sw[2][k]);
                } else {
#pragma unroll 4
                    for (int k = 0; k < 32; ++k) a = recur_step(a, sw[0][k], sw[1][k], sw[2][k]);
                }
            }
            __syncwarp();                                                           // the steps have read the block's values
            if (__double_as_longlong(a) == __double_as_longlong(cmp)) { merged = true; break; }
            if (lane == 0) bend[blk] = a;                                           // the stored trajectory becomes the true one
        }
        // the samples of blocks b0 .. last are to be recomputed: mark the 1024-frame pieces that cover them
        // (once each: the three bands of a stream mark the same pieces)
        for (int ft = (b0 >> 5) + lane; ft <= (last >> 5); ft += 32) {
            const unsigned gidx = (unsigned)s * (unsigned)fine_tiles + (unsigned)ft, bit = 1u << (gidx & 31u);
            if (!(atomicOr(&fine_bitmap[gidx >> 5], bit) & bit)) fine_list[atomicAdd(fine_count, 1u)] = gidx;
        }
        if (lane == 0) { ss_true[t] = truth; atomicAdd(&counters[2], 1ull); }
        truth = merged ? se[t] : a;
        if (lane == 0) se_true[t] = truth;
    }
}

// counters[0] = tiles repaired sequentially, counters[1] = frames re-run sequentially,
// counters[2] = tiles repaired in the parallel rounds
// One warp per stream: all lanes check the tile joints of every band in parallel (assumed start ==
// predecessor's end, bit for bit); only a stream with a broken joint is walked, tile by tile in order:
// lane b carries band b's true state, a broken tile is re-run block by block (lanes 0 .. NB-1 step
// their band through the block, then all 32 lanes apply the gains, one frame each) until every band
// meets its stored block-end state.
template <int CH, int NB>
__global__ void __launch_bounds__(128)
k_comp_fix(const StreamDesc *__restrict__ streams, const PlanDev *__restrict__ plans, RecurParams P, BandPtrs bp,
           int16_t *__restrict__ proc, const double *__restrict__ spec_start, const double *__restrict__ spec_end,
           unsigned long long *__restrict__ counters)
{
    __shared__ double s_att[4][NB][32];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int s = blockIdx.x * 4 + wid;
    if (s >= P.n_streams) return;
    const StreamDesc sd = streams[s];
    const PlanDev *__restrict__ pl = plans + sd.plan;
    if (!pl->multiband || sd.out_frames <= 0) return;
    const int ntiles = (sd.out_frames + P.tile_len - 1) / P.tile_len;
    bool broken = false;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        const double *ss = spec_start + ((size_t)s * NB + b) * P.tiles, *se = spec_end + ((size_t)s * NB + b) * P.tiles;
        for (int t = 1 + lane; t < ntiles; t += 32)
            broken |= __double_as_longlong(ss[t]) != __double_as_longlong(se[t - 1]);
    }
    if (!__any_sync(FULL, broken)) return;
    // lane b < NB: band b
    const int myb = lane < NB ? lane : 0, band = NB == 3 ? myb : P.band_base;
    const double *ss = spec_start + ((size_t)s * NB + myb) * P.tiles, *se = spec_end + ((size_t)s * NB + myb) * P.tiles;
    const uint16_t *__restrict__ rms = bp.rms[band] + sd.out_off;
    double *__restrict__ bend = bp.bend[band] + (size_t)sd.blk_off * 32;
    const double *__restrict__ curve = pl->curve[band];
    const BandDev bd = pl->band[band];
    const bool exact = bd.div_trick != 0;
    double truth = se[0];                        // tile 0 starts from att = 0 exactly
    for (int t = 1; t < ntiles; ++t) {
        const bool ok = lane >= NB || __double_as_longlong(ss[t]) == __double_as_longlong(truth);
        if (__all_sync(FULL, ok)) { truth = se[t]; continue; }
        if (lane == 0) atomicAdd(&counters[0], 1ull);
        const int i0 = t * P.tile_len, i1 = min(i0 + P.tile_len, sd.out_frames);
        double a = truth;
        bool merged = false;
        int done = 0;
        for (int blk = i0 >> 5; (blk << 5) < i1 && !merged; ++blk) {
            const int f0 = blk << 5, cnt = min(32, i1 - f0);
            if (lane < NB) {
                for (int k = 0; k < cnt; ++k) {
                    const unsigned r = rms[f0 + k];
                    const double M = r ? curve[r] : 0.0;
                    a = recur_step(a, M, div_const(M, bd.attack_frames, bd.r_attack, exact), div_const(M, bd.release_frames, bd.r_release, exact));
                    s_att[wid][myb][k] = a;
                }
            }
            __syncwarp();
            if (lane < cnt) {
                const int64_t f = sd.out_off + f0 + lane;
                unsigned acc = 0u;
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    const double at = s_att[wid][b][lane];
                    const int bb = NB == 3 ? b : P.band_base;
                    const unsigned v = mul_frame<CH>(load_frame<CH>(bp.band[bb], f), gain_of_att<true>(at, g_exp10_tab));
                    if (bp.att[bb] != nullptr) bp.att[bb][f] = at;
                    acc = b == 0 ? v : add_frame_sat<CH>(acc, v);
                }
                store_frame<CH>(proc, f, acc);
            }
            bool same = true;
            if (lane < NB) {
                same = __double_as_longlong(a) == __double_as_longlong(bend[blk]);
                bend[blk] = a;
            }
            merged = __all_sync(FULL, same);
            done += cnt;
            __syncwarp();
        }
        if (lane == 0) atomicAdd(&counters[1], (unsigned long long)done);
        truth = merged ? se[t] : a;
    }
}
// =====================================================================================
// k_kweight: pyloudnorm K-weighting of the (L+R)/2 mean of the processed track
// (ENG:214-218): shelf then high-pass, float64 DF2T (scipy lfilter), each stage stored
// back to float32.  One CTA per track, sequential over 4096-sample tiles; the filter
// state runs through the WHOLE track (loudness is measured after chunk concatenation).
// IN = int16 (proc, CH interleaved) or float (mono helper entry point, CH must be 1).
// =====================================================================================
constexpr int KNT = 256;
constexpr int KTILE = SEG * KNT;               // 4096
constexpr int KTILE_PAD = KTILE + KNT;

#ifndef B200M_KW_OCC
#define B200M_KW_OCC 4
#endif
// PT: every plan of the launch has the same K-weighting filters (they depend on the rate alone) and their
// lane-independent tables arrive in `kt` (constant bank).
template <int CH, typename IN, bool PT>
__global__ void __launch_bounds__(KNT, B200M_KW_OCC)
k_kweight(const IN *__restrict__ src_all, const TrackDesc *__restrict__ tracks, const SegDesc *__restrict__ segs,
          const PlanDev *__restrict__ plans, float *__restrict__ kw, const __grid_constant__ KwTabsC kt)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SecTab *tabs = reinterpret_cast<SecTab *>(smem_raw);                    // kw[2]
    float *sx = reinterpret_cast<float *>(smem_raw + 2 * sizeof(SecTab));   // [KTILE_PAD]
    double *carry = reinterpret_cast<double *>(sx + KTILE_PAD + (KTILE_PAD & 1)); // [2][2]
    double *wtot = carry + 4;                                               // [2][8][2]
    const SegDesc sg = segs[blockIdx.x];
    const TrackDesc td = tracks[sg.owner];
    const PlanDev *__restrict__ pl = plans + td.plan;
    if (!pl->has_lufs) return;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    {
        const double *s = reinterpret_cast<const double *>(pl->kw);
        double *d = reinterpret_cast<double *>(tabs);
        for (int i = tid; i < 2 * (int)(sizeof(SecTab) / 8); i += KNT) d[i] = s[i];
        if (tid < 4) carry[tid] = 0.0;
    }
    __syncthreads();
    const IN *__restrict__ src = src_all + td.off * CH;
    float *__restrict__ dst = kw + td.off;
    float *myx = sx + tid * (SEG + 1);
    unsigned round = 0;
    // Whole tiles inside the track come in with cp.async (16-byte pieces, BPF per thread) while the tile before
    // is filtered; every thread converts exactly the pieces it fetched, so cp.async.wait_group is the only
    // synchronisation the staging needs.  The last, partial tile (and unaligned sources) are read directly.
    constexpr int BPF = (int)sizeof(IN) * CH, FPP = 16 / BPF;       // bytes per frame, frames per piece
    unsigned char *sraw = reinterpret_cast<unsigned char *>(wtot + 2 * 8 * 2);     // [KTILE * BPF], 16-byte aligned
    const bool al = (reinterpret_cast<unsigned long long>(src) & 15ull) == 0;
    auto staged = [&](int64_t t0) { return al && t0 + KTILE <= td.frames; };
    auto fetch = [&](int64_t t0) {
        if (staged(t0)) {
            const unsigned char *g = reinterpret_cast<const unsigned char *>(src) + t0 * BPF;
#pragma unroll
            for (int k = 0; k < BPF; ++k) {
                const int p = tid + k * KNT;
                const unsigned sa = (unsigned)__cvta_generic_to_shared(sraw + 16 * p);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(g + 16 * p) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    const int64_t t_first = max((int64_t)0, sg.begin - sg.warm);
    fetch(t_first);
    for (int64_t t0 = t_first; t0 < sg.end; t0 += KTILE) {
        const bool store = t0 >= sg.begin;           // warm-up tiles only advance the filter states
        const int nload = (int)min((int64_t)KTILE, td.frames - t0);
        const int nvalid = store ? (int)min((int64_t)KTILE, sg.end - t0) : 0;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        if (staged(t0)) {
#pragma unroll
            for (int k = 0; k < BPF; ++k) {
                const int p = tid + k * KNT, f0 = p * FPP;
                const uint4 w = *reinterpret_cast<const uint4 *>(sraw + 16 * p);
                const unsigned ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (sizeof(IN) == 2 && CH == 2) {
                        // float32 (vL + vR) / 2 with vL, vR multiples of 2^-15: exact
                        sx[pidx(f0 + i)] = (float)((int)(short)(ww[i] & 0xffffu) + ((int)ww[i] >> 16)) * (1.0f / 65536.0f);
                    } else if (sizeof(IN) == 2) {
                        sx[pidx(f0 + 2 * i)] = (float)(int)(short)(ww[i] & 0xffffu) * (1.0f / 32768.0f);
                        sx[pidx(f0 + 2 * i + 1)] = (float)((int)ww[i] >> 16) * (1.0f / 32768.0f);
                    } else {
                        sx[pidx(f0 + i)] = __uint_as_float(ww[i]);
                    }
                }
            }
        } else
        for (int f = tid; f < KTILE; f += KNT) {
            float m = 0.0f;
            if (f < nload) {
                if (sizeof(IN) == 2) {
                    if (CH == 2) {
                        const short2 q = *reinterpret_cast<const short2 *>(
                            reinterpret_cast<const int16_t *>(src) + (t0 + f) * 2);
                        // float32 (vL + vR) / 2 with vL, vR multiples of 2^-15: exact
                        m = (float)((int)q.x + (int)q.y) * (1.0f / 65536.0f);
                    } else {
                        m = (float)reinterpret_cast<const int16_t *>(src)[t0 + f] * (1.0f / 32768.0f);
                    }
                } else {
                    m = (float)src[t0 + f];
                }
            }
            sx[pidx(f)] = m;
        }
        if (t0 + KTILE < sg.end) fetch(t0 + KTILE);  // this thread's pieces of sraw are consumed: the next tile may land
        __syncthreads();
        double x[SEG];
#pragma unroll
        for (int n = 0; n < SEG; ++n) x[n] = (double)myx[n];
        if (PT) section_round<8>(x, kt.sec[0], tabs[0].Q, carry + 0, wtot + (round & 1) * 16, lane, wid, tid == 0);
        else    section_round<8>(x, tabs[0], tabs[0].Q, carry + 0, wtot + (round & 1) * 16, lane, wid, tid == 0);
        ++round;
#pragma unroll
        for (int n = 0; n < SEG; ++n) x[n] = (double)(float)x[n];   // input_data[:,ch] = ... (float32 store)
        if (PT) section_round<8>(x, kt.sec[1], tabs[1].Q, carry + 2, wtot + (round & 1) * 16, lane, wid, tid == 0);
        else    section_round<8>(x, tabs[1], tabs[1].Q, carry + 2, wtot + (round & 1) * 16, lane, wid, tid == 0);
        ++round;
#pragma unroll
        for (int n = 0; n < SEG; ++n) myx[n] = (float)x[n];
        __syncthreads();
        for (int f = tid; f < nvalid; f += KNT) dst[t0 + f] = sx[pidx(f)];
        __syncthreads();
    }
}

constexpr size_t kweight_smem_bytes()
{
    return 2 * sizeof(SecTab) + (size_t)(KTILE_PAD + (KTILE_PAD & 1)) * 4 + 4 * 8 + 2 * 8 * 2 * 8 + (size_t)KTILE * 4 /* raw tile */;
}

// =====================================================================================
// k_kweightw: the same K-weighting for a large batch, one WARP per segment of a track
// (the shape of k_chainw: one warp per CTA, provably uniform control flow, no CTA barrier
// in the loop).  A warp tile is 512 frames -- lane j owns frames 16 j .. 16 j + 15 of the
// mono mean -- the processed PCM of the next tile arrives with cp.async while this one
// is filtered, the filter tables sit in the constant bank (every plan of a batch has the
// same K-weighting: it depends on the rate alone) and each lane stores its 16 results
// as four 16-byte words.  k_kweight (eight warps, a CTA-wide scan and three barriers per
// 4096 frames) ran at 46 % of the fp64 pipe and half the DRAM rate; small batches, which
// cannot give 148 x 16 warps a run of four warm-ups each, stay with it.
// =====================================================================================
template <int CH> struct KwW {
    static constexpr int WT = 32 * SEG;                     // frames per warp tile
    static constexpr int RSTRIDE = CH == 2 ? 20 : 12;       // 32-bit words per lane's 16 raw frames (16 / 8 used): conflict-free LDS.128
    static constexpr int RAW_WORDS = 32 * RSTRIDE;
    static constexpr int DEPTH = 3;                         // raw tiles in flight or in use
    static constexpr size_t SMEM = 2 * 32 * 4 * 8 /* Q[lane] of the two sections */ + (size_t)DEPTH * RAW_WORDS * 4 + 4 * 8 /* carry */;
};
#ifndef B200M_KWW_OCC
#define B200M_KWW_OCC 16           // resident warps per SM the kernel is built for (<= 128 registers)
#endif
template <int CH>
__global__ void __launch_bounds__(32, B200M_KWW_OCC)
k_kweightw(const int16_t *__restrict__ src_all, const TrackDesc *__restrict__ tracks, const SegDesc *__restrict__ segs,
           const PlanDev *__restrict__ plans, float *__restrict__ kw, const __grid_constant__ KwTabsC kt)
{
    using W = KwW<CH>;
    constexpr int WT = W::WT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double (*sQ)[32][4] = reinterpret_cast<double (*)[32][4]>(smem_raw);
    unsigned *raw_all = reinterpret_cast<unsigned *>(smem_raw + 2 * 32 * 4 * 8);         // [DEPTH][RAW_WORDS]
    double *carry = reinterpret_cast<double *>(smem_raw + 2 * 32 * 4 * 8 + (size_t)W::DEPTH * W::RAW_WORDS * 4);
    const int lane = threadIdx.x;
    const SegDesc sg = segs[blockIdx.x];
    const TrackDesc td = tracks[sg.owner];
    const PlanDev *__restrict__ pl = plans + td.plan;
    if (!pl->has_lufs || sg.begin >= sg.end) return;                 // CTA-uniform
    for (int i = lane; i < 2 * 128; i += 32) reinterpret_cast<double *>(sQ)[i] = reinterpret_cast<const double *>(pl->kw[i >> 7].Q)[i & 127];
    if (lane < 4) carry[lane] = 0.0;
    __syncwarp();
    const int16_t *__restrict__ src = src_all + td.off * CH;
    float *__restrict__ dst = kw + td.off;
    const bool in16 = (reinterpret_cast<unsigned long long>(src) & 15ull) == 0;
    const int64_t frames = td.frames;

    // processed PCM of tile t0 -> raw[]: 16-byte pieces (4 stereo / 8 mono frames each), zeros past the end of the track
    auto fetch = [&](int64_t t0, int slot) {
        unsigned *raw = raw_all + slot * W::RAW_WORDS;
        constexpr int FPP = 8 / CH, PPL = 16 / FPP;                   // frames per piece, pieces per lane's 16 frames
        if (in16 && t0 + WT <= frames) {
            // the whole tile lies inside the track (warp-uniform; all but a track's last tile): piece lane + 32 h sits
            // 512 h bytes behind piece `lane` in global memory and 8 h / PPL ... rows further in raw[] -- constant offsets
            const char *gp = reinterpret_cast<const char *>(src + t0 * CH) + 16 * lane;
            const unsigned sa = (unsigned)__cvta_generic_to_shared(raw + W::RSTRIDE * (lane / PPL) + 4 * (lane % PPL));
#pragma unroll
            for (int h = 0; h < PPL; ++h)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa + h * (W::RSTRIDE * (32 / PPL) * 4)), "l"(gp + 512 * h) : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
            return;
        }
#pragma unroll 1
        for (int h = 0; h < PPL; ++h) {
            const int p = lane + 32 * h;
            const int64_t gf = t0 + (int64_t)FPP * p;
            unsigned *d = raw + W::RSTRIDE * (p / PPL) + 4 * (p % PPL);
            if (in16 && gf + FPP <= frames) {
                const unsigned sa = (unsigned)__cvta_generic_to_shared(d);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(src + gf * CH) : "memory");
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (CH == 2) {
                        d[k] = gf + k < frames ? reinterpret_cast<const unsigned *>(src)[gf + k] : 0u;
                    } else {
                        const unsigned lo = gf + 2 * k < frames ? (unsigned)(unsigned short)src[gf + 2 * k] : 0u;
                        const unsigned hi = gf + 2 * k + 1 < frames ? (unsigned)(unsigned short)src[gf + 2 * k + 1] : 0u;
                        d[k] = lo | (hi << 16);
                    }
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    const int64_t t_first = max((int64_t)0, sg.begin - sg.warm);
    // two tiles ahead: with 12 warps per SM one tile in flight per warp leaves the DRAM pipe half empty
    fetch(t_first, 0);
    fetch(t_first + WT, 1);
    int slot = 0;
    for (int64_t t0 = t_first; t0 < sg.end; t0 += WT) {
        const bool store = t0 >= sg.begin;                           // warm-up tiles only advance the filter states
        const int nvalid = store ? (int)min((int64_t)WT, sg.end - t0) : 0;
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncwarp();
        double x[SEG];
        {
            const uint4 *rp = reinterpret_cast<const uint4 *>(raw_all + slot * W::RAW_WORDS + W::RSTRIDE * lane);
            if (CH == 2) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint4 w = rp[i];
                    const unsigned ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll      // float32 (vL + vR) / 2 with vL, vR multiples of 2^-15: exact (ENG:214)
                    for (int k = 0; k < 4; ++k) x[4 * i + k] = (double)((float)(prmt_sx(ww[k], 0x9910u) + prmt_sx(ww[k], 0xbb32u)) * (1.0f / 65536.0f));
                }
            } else {
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const uint4 w = rp[i];
                    const unsigned ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        x[8 * i + 2 * k] = (double)((float)prmt_sx(ww[k], 0x9910u) * (1.0f / 32768.0f));
                        x[8 * i + 2 * k + 1] = (double)((float)prmt_sx(ww[k], 0xbb32u) * (1.0f / 32768.0f));
                    }
                }
            }
        }
        // the tile two ahead goes into the third slot (a group is committed every round, so wait_group 1 always
        // means "the next tile to be read has landed"; past the track's end a fetch is a zero fill)
        fetch(t0 + 2 * WT, slot == 0 ? 2 : slot - 1);
        slot = slot == 2 ? 0 : slot + 1;
        section_round_w<32>(x, kt.sec[0], sQ[0], carry + 0, lane);
#pragma unroll
        for (int n = 0; n < SEG; ++n) x[n] = (double)(float)x[n];    // input_data[:, ch] = ... (float32 store between the stages)
        section_round_w<32>(x, kt.sec[1], sQ[1], carry + 2, lane);
        if (store) {
            float *o = dst + t0 + SEG * lane;
            const int f0 = SEG * lane;
            if ((reinterpret_cast<unsigned long long>(o) & 15ull) == 0 && f0 + SEG <= nvalid) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    reinterpret_cast<float4 *>(o)[i] = make_float4((float)x[4 * i], (float)x[4 * i + 1], (float)x[4 * i + 2], (float)x[4 * i + 3]);
            } else {
#pragma unroll
                for (int n = 0; n < SEG; ++n) if (f0 + n < nvalid) o[n] = (float)x[n];
            }
        }
    }
}

// =====================================================================================
// numpy's pairwise summation (numpy/_core/src/umath/loops_utils.h.src, @TYPE@_pairwise_sum),
// reproduced operation for operation so block energies and gated means round exactly
// like np.sum / np.mean do inside pyloudnorm.  `get(i)` yields element i.
// =====================================================================================
struct F32 {   // float whose adds can never be contracted into FMAs
    float v;
    __device__ F32() : v(0.f) {}
    __device__ explicit F32(int) : v(0.f) {}
    __device__ F32(float f) : v(f) {}
    __device__ F32 operator+(const F32 &o) const { return F32(__fadd_rn(v, o.v)); }
};

template <typename T, typename Get>
__device__ __forceinline__ T np_pairwise_leaf(Get get, int64_t off, int n)
{
    if (n < 8) {
        T res = (T)0;
        for (int i = 0; i < n; ++i) res = res + get(off + i);
        return res;
    }
    T r[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) r[k] = get(off + k);
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int k = 0; k < 8; ++k) r[k] = r[k] + get(off + i + k);
    }
    T res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res = res + get(off + i);
    return res;
}

// The recursion itself, with the leaves (<= 128 elements) left to `leaf(offset, count)`: the same walk
// enumerates the leaves, sums serially, or combines leaf sums that were computed in parallel.
template <typename T, typename Leaf>
__device__ T np_pairwise_walk(Leaf leaf, int64_t n)
{
    if (n <= 128) return leaf((int64_t)0, (int)n);
    int64_t s_off[40], s_n[40];
    T s_acc[40];
    int s_stage[40];
    int sp = 0;
    s_off[0] = 0; s_n[0] = n; s_stage[0] = 0; s_acc[0] = (T)0;
    T ret = (T)0;
    while (sp >= 0) {
        if (s_n[sp] <= 128) { ret = leaf(s_off[sp], (int)s_n[sp]); --sp; continue; }
        int64_t n2 = s_n[sp] / 2;
        n2 -= n2 % 8;
        if (s_stage[sp] == 0) {
            s_stage[sp] = 1;
            s_off[sp + 1] = s_off[sp]; s_n[sp + 1] = n2; s_stage[sp + 1] = 0; ++sp;
        } else if (s_stage[sp] == 1) {
            s_acc[sp] = ret; s_stage[sp] = 2;
            s_off[sp + 1] = s_off[sp] + n2; s_n[sp + 1] = s_n[sp] - n2; s_stage[sp + 1] = 0; ++sp;
        } else {
            ret = s_acc[sp] + ret; --sp;
        }
    }
    return ret;
}

template <typename T, typename Get>
__device__ T np_pairwise_sum(Get get, int64_t n)
{
    return np_pairwise_walk<T>([&](int64_t off, int cnt) { return np_pairwise_leaf<T>(get, off, cnt); }, n);
}

// =====================================================================================
// k_blocks: 400 ms / 75 %-overlap block mean squares z_j (pyloudnorm meter.py):
//   l = int(T_g*(j*step)*rate), u = int(T_g*(j*step+1)*rate)
//   z_j = float32(1/(T_g*rate)) * np.sum(np.square(y[l:u]))        (float32 throughout)
// np.sum's pairwise order is reproduced exactly AND in parallel: the recursion tree of numpy's
// pairwise_sum depends only on the block length n, so the host tabulates it once per sample
// rate (PlanDev::ptree: leaves of <= 128 elements in order, internal nodes by level).  One CTA
// per block: 8 lanes per leaf play numpy's 8 partial accumulators r[0..7] (a 3-step xor
// butterfly is exactly ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7))), then the internal nodes are added
// level by level.  Blocks of any other length (clamped last block) take the serial walk.
// =====================================================================================
constexpr int BNT = 256;

// Parallel evaluation of one tabulated pairwise tree over y[0 .. T[0]): returns the root (valid in
// every thread).  bval: [nleaves + ninternal] floats of shared memory.  Ends with a barrier.
__device__ __forceinline__ float pw_tree_eval(const float *__restrict__ y, const int32_t *__restrict__ T, float *bval, int tid)
{
    const int nleaves = T[1], nint = T[2], nlev = T[3];
    const int32_t *lvl = T + 4, *loff = lvl + nlev + 1, *ln = loff + nleaves, *nl = ln + nleaves, *nr = nl + nint;
    const int sub = tid >> 3, k = tid & 7;
    for (int L0 = 0; L0 < nleaves; L0 += BNT / 8) {
        const int L = L0 + sub;
        const bool on = L < nleaves;
        const int cnt = on ? ln[L] : 8;
        const float *p = y + (on ? loff[L] : 0);
        const int lim = cnt - (cnt % 8);
        float v0 = p[k];
        float acc = __fmul_rn(v0, v0);
#pragma unroll 8
        for (int i = 8; i < lim; i += 8) {
            const float v = p[i + k];
            acc = __fadd_rn(acc, __fmul_rn(v, v));
        }
        acc = __fadd_rn(acc, __shfl_xor_sync(FULL, acc, 1));
        acc = __fadd_rn(acc, __shfl_xor_sync(FULL, acc, 2));
        acc = __fadd_rn(acc, __shfl_xor_sync(FULL, acc, 4));
        if (on && k == 0) {
            for (int i = lim; i < cnt; ++i) { const float v = p[i]; acc = __fadd_rn(acc, __fmul_rn(v, v)); }
            bval[L] = acc;
        }
    }
    __syncthreads();
    for (int lv = 0; lv < nlev; ++lv) {
        for (int i = lvl[lv] + tid; i < lvl[lv + 1]; i += BNT) bval[nleaves + i] = __fadd_rn(bval[nl[i]], bval[nr[i]]);
        __syncthreads();
    }
    return nint ? bval[nleaves + nint - 1] : bval[0];
}

// k_hops: blocks overlap by 75 %, and at rates whose block length n is a multiple of 32 (48 kHz,
// 96 kHz, ...) numpy's tree over a block splits exactly at n/2 and n/4: sum(block j) =
// (h_j + h_{j+1}) + (h_{j+2} + h_{j+3}) with h the pairwise sum of one 100 ms hop.  The hop sums
// are computed once (one CTA each) into `hops` (floats; slot = hop index - first block of the
// buffer); k_blocks adds four of them whenever pyloudnorm's float bounds of the block are the
// hop multiples (checked per block), and walks the block itself otherwise.
__global__ void __launch_bounds__(BNT)
k_hops(const float *__restrict__ kw, const TrackDesc *__restrict__ tracks, const PlanDev *__restrict__ plans,
       double *__restrict__ hops, int stage_floats)
{
    extern __shared__ float bval[];
    const TrackDesc td = tracks[blockIdx.y];
    const PlanDev *__restrict__ pl = plans + td.plan;
    const int hs = blockIdx.x;
    if (!pl->has_lufs || pl->hop <= 0 || td.nblocks < 3 || hs >= td.nblocks + 3) return;
    const int64_t b = (int64_t)(td.j0 + hs) * pl->hop, e = b + pl->hop;
    if (b < td.abs0 || e > td.total_frames || e - td.abs0 > td.frames) return;     // not (entirely) in this buffer: no block will ask for it
    // The hop's samples come into shared memory as 16-byte cp.async pieces (all of them in flight at once), and the
    // leaves read them there: leaves walk with a stride of eight floats, which from global memory meant 32-byte
    // requests and a dependent load every iteration (56 % of the DRAM rate; staged: the kernel streams).
    const float *src = kw + td.off + (b - td.abs0);
    const int32_t *T = pl->htree;
    float *stage = bval + ((T[1] + T[2] + 3) & ~3);
    const int n = pl->hop;
    if (stage_floats >= n && (reinterpret_cast<unsigned long long>(src) & 15ull) == 0 && (n & 3) == 0) {
        for (int i = threadIdx.x; i < n / 4; i += BNT) {
            const unsigned sa = (unsigned)__cvta_generic_to_shared(stage + 4 * i);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(src + 4 * i) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        src = stage;
    }
    const float v = pw_tree_eval(src, T, bval, threadIdx.x);
    if (threadIdx.x == 0) reinterpret_cast<float *>(hops + td.zoff)[hs] = v;
}

__global__ void __launch_bounds__(BNT)
k_blocks(const float *__restrict__ kw, const TrackDesc *__restrict__ tracks,
         const PlanDev *__restrict__ plans, const double *__restrict__ hops, double *__restrict__ z)
{
    extern __shared__ float bval[];              // [nleaves + ninternal]
    const TrackDesc td = tracks[blockIdx.y];
    const PlanDev *__restrict__ pl = plans + td.plan;
    const int jb = blockIdx.x, j = td.j0 + jb, tid = threadIdx.x;      // jb: slot in this buffer's z, j: block of the track
    if (!pl->has_lufs || jb >= td.nblocks) return;
    const double rate = (double)pl->rate, Tg = 0.4, step = 0.25;
    int64_t l = (int64_t)__dmul_rn(__dmul_rn(Tg, __dmul_rn((double)j, step)), rate);
    int64_t u = (int64_t)__dmul_rn(__dmul_rn(Tg, __dadd_rn(__dmul_rn((double)j, step), 1.0)), rate);
    const float scale = (float)(1.0 / (Tg * rate));
    const int hop = pl->hop;
    if (hops != nullptr && hop > 0 && td.nblocks >= 3 && l == (int64_t)j * hop && u == l + 4 * (int64_t)hop &&
        u <= td.total_frames && l >= td.abs0 && u - td.abs0 <= td.frames) {
        if (tid == 0) {
            const float *H = reinterpret_cast<const float *>(hops + td.zoff) + jb;
            z[td.zoff + jb] = (double)__fmul_rn(scale, __fadd_rn(__fadd_rn(H[0], H[1]), __fadd_rn(H[2], H[3])));
        }
        return;
    }
    if (u > td.total_frames) u = td.total_frames;       // numpy slicing clamps
    if (l > u) l = u;
    const float *__restrict__ y = kw + td.off + (l - td.abs0);
    const int64_t n = u - l;
    const int32_t *__restrict__ T = pl->ptree;
    if (T == nullptr || (int64_t)T[0] != n) {
        if (tid == 0) {
            auto get = [y](int64_t i) { const float v = y[i]; return F32(__fmul_rn(v, v)); };
            z[td.zoff + jb] = (double)__fmul_rn(scale, np_pairwise_sum<F32>(get, n).v);
        }
        return;
    }
    const float root = pw_tree_eval(y, T, bval, tid);
    if (tid == 0) z[td.zoff + jb] = (double)__fmul_rn(scale, root);
}

// =====================================================================================
// k_gate: absolute (-70 LUFS) and relative (-10 LU) gating over the block energies,
// integrated loudness and the linear gain of ENG:219-220.  One CTA per track: block-wide ordered
// compaction, gated means in numpy's pairwise order with the leaves summed in parallel (a 2-hour
// track has 72 000 blocks).
// out[track] = {loudness, gain}
// =====================================================================================
constexpr int GNT = 256;            // threads per track in k_gate
constexpr int GLEV = 11;            // the pairwise tree of a gated mean is expanded this deep in parallel
constexpr int GSLOTS = 1 << GLEV;   // 2048 leaf slots: up to 225 k gated blocks (6.2 h of audio), beyond that the serial walk

// np.mean of sel[0 .. cnt) in numpy's pairwise order by the whole CTA.  The recursion (n > 128: n2 = n/2 -
// (n/2) % 8, sum(n2) + sum(n - n2)) is expanded level by level in place: node i of level d lives in slot
// i * 2^(GLEV-d); a split leaves its left child there and puts the right child half a stride further, a node
// that is already a leaf just stays.  Then every slot with elements is summed like numpy's leaf (eight
// accumulators), and the levels are folded back, adding a right child wherever a split created one -- the
// additions of the sequential recursion, each exactly once and with the same operands.
__device__ double gate_mean(const double *__restrict__ sel, int cnt, int *s_off, int *s_n, double *s_val)
{
    const int tid = threadIdx.x;
    auto get = [sel](int64_t i) { return sel[i]; };
    __shared__ double s_res;
    if (cnt <= 0) return __longlong_as_double(0x7ff8000000000000LL);      // np.mean([]) -> nan
    bool serial = cnt <= 128 || cnt > 110 * GSLOTS;        // 110 + 16 (rounding of eleven splits) <= 128
    if (!serial) {
        for (int i = tid; i < GSLOTS; i += GNT) { s_n[i] = 0; s_off[i] = 0; }
        __syncthreads();
        if (tid == 0) s_n[0] = cnt;
        __syncthreads();
        for (int d = 0; d < GLEV; ++d) {
            const int S = GSLOTS >> d;
            for (int p = tid * S; p < GSLOTS; p += GNT * S) {
                const int n = s_n[p];
                if (n > 128) {
                    int n2 = n / 2;
                    n2 -= n2 % 8;
                    const int off = s_off[p];
                    s_n[p] = n2;
                    s_off[p + S / 2] = off + n2;
                    s_n[p + S / 2] = n - n2;
                }
            }
            __syncthreads();
        }
        bool big = false;
        for (int i = tid; i < GSLOTS; i += GNT) {
            const int n = s_n[i];
            big |= n > 128;
            if (n > 0 && n <= 128) s_val[i] = np_pairwise_leaf<double>(get, s_off[i], n);
        }
        serial = __syncthreads_or(big);                                   // deeper than GLEV levels (cannot happen below 110 * GSLOTS)
        if (!serial) {
            for (int d = GLEV - 1; d >= 0; --d) {
                const int S = GSLOTS >> d;
                for (int p = tid * S; p < GSLOTS; p += GNT * S)
                    if (s_n[p + S / 2] > 0) s_val[p] = s_val[p] + s_val[p + S / 2];
                __syncthreads();
            }
            if (tid == 0) s_res = s_val[0] / (double)cnt;
        }
    }
    if (serial && tid == 0) s_res = np_pairwise_sum<double>(get, cnt) / (double)cnt;
    __syncthreads();
    const double r = s_res;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(GNT)
k_gate(const TrackDesc *__restrict__ tracks, const PlanDev *__restrict__ plans,
       const double *__restrict__ z, double *__restrict__ zsel, double2 *__restrict__ out)
{
    __shared__ int s_off[GSLOTS], s_n[GSLOTS], s_wcnt[GNT / 32];
    __shared__ double s_val[GSLOTS];
    const TrackDesc td = tracks[blockIdx.x];
    const PlanDev *__restrict__ pl = plans + td.plan;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (!pl->has_lufs) {
        if (tid == 0) out[blockIdx.x] = make_double2(__longlong_as_double(0x7ff8000000000000LL), 1.0);
        return;
    }
    const double *__restrict__ zt = z + td.zoff;
    double *__restrict__ sel = zsel + td.zoff;
    const int nb = td.nblocks;
    double gamma_r = 0.0;
    double mean = 0.0;
    for (int pass = 0; pass < 2; ++pass) {
        // ordered compaction of the blocks that pass the gate (the order is what np.mean's pairwise tree sees)
        int cnt = 0;
        for (int j0 = 0; j0 < nb; j0 += GNT) {
            const int j = j0 + tid;
            bool keep = false;
            double zj = 0.0;
            if (j < nb) {
                zj = zt[j];
                const double lj = __dadd_rn(-0.691, __dmul_rn(10.0, log10(zj)));
                keep = pass == 0 ? (lj >= -70.0) : (lj > gamma_r && lj > -70.0);
            }
            const unsigned m = __ballot_sync(FULL, keep);
            if (lane == 0) s_wcnt[wid] = __popc(m);
            __syncthreads();
            int off = cnt, tot = 0;
#pragma unroll
            for (int w = 0; w < GNT / 32; ++w) { if (w < wid) off += s_wcnt[w]; tot += s_wcnt[w]; }
            if (keep) sel[off + __popc(m & ((1u << lane) - 1u))] = zj;
            cnt += tot;
            __syncthreads();
        }
        mean = gate_mean(sel, cnt, s_off, s_n, s_val);
        if (pass == 0) gamma_r = __dsub_rn(__dadd_rn(-0.691, __dmul_rn(10.0, log10(mean))), 10.0);
    }
    if (tid == 0) {
        if (mean != mean) mean = 0.0;           // np.nan_to_num
        const double lufs = __dadd_rn(-0.691, __dmul_rn(10.0, log10(mean)));
        const double gain = pow(10.0, (pl->lufs - lufs) / 20.0);
        out[blockIdx.x] = make_double2(lufs, gain);
    }
}

// k_wav_headers: ENG:96-99 `export(format="wav")` folded into the batch.  pydub hands the samples to the
// stdlib wave module, whose file is the canonical 44-byte RIFF/WAVE header (PCM, 16 bit) followed by the
// data; the header of every track is written immediately ahead of its samples in the output buffer, so a
// track's file image is one contiguous span that the host writes out (or uploads) as it is.
__global__ void k_wav_headers(const TrackDesc *__restrict__ tracks, const PlanDev *__restrict__ plans, int n_tracks, int ch,
                              int16_t *__restrict__ out)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tracks) return;
    const TrackDesc td = tracks[t];
    const unsigned rate = (unsigned)plans[td.plan].rate, bytes = (unsigned)(td.frames * ch * 2);
    unsigned *hd = reinterpret_cast<unsigned *>(out + td.dst_off * ch) - 11;      // 44 bytes ahead of the first sample (4-byte aligned)
    hd[0] = 0x46464952u;                         // "RIFF"
    hd[1] = 36u + bytes;
    hd[2] = 0x45564157u;                         // "WAVE"
    hd[3] = 0x20746d66u;                         // "fmt "
    hd[4] = 16u;
    hd[5] = 1u | ((unsigned)ch << 16);           // WAVE_FORMAT_PCM, channels
    hd[6] = rate;
    hd[7] = rate * (unsigned)ch * 2u;            // bytes per second
    hd[8] = ((unsigned)ch * 2u) | (16u << 16);   // block align, bits per sample
    hd[9] = 0x61746164u;                         // "data"
    hd[10] = bytes;
}

// k_regain: the gain of ENG:219-220 for ANOTHER loudness target of the same measurement (a sweep over
// targets shares everything ahead of ENG:84): out[t] = {lufs[t], 10^((target - lufs[t]) / 20)}, the
// expression k_gate evaluates for the plan's own target.
__global__ void k_regain(const double2 *__restrict__ loud, int n_tracks, double target, double2 *__restrict__ out)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n_tracks) {
        const double lufs = loud[t].x;
        out[t] = make_double2(lufs, pow(10.0, (target - lufs) / 20.0));
    }
}

// =====================================================================================
// k_final: re-float the processed track (ENG:82), one gain (ENG:222, float64 because the
// loudness is a numpy float64 scalar), rational soft limiter (ENG:224-227), quantise #3
// (ENG:89).  Without a loudness target the limiter runs in float32, as in the reference.
// grid = (tiles, tracks).
// =====================================================================================
__device__ __forceinline__ double limiter64(double x, double thr)
{
    const double ax = fabs(x);
    if (ax > thr) {
        const double d = __dsub_rn(ax, thr);
        const double t = __ddiv_rn(d, 0.02);
        const double r = __ddiv_rn(d, __dsqrt_rn(__dadd_rn(1.0, __dmul_rn(t, t))));
        const double sgn = x > 0.0 ? 1.0 : (x < 0.0 ? -1.0 : x);   // np.sign
        return __dmul_rn(__dadd_rn(thr, r), sgn);
    }
    return x;
}

__device__ __forceinline__ float limiter32(float x, float thr)
{
    const float ax = fabsf(x);
    if (ax > thr) {
        const float d = __fsub_rn(ax, thr);
        const float t = __fdiv_rn(d, 0.02f);
        const float r = __fdiv_rn(d, __fsqrt_rn(__fadd_rn(1.0f, __fmul_rn(t, t))));
        const float sgn = x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : x);
        return __fmul_rn(__fadd_rn(thr, r), sgn);
    }
    return x;
}

// One sample of ENG:82-89.  v = q / 2^15 is exact in float32 and scaling by a power of two commutes
// with rounding, so fl(v * gain) * 2^15 == fl(q * gain): below the limiter threshold the whole
// re-float / gain / quantise sequence is one multiply and one truncation, and the limiter test
// |fl(v * gain)| > 0.98 is |fl(q * gain)| > 0.98 * 2^15 exactly.  Anything else (limited samples,
// NaN from 0 * inf on silent input) takes the reference's formula step by step.
__device__ __forceinline__ int final_sample(int q, bool has, double gain)
{
    if (has) {
        const double p = __dmul_rn((double)q, gain);
        if (fabs(p) <= 0.98 * 32768.0) return __double2int_rz(p);
        const float v = (float)q * (1.0f / 32768.0f);
        return quant16(limiter64(__dmul_rn((double)v, gain), 0.98));
    }
    // float32 limiter (no loudness target): |q / 2^15| > 0.98f  <=>  |q| >= 32113, and below it the
    // quantiser returns q itself
    if (abs(q) <= 32112) return q;
    return quant16((double)limiter32((float)q * (1.0f / 32768.0f), 0.98f));
}

// eight samples through final_sample: the path of a vector in which some sample meets the limiter (one copy of the code)
__device__ __noinline__ uint4 final_vector(uint4 w, bool has, double gain)
{
    const unsigned in[4] = {w.x, w.y, w.z, w.w};
    unsigned o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int r0 = final_sample(prmt_sx(in[k], 0x9910u), has, gain);
        const int r1 = final_sample(prmt_sx(in[k], 0xbb32u), has, gain);
        o[k] = (unsigned)(r0 & 0xffff) | ((unsigned)r1 << 16);
    }
    return make_uint4(o[0], o[1], o[2], o[3]);
}

template <int CH>
__global__ void __launch_bounds__(256)
k_final(const int16_t *__restrict__ proc, const TrackDesc *__restrict__ tracks,
        const PlanDev *__restrict__ plans, const double2 *__restrict__ loud,
        int16_t *__restrict__ out)
{
    const TrackDesc td = tracks[blockIdx.y];
    const PlanDev *__restrict__ pl = plans + td.plan;
    const bool has = pl->has_lufs != 0;
    const double gain = has ? loud[blockIdx.y].y : 1.0;
    const int16_t *__restrict__ src = proc + td.off * CH;
    int16_t *__restrict__ dst = out + td.dst_off * CH;
    const int64_t nsamp = td.frames * CH;
    // eight samples per thread and step (16-byte accesses) where both ends are aligned; workspace track
    // starts always are, packed output starts are when the preceding tracks have multiples of 8 / CH frames
    const bool vec = ((reinterpret_cast<unsigned long long>(src) | reinterpret_cast<unsigned long long>(dst)) & 15ull) == 0;
    const int64_t nvec = vec ? nsamp / 8 : 0;
    const int64_t stride = (int64_t)gridDim.x * 256;
    for (int64_t g = (int64_t)blockIdx.x * 256 + threadIdx.x; g < nvec; g += 2 * stride) {
        // two vectors per thread and round, both loads in flight before the first sample is touched
        const int64_t g2 = g + stride;
        const bool second = g2 < nvec;
        uint4 w[2];
        w[0] = __ldg(reinterpret_cast<const uint4 *>(src) + g);
        w[1] = second ? __ldg(reinterpret_cast<const uint4 *>(src) + g2) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int v = 0; v < 2; ++v) {
            if (v == 1 && !second) break;
            const unsigned in[4] = {w[v].x, w[v].y, w[v].z, w[v].w};
            unsigned o[4];
            // the eight samples of a vector take the common case without a branch (one multiply, one truncation: see
            // final_sample; without a loudness target the sample itself), the "needs the limiter" tests accumulate
            // in one predicate, and only a vector with such a sample (or a NaN) goes back through final_sample.  A branch
            // per sample kept the eight conversion / multiply chains from overlapping.
            bool slow = false;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int q0 = prmt_sx(in[k], 0x9910u), q1 = prmt_sx(in[k], 0xbb32u);
                if (has) {
                    const double p0 = __dmul_rn((double)q0, gain), p1 = __dmul_rn((double)q1, gain);
                    slow = slow || !(fabs(p0) <= 0.98 * 32768.0) || !(fabs(p1) <= 0.98 * 32768.0);
                    o[k] = __byte_perm((unsigned)__double2int_rz(p0), (unsigned)__double2int_rz(p1), 0x5410);
                } else {
                    slow = slow || abs(q0) > 32112 || abs(q1) > 32112;
                    o[k] = in[k];
                }
            }
            uint4 ov = make_uint4(o[0], o[1], o[2], o[3]);
            if (slow) ov = final_vector(w[v], has, gain);
            reinterpret_cast<uint4 *>(dst)[v ? g2 : g] = ov;
        }
    }
    for (int64_t i = nvec * 8 + (int64_t)blockIdx.x * 256 + threadIdx.x; i < nsamp; i += (int64_t)gridDim.x * 256)
        dst[i] = (int16_t)final_sample((int)src[i], has, gain);
}

// =====================================================================================
// k_stage_s24 / k_stage_f32: declared extension (the reference handles 16-bit PCM only, ENG:125).
// Packed little-endian 24-bit PCM is reduced to the reference's 16-bit domain exactly like
// pydub's AudioSegment.set_sample_width(2) = audioop.lin2lin (keep the high-order 16 bits);
// float32 PCM goes through the reference's own quantiser (ENG:123-126).  Each thread converts
// four samples: three aligned 32-bit loads for 12 bytes of 24-bit PCM, one 8-byte store.
// =====================================================================================
__global__ void __launch_bounds__(256)
k_stage_s24(const unsigned char *__restrict__ in, int64_t n, int16_t *__restrict__ out)
{
    const bool al = ((reinterpret_cast<unsigned long long>(in) & 3ull) | (reinterpret_cast<unsigned long long>(out) & 7ull)) == 0;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g * 4 < n; g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = g * 4;
        if (al && i + 4 <= n) {
            const unsigned w0 = reinterpret_cast<const unsigned *>(in)[3 * g], w1 = reinterpret_cast<const unsigned *>(in)[3 * g + 1],
                           w2 = reinterpret_cast<const unsigned *>(in)[3 * g + 2];
            // bytes b0..b11; sample k = b[3k+1] | b[3k+2] << 8
            const unsigned s0 = (w0 >> 8) & 0xffffu, s1 = w1 & 0xffffu, s2 = (w1 >> 24) | ((w2 & 0xffu) << 8), s3 = w2 >> 16;
            reinterpret_cast<uint2 *>(out)[g] = make_uint2(s0 | (s1 << 16), s2 | (s3 << 16));
        } else {
            for (int64_t k = i; k < n && k < i + 4; ++k)
                out[k] = (int16_t)((unsigned)in[3 * k + 1] | ((unsigned)in[3 * k + 2] << 8));
        }
    }
}

__global__ void __launch_bounds__(256)
k_stage_f32(const float *__restrict__ in, int64_t n, int16_t *__restrict__ out)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (int16_t)quant16((double)in[i]);
}

// =====================================================================================
// Stage-level helper kernels (back the reference's per-function API; parity tests).
// =====================================================================================
__global__ void k_pcm16_to_float(const int16_t *__restrict__ in, int64_t n, float *__restrict__ out)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (float)in[i] * (1.0f / 32768.0f);
}

template <typename T>
__global__ void k_float_to_pcm16(const T *__restrict__ in, int64_t n, int16_t *__restrict__ out)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (int16_t)quant16((double)in[i]);
}

__global__ void k_saturation(const float *__restrict__ in, int64_t n, float clean, float mix, float drive,
                             float *__restrict__ out)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = exciter(in[i], clean, mix, drive);
}

// ENG:117-134 on int16 samples: lut = 2^15 * exciter(s / 2^15) (PlanDev::sat_lut); the division by 2^15 is exact
__global__ void k_saturation_pcm(const int16_t *__restrict__ in, int64_t n, const float *__restrict__ lut, float *__restrict__ out)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = __ldg(lut + (unsigned)(unsigned short)in[i]) * (1.0f / 32768.0f);
}

template <typename T>
__global__ void k_width(const T *__restrict__ in, int64_t nframes, double width, T *__restrict__ out)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nframes; i += (int64_t)gridDim.x * blockDim.x) {
        const T l = in[2 * i], r = in[2 * i + 1];
        if (sizeof(T) == 8) {
            const double mid = __dmul_rn(__dadd_rn((double)l, (double)r), 0.5);
            const double side = __dmul_rn(__dmul_rn(__dsub_rn((double)l, (double)r), 0.5), width);
            out[2 * i] = (T)__dadd_rn(mid, side);
            out[2 * i + 1] = (T)__dsub_rn(mid, side);
        } else {
            const float mid = __fmul_rn(__fadd_rn((float)l, (float)r), 0.5f);
            const float side = __fmul_rn(__fmul_rn(__fsub_rn((float)l, (float)r), 0.5f), (float)width);
            out[2 * i] = (T)__fadd_rn(mid, side);
            out[2 * i + 1] = (T)__fsub_rn(mid, side);
        }
    }
}

template <typename T>
__global__ void k_limiter(const T *__restrict__ in, int64_t n, double thr, T *__restrict__ out)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (sizeof(T) == 8) out[i] = (T)limiter64((double)in[i], thr);
        else out[i] = (T)limiter32((float)in[i], (float)thr);
    }
}

// ENG:215 samples.mean(axis=1) in float32: fl(l + r) / 2
__global__ void k_mono_mean(const float *__restrict__ in, int64_t nframes, float *__restrict__ out)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nframes; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = __fmul_rn(__fadd_rn(in[2 * i], in[2 * i + 1]), 0.5f);
}

// ENG:222 samples * gain_linear (float32 array * float64 scalar -> float64)
__global__ void k_scale(const float *__restrict__ in, int64_t n, const double2 *__restrict__ loud, double *__restrict__ out)
{
    const double gain = loud[0].y;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = __dmul_rn((double)in[i], gain);
}

// scipy.signal.sosfilt from zero state on `channels` interleaved channels: one CTA per
// channel, sequential 4096-sample tiles, up to 8 sections.  IN float or double, out double.
template <typename IN>
__global__ void __launch_bounds__(KNT, 2)
k_sosfilt(const IN *__restrict__ in, int64_t nframes, int channels, const SecTab *__restrict__ secs,
          int nsec, double *__restrict__ out)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SecTab *tabs = reinterpret_cast<SecTab *>(smem_raw);                       // [8]
    double *sx = reinterpret_cast<double *>(smem_raw + 8 * sizeof(SecTab));    // [KTILE_PAD]
    double *carry = sx + KTILE_PAD;                                            // [8][2]
    double *wtot = carry + 16;                                                 // [2][8][2]
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, ch = blockIdx.x;
    {
        const double *s = reinterpret_cast<const double *>(secs);
        double *d = reinterpret_cast<double *>(tabs);
        for (int i = tid; i < nsec * (int)(sizeof(SecTab) / 8); i += KNT) d[i] = s[i];
        if (tid < 16) carry[tid] = 0.0;
    }
    __syncthreads();
    double *myx = sx + tid * (SEG + 1);
    unsigned round = 0;
    for (int64_t t0 = 0; t0 < nframes; t0 += KTILE) {
        const int nvalid = (int)min((int64_t)KTILE, nframes - t0);
        for (int f = tid; f < KTILE; f += KNT)
            sx[pidx(f)] = f < nvalid ? (double)in[(t0 + f) * channels + ch] : 0.0;
        __syncthreads();
        double x[SEG];
#pragma unroll
        for (int n = 0; n < SEG; ++n) x[n] = myx[n];
        for (int s = 0; s < nsec; ++s) {
            section_round<8>(x, tabs[s], tabs[s].Q, carry + 2 * s, wtot + (round & 1) * 16, lane, wid, tid == 0);
            ++round;
        }
#pragma unroll
        for (int n = 0; n < SEG; ++n) myx[n] = x[n];
        __syncthreads();
        for (int f = tid; f < nvalid; f += KNT) out[(t0 + f) * channels + ch] = sx[pidx(f)];
        __syncthreads();
    }
}

constexpr size_t sosfilt_smem_bytes() { return 8 * sizeof(SecTab) + (size_t)KTILE_PAD * 8 + 16 * 8 + 32 * 8; }

}  // namespace b200m
