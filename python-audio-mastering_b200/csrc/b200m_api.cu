// b200m_api.cu -- C-ABI of libb200master.so (include/b200_master.h): handle, host-side
// filter design, plan upload, batch orchestration and the stage-level entry points.
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>       // header-only NVTX 3: a range per kernel launch and per batch call (nsys / ncu --nvtx)

#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/b200_master.h"
#include "b200m_kernels.cuh"

using namespace b200m;

// ------------------------------------------------------------------------------------
struct ProfRec { int name; cudaEvent_t a, b; };

struct b200m_handle {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    // scratch arena
    char *ws = nullptr;
    size_t ws_cap = 0, ws_used = 0;
    int64_t ws_limit = (int64_t)64 << 30;
    // plan cache
    std::vector<b200m_plan> plans_key;
    std::vector<PlanDev> plans_host;
    PlanDev *d_plans = nullptr;
    size_t d_plans_cap = 0;
    // device tables keyed by content, least-recently-used entries evicted beyond a bound (a long-lived host --
    // the reference's worker is one -- sees a new slider combination per job; ensure_plans)
    struct DevTab { void *ptr; uint64_t stamp; int flag; };
    std::map<std::tuple<double, double>, DevTab> curves;     // (thresh_rms, slope) -> CURVE_N doubles
    std::map<uint64_t, DevTab> sat_luts;                     // content key -> 65536 floats (2^15 * exciter)
    uint64_t tab_tick = 0;
    std::map<int, std::pair<int32_t *, int>> pw_trees;      // block length -> (device table, smem floats)
    int blocks_smem_floats = 0;                              // max over the current plans
    int hops_smem_floats = 0;                                // the same for the hop trees (0: no plan shares hops)
    int hops_stage_floats = 0;                               // the longest hop of the current plans (k_hops stages a hop in shared memory), 0 when it would not fit
    // pinned staging for descriptors / small results
    char *pin = nullptr;
    size_t pin_cap = 0;
    // b200m_master_batch alternates between `pin` and `pinB` for its descriptors and results, and waits only for the call
    // that used a buffer last (two calls ago): the host side of call i + 1 (cutting groups and segments, ~1.5 ms for one
    // track) then runs while the kernels of call i do.  Every other entry point synchronises the stream before it touches `pin`.
    char *pinB = nullptr;
    size_t pinB_cap = 0;
    int pin_turn = 0;
    cudaEvent_t pin_ev[2] = {nullptr, nullptr};
    // profiling
    bool profiling = false;
    std::vector<std::string> prof_names;
    std::vector<ProfRec> prof_recs;
    std::vector<cudaEvent_t> ev_pool;
    std::map<std::string, std::pair<double, int64_t>> prof_acc;
    int64_t launches = 0;
    // recurrence tiling (0 = automatic) and its verification counters
    int recur_tile = 0, recur_warm = 0, recur_rounds = -1;
    bool comp_sprint = true;         // automatic repair = k_comp_sprint + one round (B200M_COMP_SPRINT=0: Jacobi rounds, as when a round count is set)
    // time segmentation of k_chain / k_kweight: 0 = automatic, < 0 = off, > 0 = tiles per segment
    int seg_chain = 0, seg_kweight = 0;
    int chain_kernel = 0;            // 0 = automatic, 1 = k_chain (a CTA per segment), 2 = k_chainw (a warp per segment)
    int kweight_kernel = 0;          // the same choice for k_kweight / k_kweightw (b200m_set_chain_kernel sets both)
    int detect_kernel = 0;           // 0 = automatic (k_detectw where the look-back fits its ring), 1 = k_detect always (B200M_DETECT_KERNEL: experiments, tests)
    int chain_waves = 4;             // k_chainw: the most resident waves of warps a group is cut into (B200M_CHAIN_WAVES)
    bool chain_slut_ok = true;       // k_chainw may take its 16-warp shape (B200M_CHAIN_SLUT=0 switches it off: experiments)
    unsigned long long *d_counters = nullptr;
    // host-buffer pipeline: side streams for H2D / D2H and the events that order the groups
    bool pipeline = true;
    int pipe_groups = 8, pipe_streams = 1;      // groups per batch; compute streams the groups alternate over
    double pipe_max_frames = 140e6;             // ... and the most frames a pipelined group takes (B200M_PIPE_MAX_FRAMES)
    cudaStream_t s_in = nullptr, s_out = nullptr, s_comp2 = nullptr;
    std::vector<cudaEvent_t> sync_events;
};

static std::string g_create_err;

static int fail(b200m_handle *h, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (h) h->err = buf; else g_create_err = buf;
    return code;
}

#define CK(call)                                                                          \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess)                                                            \
            return fail(h, B200M_ERR_CUDA, "%s failed: %s (%s:%d)", #call,                \
                        cudaGetErrorString(e_), __FILE__, __LINE__);                      \
    } while (0)

// ---- profiling: CUDA events around every launch on the handle's stream ----------------
static int prof_name_id(b200m_handle *h, const char *name)
{
    for (size_t i = 0; i < h->prof_names.size(); ++i)
        if (h->prof_names[i] == name) return (int)i;
    h->prof_names.push_back(name);
    return (int)h->prof_names.size() - 1;
}

static cudaEvent_t prof_event(b200m_handle *h)
{
    if (!h->ev_pool.empty()) { cudaEvent_t e = h->ev_pool.back(); h->ev_pool.pop_back(); return e; }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}

struct LaunchScope {
    b200m_handle *h; ProfRec r; bool on;
    LaunchScope(b200m_handle *h_, const char *name) : h(h_), on(h_->profiling)
    {
        ++h->launches;
        nvtxRangePushA(name);
        if (on) { r.name = prof_name_id(h, name); r.a = prof_event(h); r.b = prof_event(h); cudaEventRecord(r.a, h->stream); }
    }
    ~LaunchScope() { if (on) { cudaEventRecord(r.b, h->stream); h->prof_recs.push_back(r); } nvtxRangePop(); }
};
#define LAUNCH(name, ...) do { LaunchScope ls_(h, name); __VA_ARGS__; } while (0)

static int prof_collect(b200m_handle *h)
{
    if (h->prof_recs.empty()) return B200M_OK;
    CK(cudaStreamSynchronize(h->stream));
    for (auto &r : h->prof_recs) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r.a, r.b);
        auto &acc = h->prof_acc[h->prof_names[r.name]];
        acc.first += ms; acc.second += 1;
        h->ev_pool.push_back(r.a); h->ev_pool.push_back(r.b);
    }
    h->prof_recs.clear();
    return B200M_OK;
}

// ---- arena ---------------------------------------------------------------------------
static int ws_reserve(b200m_handle *h, size_t bytes)
{
    if (bytes <= h->ws_cap) return B200M_OK;
    CK(cudaStreamSynchronize(h->stream));
    if (h->ws) { cudaFree(h->ws); h->ws = nullptr; h->ws_cap = 0; }
    size_t want = bytes + (bytes >> 3);
    if (cudaMalloc(&h->ws, want) != cudaSuccess) {
        cudaGetLastError();
        want = bytes;
        if (cudaMalloc(&h->ws, want) != cudaSuccess) {
            cudaGetLastError();
            return fail(h, B200M_ERR_NOMEM, "cudaMalloc of %zu workspace bytes failed", bytes);
        }
    }
    h->ws_cap = want;
    return B200M_OK;
}

struct Arena {
    char *base; size_t used = 0;
    explicit Arena(char *b) : base(b) {}
    template <typename T> T *take(size_t n)
    {
        used = (used + 255) & ~(size_t)255;
        T *p = reinterpret_cast<T *>(base + used);
        used += n * sizeof(T);
        return p;
    }
};

static int pin_reserve(b200m_handle *h, size_t bytes)
{
    if (bytes <= h->pin_cap) return B200M_OK;
    CK(cudaStreamSynchronize(h->stream));
    if (h->pin) cudaFreeHost(h->pin);
    h->pin = nullptr; h->pin_cap = 0;
    CK(cudaMallocHost(&h->pin, bytes + 4096));
    h->pin_cap = bytes + 4096;
    return B200M_OK;
}

// ------------------------------------------------------------------------------------
// Host-side design
// ------------------------------------------------------------------------------------
static void build_sectab(const b200m_biquad &q, SecTab &T)
{
    typedef long double ld;
    std::memset(&T, 0, sizeof T);
    T.b0 = q.b0; T.b1 = q.b1; T.b2 = q.b2; T.a1 = q.a1; T.a2 = q.a2;
    const ld A[4] = {-(ld)q.a1, 1.0L, -(ld)q.a2, 0.0L};
    auto mm = [](const ld *X, const ld *Y, ld *Z) {
        ld r[4] = {X[0] * Y[0] + X[1] * Y[2], X[0] * Y[1] + X[1] * Y[3],
                   X[2] * Y[0] + X[3] * Y[2], X[2] * Y[1] + X[3] * Y[3]};
        for (int i = 0; i < 4; ++i) Z[i] = r[i];
    };
    ld v[2] = {(ld)q.b1 - (ld)q.a1 * (ld)q.b0, (ld)q.b2 - (ld)q.a2 * (ld)q.b0};
    for (int n = SEG - 1; n >= 0; --n) {
        T.g[n][0] = (double)v[0]; T.g[n][1] = (double)v[1];
        ld w[2] = {A[0] * v[0] + A[1] * v[1], A[2] * v[0] + A[3] * v[1]};
        v[0] = w[0]; v[1] = w[1];
    }
    ld AL[4] = {1, 0, 0, 1};
    for (int n = 0; n < SEG; ++n) mm(AL, A, AL);
    ld Pk[4] = {AL[0], AL[1], AL[2], AL[3]};
    for (int k = 0; k < 5; ++k) {
        for (int i = 0; i < 4; ++i) T.P[k][i] = (double)Pk[i];
        mm(Pk, Pk, Pk);
    }
    for (int i = 0; i < 4; ++i) T.PW[i] = (double)Pk[i];
    ld AH[4] = {1, 0, 0, 1};
    for (int n = 0; n < SEG / 2; ++n) mm(AH, A, AH);
    for (int i = 0; i < 4; ++i) T.AH[i] = (double)AH[i];
    ld Qj[4] = {1, 0, 0, 1};
    for (int j = 0; j < 32; ++j) {
        for (int i = 0; i < 4; ++i) T.Q[j][i] = (double)Qj[i];
        mm(Qj, AL, Qj);
    }
}

static const double kPi = 3.14159265358979323846;

// ENG:170-182 (doubled angle w = 2*pi*f/nyquist; `gain` used as RBJ's A)
static bool design_shelf(double rate, double f, double db, bool low, double qf, b200m_biquad *o)
{
    if (db == 0) return false;
    const double wn = f / (0.5 * rate), g = std::pow(10.0, db / 20.0), w = wn * 2 * kPi;
    const double alpha = std::sin(w) / (2.0 * qf), c = std::cos(w), rt = 2 * std::sqrt(g) * alpha;
    double b0, b1, b2, a0, a1, a2;
    if (low) {
        b0 = g * ((g + 1) - (g - 1) * c + rt); b1 = 2 * g * ((g - 1) - (g + 1) * c); b2 = g * ((g + 1) - (g - 1) * c - rt);
        a0 = (g + 1) + (g - 1) * c + rt; a1 = -2 * ((g - 1) + (g + 1) * c); a2 = (g + 1) + (g - 1) * c - rt;
    } else {
        b0 = g * ((g + 1) + (g - 1) * c + rt); b1 = -2 * g * ((g - 1) + (g + 1) * c); b2 = g * ((g + 1) + (g - 1) * c - rt);
        a0 = (g + 1) - (g - 1) * c + rt; a1 = 2 * ((g - 1) - (g + 1) * c); a2 = (g + 1) - (g - 1) * c - rt;
    }
    *o = {b0 / a0, b1 / a0, b2 / a0, a1 / a0, a2 / a0};
    return true;
}

// ENG:185-193
static bool design_peak(double rate, double f, double db, double qf, b200m_biquad *o)
{
    if (db == 0) return false;
    const double wn = f / (0.5 * rate), g = std::pow(10.0, db / 20.0), w = wn * 2 * kPi;
    const double alpha = std::sin(w) / (2.0 * qf);
    const double b0 = 1 + alpha * g, b1 = -2 * std::cos(w), b2 = 1 - alpha * g;
    const double a0 = 1 + alpha / g, a1 = -2 * std::cos(w), a2 = 1 - alpha / g;
    *o = {b0 / a0, b1 / a0, b2 / a0, a1 / a0, a2 / a0};
    return true;
}

// scipy.signal.butter(4, fc, btype, fs=rate, output='sos'): buttap -> lp2lp/lp2hp ->
// bilinear -> zpk2sos (pairing 'nearest': sections ordered with the poles closest to the
// unit circle last, overall gain on the first section).
static void design_butter4(double rate, double fc, bool highpass, b200m_biquad out[2])
{
    typedef std::complex<double> cd;
    const double wn = 2.0 * fc / rate, fs = 2.0;
    const double warped = 2 * fs * std::tan(kPi * wn / fs);
    cd p[4];
    double k = 1.0;
    for (int i = 0; i < 4; ++i) {
        const double m = -3 + 2 * i;
        p[i] = -std::exp(cd(0, kPi * m / 8.0));
    }
    cd z[4];
    if (!highpass) {
        for (int i = 0; i < 4; ++i) p[i] *= warped;
        k *= std::pow(warped, 4);
    } else {
        cd prod(1, 0);
        for (int i = 0; i < 4; ++i) prod *= -p[i];
        for (int i = 0; i < 4; ++i) { p[i] = warped / p[i]; z[i] = 0; }
        k *= (cd(1, 0) / prod).real();
    }
    const double fs2 = 2.0 * fs;
    cd num(1, 0), den(1, 0);
    if (highpass) for (int i = 0; i < 4; ++i) num *= (fs2 - z[i]);
    for (int i = 0; i < 4; ++i) den *= (fs2 - p[i]);
    cd pz[4];
    for (int i = 0; i < 4; ++i) pz[i] = (fs2 + p[i]) / (fs2 - p[i]);
    k *= (num / den).real();
    // conjugate pairs: (0,3) and (1,2) by construction of buttap
    double a1[2], a2[2], rad[2];
    const int pair[2][2] = {{0, 3}, {1, 2}};
    for (int s = 0; s < 2; ++s) {
        const cd q = pz[pair[s][0]];
        a1[s] = -2.0 * q.real();
        a2[s] = std::norm(q);
        rad[s] = std::abs(q);
    }
    const int first = rad[0] <= rad[1] ? 0 : 1, second = 1 - first;
    const double zb1 = highpass ? -2.0 : 2.0;
    out[0] = {k, k * zb1, k, a1[first], a2[first]};
    out[1] = {1.0, zb1, 1.0, a1[second], a2[second]};
}

// pyloudnorm IIRfilter coefficients of the K-weighting pre-filter
static void design_kweight(double rate, b200m_biquad out[2])
{
    {
        const double G = 4.0, Q = 1.0 / std::sqrt(2.0), fc = 1500.0;
        const double A = std::pow(10.0, G / 40.0), w0 = 2.0 * kPi * (fc / rate);
        const double alpha = std::sin(w0) / (2.0 * Q), c = std::cos(w0), rt = 2 * std::sqrt(A) * alpha;
        const double b0 = A * ((A + 1) + (A - 1) * c + rt), b1 = -2 * A * ((A - 1) + (A + 1) * c), b2 = A * ((A + 1) + (A - 1) * c - rt);
        const double a0 = (A + 1) - (A - 1) * c + rt, a1 = 2 * ((A - 1) - (A + 1) * c), a2 = (A + 1) - (A - 1) * c - rt;
        out[0] = {b0 / a0, b1 / a0, b2 / a0, a1 / a0, a2 / a0};
    }
    {
        const double Q = 0.5, fc = 38.0;
        const double w0 = 2.0 * kPi * (fc / rate), alpha = std::sin(w0) / (2.0 * Q), c = std::cos(w0);
        const double b0 = (1 + c) / 2, b1 = -(1 + c), b2 = (1 + c) / 2, a0 = 1 + alpha, a1 = -2 * c, a2 = 1 - alpha;
        out[1] = {b0 / a0, b1 / a0, b2 / a0, a1 / a0, a2 / a0};
    }
}

static void design_band(double rate, double thr_db, double ratio, double attack_ms, double release_ms, b200m_band *b)
{
    b->thresh_rms = 32768.0 * std::pow(10.0, thr_db / 20.0);
    b->attack_frames = attack_ms * (rate / 1000.0);
    b->release_frames = release_ms * (rate / 1000.0);
    b->slope = 1 - (1.0 / ratio);
    b->look_frames = (int32_t)b->attack_frames;
    b->reserved = 0;
}

// Frames after which a section's homogeneous response has decayed below 2^-64 (nothing in fp64).
static double settle_frames(const b200m_biquad &q)
{
    double r;
    const double disc = q.a1 * q.a1 - 4.0 * q.a2;
    if (disc < 0) r = std::sqrt(q.a2);
    else r = std::max(std::fabs((-q.a1 + std::sqrt(disc)) / 2), std::fabs((-q.a1 - std::sqrt(disc)) / 2));
    if (!(r < 1.0)) return 1e30;
    if (r < 1e-6) return 3.0;
    return std::ceil(44.4 / -std::log(r)) + 2.0;
}

// Warm-up (frames) that makes a k_chain / k_kweight segment independent of what came before it:
// the sections of a cascade settle one after the other, parallel branches (LP / HP) together.
constexpr size_t HOPS_SMEM_MAX = 100 * 1024;                // k_hops: tree values + one staged hop (19 KB at 48 kHz, 77 KB at 192 kHz)
static size_t hops_smem_bytes(const b200m_handle *h)
{
    return (size_t)(h->hops_smem_floats + 4 + h->hops_stage_floats) * 4 + 16;
}

static double chain_warm_frames(const b200m_plan &p)
{
    double w = 0;
    for (int s = 0; s < p.n_eq; ++s) w += settle_frames(p.eq[s]);
    if (p.multiband)
        w += std::max(settle_frames(p.lp[0]) + settle_frames(p.lp[1]), settle_frames(p.hp[0]) + settle_frames(p.hp[1]));
    return w;
}

static double kweight_warm_frames(const b200m_plan &p) { return settle_frames(p.kw[0]) + settle_frames(p.kw[1]); }

// Cut [0, frames) into segments of about `seg_tiles` tiles (never shorter than twice the warm-up);
// a single segment when seg_tiles <= 0 or the warm-up is unbounded / too long.
static void make_segments(std::vector<SegDesc> &out, int owner, int64_t frames, int tile, double warm_frames, int seg_tiles)
{
    if (frames <= 0) return;
    const int64_t ntiles = (frames + tile - 1) / tile;
    const double warm_tiles = std::ceil(warm_frames / tile);
    if (seg_tiles <= 0 || !(warm_tiles <= 64)) { out.push_back({0, frames, owner, 0}); return; }
    const int64_t seg = std::max<int64_t>(seg_tiles, 2 * (int64_t)warm_tiles);
    if (ntiles <= seg) { out.push_back({0, frames, owner, 0}); return; }
    const int warm = (int)warm_tiles * tile;
    for (int64_t t = 0; t < ntiles; t += seg) {
        const int64_t b = t * tile, e = std::min<int64_t>(frames, (t + seg) * tile);
        out.push_back({b, e, owner, b == 0 ? 0 : warm});
    }
}

// Warp segments of k_chainw: the stream is cut into the fewest equal runs of at most `seg_tiles` warp tiles (never
// shorter than twice the warm-up), one warp each.
static void make_segments_w(std::vector<SegDesc> &out, int owner, int64_t frames, int tile, double warm_frames, int seg_tiles)
{
    if (frames > 0) {
        const int64_t ntiles = (frames + tile - 1) / tile;
        const int64_t warm_tiles = (int64_t)std::ceil(warm_frames / tile);
        const int64_t min_seg = std::max<int64_t>(1, 2 * warm_tiles);
        int64_t nseg = std::max<int64_t>(1, (ntiles + std::max(1, seg_tiles) - 1) / std::max(1, seg_tiles));
        while (nseg > 1 && (ntiles + nseg - 1) / nseg < min_seg) --nseg;
        const int64_t seg = std::max(min_seg, (ntiles + nseg - 1) / nseg);
        const int warm = (int)(warm_tiles * tile);
        for (int64_t t = 0; t < ntiles; t += seg) {
            const int64_t b = t * tile, e = std::min<int64_t>(frames, (t + seg) * tile);
            out.push_back({b, e, owner, b == 0 ? 0 : warm});
        }
    } else {
        out.push_back({0, 0, owner, 0});
    }
}

// Segment length (in warp tiles) for a kernel whose warps each walk one segment, `resident` of them at a time: the
// kernel's time is waves x (longest segment + warm-up), so for w = 1 .. max_waves take the shortest length whose
// segment count still fits w resident waves and keep the cheapest (a few segments beyond a wave cost a whole extra
// pass).  0 when nothing fits.
static double plan_warp_segments(const std::vector<int64_t> &wtiles, double warm_tiles, double min_len, int max_waves, double resident)
{
    int64_t total_wtiles = 0;
    for (int64_t nt : wtiles) total_wtiles += nt;
    double best_len = 0, best_cost = 1e300;
    for (int w = 1; w <= max_waves; ++w) {
        const double cap = resident * w;
        double len = std::max(min_len, std::ceil((double)total_wtiles / cap));
        for (int it = 0; it < 200; ++it) {
            double total = 0;
            for (int64_t nt : wtiles) total += nt > 0 ? std::ceil((double)nt / len) : 0;
            if (total <= cap) break;
            len = std::ceil(len * std::max(1.003, total / cap));
        }
        double total = 0, longest = 0;
        for (int64_t nt : wtiles) if (nt > 0) { const double n = std::ceil((double)nt / len); total += n; longest = std::max(longest, std::ceil((double)nt / n)); }
        if (total > cap) continue;
        const double cost = w * (longest + warm_tiles);
        if (cost < best_cost) { best_cost = cost; best_len = len; }
    }
    return best_len;
}

static int auto_seg_tiles(int64_t total_tiles, int lo, int hi)
{
    return (int)std::max<int64_t>(lo, std::min<int64_t>(hi, total_tiles / (148 * 6)));
}

extern "C" int b200m_plan_from_settings(const b200m_settings *s, int sample_rate, int channels, b200m_plan *p)
{
    if (!s || !p || sample_rate <= 0 || (channels != 1 && channels != 2)) return B200M_ERR_INVALID;
    std::memset(p, 0, sizeof *p);
    p->sample_rate = sample_rate; p->channels = channels;
    p->sat_on = s->saturation != 0;
    const double mix = (s->saturation / 100.0) * (s->saturation / 100.0);
    p->sat_clean = (float)(1 - mix); p->sat_mix = (float)mix; p->sat_drive = (float)(1 + mix * 4);
    int n = 0;
    b200m_biquad q;
    if (design_shelf(sample_rate, 250, s->bass_boost, true, 0.707, &q)) p->eq[n++] = q;
    if (design_peak(sample_rate, 1000, -s->mid_cut, 1.0, &q)) p->eq[n++] = q;
    if (design_peak(sample_rate, 4000, s->presence_boost, 1.0, &q)) p->eq[n++] = q;
    if (design_shelf(sample_rate, 8000, s->treble_boost, false, 0.707, &q)) p->eq[n++] = q;
    p->n_eq = n;
    p->width_on = (channels == 2) && (s->width != 1.0);
    p->width = s->width;
    p->multiband = s->multiband != 0;
    if (p->multiband) {
        if (s->low_ratio == 0 || s->mid_ratio == 0 || s->high_ratio == 0) return B200M_ERR_INVALID;
        if (!(4000.0 < 0.5 * sample_rate)) return B200M_ERR_INVALID;   // scipy butter raises too
        design_butter4(sample_rate, 250, false, p->lp);
        design_butter4(sample_rate, 4000, true, p->hp);
        design_band(sample_rate, s->low_thresh, s->low_ratio, 10.0, 200.0, &p->band[0]);
        design_band(sample_rate, s->mid_thresh, s->mid_ratio, 5.0, 150.0, &p->band[1]);
        design_band(sample_rate, s->high_thresh, s->high_ratio, 1.0, 50.0, &p->band[2]);
    }
    design_kweight(sample_rate, p->kw);
    p->has_lufs = s->has_lufs != 0;
    p->lufs = s->lufs;
    return B200M_OK;
}

// ------------------------------------------------------------------------------------
// Plan upload (device tables), cached across calls with identical plans
// ------------------------------------------------------------------------------------
constexpr size_t MAX_CURVES = 96, MAX_SAT_LUTS = 32;        // 25 MB + 8 MB of device tables at most (beyond what one call uses)

// Make room in a table cache: entries not touched by the rebuild in progress (stamp < tick) go first, oldest first.
// The caller has synchronised the handle's stream, so nothing in flight reads them.
template <typename M> static void evict_tables(M &m, size_t cap, uint64_t tick)
{
    while (m.size() >= cap) {
        auto victim = m.end();
        for (auto it = m.begin(); it != m.end(); ++it)
            if (it->second.stamp < tick && (victim == m.end() || it->second.stamp < victim->second.stamp)) victim = it;
        if (victim == m.end()) return;           // everything is in use by this very call: grow
        cudaFree(victim->second.ptr);
        m.erase(victim);
    }
}

static int get_curve(b200m_handle *h, const b200m_band &b, const double **out, std::vector<double> *host_copy)
{
    // pydub: db = 20 * math.log(rms / thresh_rms, 10) = 20 * (log(x) / log(10));
    //        max_attenuation = (1 - 1/ratio) * max(db, 0)      (0 whenever rms <= thresh_rms)
    std::vector<double> tab(CURVE_N);
    const double l10 = std::log(10.0);
    for (int r = 0; r < CURVE_N; ++r) {
        double over = 0.0;
        if (r != 0) {
            const double db = 20 * (std::log((double)r / b.thresh_rms) / l10);
            over = db > 0 ? db : 0.0;
        }
        double M = b.slope * over;
        if (M == 0) M = 0.0;                     // no negative zero: decisions are integer compares
        tab[r] = M;
    }
    if (host_copy) *host_copy = tab;
    auto key = std::make_tuple(b.thresh_rms, b.slope);
    auto it = h->curves.find(key);
    if (it != h->curves.end()) { it->second.stamp = h->tab_tick; *out = (const double *)it->second.ptr; return B200M_OK; }
    evict_tables(h->curves, MAX_CURVES, h->tab_tick);
    double *d = nullptr;
    CK(cudaMalloc(&d, CURVE_N * sizeof(double)));
    if (cudaMemcpy(d, tab.data(), CURVE_N * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaFree(d);
        return fail(h, B200M_ERR_CUDA, "upload of a compressor curve failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    h->curves[key] = {d, h->tab_tick, 0};
    *out = d;
    return B200M_OK;
}

static uint64_t hash_words(const void *p, size_t bytes, uint64_t seed)
{
    uint64_t hsh = 0xcbf29ce484222325ull ^ seed;
    const unsigned char *c = (const unsigned char *)p;
    size_t i = 0;
    for (; i + 8 <= bytes; i += 8) { uint64_t w; std::memcpy(&w, c + i, 8); hsh = (hsh ^ w) * 0x100000001b3ull; hsh ^= hsh >> 29; }
    for (; i < bytes; ++i) hsh = (hsh ^ c[i]) * 0x100000001b3ull;
    return hsh;
}

// ENG:128-134 tabulated over the 65536 int16 samples (b200m_plan::sat_lut), uploaded pre-scaled by 2^15 (exact):
// the chain kernels work on 2^15 x up to quantise #1.
static int get_sat_lut(b200m_handle *h, const b200m_plan &p, const float **out, int *odd = nullptr)
{
    std::vector<float> own;
    const float *src = p.sat_lut;
    uint64_t key = p.sat_lut_key;
    if (!src) {
        // no table from the host: the same expression with libm's tanhf, every product / sum rounded to float32
        own.resize(65536);
        for (int i = 0; i < 65536; ++i) {
            const float x = (float)(int16_t)(uint16_t)i / 32768.0f;
            volatile float arg = x * p.sat_drive;
            volatile float a = p.sat_clean * x, b = p.sat_mix * std::tanh((float)arg);
            own[i] = a + b;
        }
        src = own.data();
        const float par[3] = {p.sat_clean, p.sat_mix, p.sat_drive};
        key = hash_words(par, sizeof par, 0x6c69626dull /* "libm" */);
    } else if (key == 0) {
        key = hash_words(src, 65536 * sizeof(float), 0);
    }
    auto it = h->sat_luts.find(key);
    if (it != h->sat_luts.end()) {
        it->second.stamp = h->tab_tick; *out = (const float *)it->second.ptr;
        if (odd) *odd = it->second.flag;
        return B200M_OK;
    }
    evict_tables(h->sat_luts, MAX_SAT_LUTS, h->tab_tick);
    std::vector<float> scaled(65536);
    for (int i = 0; i < 65536; ++i) scaled[i] = src[i] * 32768.0f;      // exact (power of two; |value| ~ 1)
    // odd in the sample?  entry(-s) == -entry(s) bit for bit for s = 1 .. 32767 (numpy's tanh and IEEE products are)
    int sym = 1;
    for (int m = 1; m < 32768 && sym; ++m) {
        uint32_t pos, neg;
        std::memcpy(&pos, &scaled[m], 4); std::memcpy(&neg, &scaled[65536 - m], 4);
        if ((pos ^ 0x80000000u) != neg) sym = 0;
    }
    if (odd) *odd = sym;
    float *d = nullptr;
    CK(cudaMalloc(&d, 65536 * sizeof(float)));
    if (cudaMemcpy(d, scaled.data(), 65536 * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaFree(d);
        return fail(h, B200M_ERR_CUDA, "upload of an exciter table failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    h->sat_luts[key] = {d, h->tab_tick, sym};
    *out = d;
    return B200M_OK;
}

// numpy's pairwise_sum recursion tree for n elements (numpy/_core/src/umath/loops_utils.h.src):
//   n <= 128: leaf;  else n2 = n / 2, n2 -= n2 % 8, sum(a, n2) + sum(a + n2, n - n2).
// Table: [n, nleaves, ninternal, nlevels, level_start[nlevels + 1], leaf_off[], leaf_n[], node_l[], node_r[]];
// internal nodes are ordered deepest level first (the root is last) and store to value slot
// nleaves + i; children are value slots.
static void build_pw_tree(int n, std::vector<int32_t> &tab)
{
    struct Node { int l, r, depth; };
    std::vector<int32_t> leaf_off, leaf_n;
    std::vector<Node> internal;
    struct Rec {
        std::vector<int32_t> &lo, &ln; std::vector<Node> &in;
        int run(int off, int cnt, int depth)
        {
            if (cnt <= 128) { lo.push_back(off); ln.push_back(cnt); return (int)lo.size() - 1; }
            int n2 = cnt / 2;
            n2 -= n2 % 8;
            const int L = run(off, n2, depth + 1), R = run(off + n2, cnt - n2, depth + 1);
            in.push_back({L, R, depth});
            return -(int)in.size();
        }
    } rec{leaf_off, leaf_n, internal};
    rec.run(0, n, 0);
    const int nl = (int)leaf_off.size(), ni = (int)internal.size();
    std::vector<int> order(ni), pos(ni);
    for (int i = 0; i < ni; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return internal[a].depth > internal[b].depth; });
    for (int i = 0; i < ni; ++i) pos[order[i]] = i;
    std::vector<int32_t> lvl;
    for (int i = 0; i < ni; ++i)
        if (i == 0 || internal[order[i]].depth != internal[order[i - 1]].depth) lvl.push_back(i);
    const int nlev = (int)lvl.size();
    lvl.push_back(ni);
    auto slot = [&](int id) { return id >= 0 ? id : nl + pos[-id - 1]; };
    tab.clear();
    tab.push_back(n); tab.push_back(nl); tab.push_back(ni); tab.push_back(nlev);
    tab.insert(tab.end(), lvl.begin(), lvl.end());
    tab.insert(tab.end(), leaf_off.begin(), leaf_off.end());
    tab.insert(tab.end(), leaf_n.begin(), leaf_n.end());
    for (int i = 0; i < ni; ++i) tab.push_back(slot(internal[order[i]].l));
    for (int i = 0; i < ni; ++i) tab.push_back(slot(internal[order[i]].r));
}

static int get_pw_tree(b200m_handle *h, int n, const int32_t **out, int *smem_floats)
{
    *out = nullptr; *smem_floats = 0;
    if (n < 8 || n > (1 << 22)) return B200M_OK;             // tiny / absurd block lengths: serial walk
    auto it = h->pw_trees.find(n);
    if (it == h->pw_trees.end()) {
        std::vector<int32_t> tab;
        build_pw_tree(n, tab);
        const int floats = tab[1] + tab[2];
        int32_t *d = nullptr;
        if (floats * 4 <= 40 * 1024) {
            CK(cudaMalloc(&d, tab.size() * sizeof(int32_t)));
            CK(cudaMemcpy(d, tab.data(), tab.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
        }
        it = h->pw_trees.emplace(n, std::make_pair(d, d ? floats : 0)).first;
    }
    *out = it->second.first; *smem_floats = it->second.second;
    return B200M_OK;
}

// Is q = fma(fma(-m*rc, c, m), rc, m*rc) the correctly rounded m / c for every curve value?
static bool div_trick_exact(const std::vector<double> &tab, double c)
{
    if (!(c > 0) || !std::isfinite(c)) return false;
    const double rc = 1.0 / c;
    for (double m : tab) {
        volatile double q = m * rc;
        const double r = std::fma(-q, c, m);
        const double q2 = std::fma(r, rc, q);
        if (q2 != m / c) return false;
    }
    return true;
}

static int ensure_plans(b200m_handle *h, const b200m_plan *plans, int n)
{
    if ((int)h->plans_key.size() == n && n > 0 &&
        std::memcmp(h->plans_key.data(), plans, n * sizeof(b200m_plan)) == 0)
        return B200M_OK;
    CK(cudaStreamSynchronize(h->stream));   // previous launches may still read d_plans and the cached tables
    // Everything is built into locals and swapped in only after the upload succeeded; the cache key is
    // dropped first, so a rebuild that fails half-way can never be mistaken for the plans it replaced.
    h->plans_key.clear();
    ++h->tab_tick;
    std::vector<PlanDev> host(n);
    int blocks_smem = 0, hops_smem = 0, hops_stage = 0;
    for (int i = 0; i < n; ++i) {
        const b200m_plan &p = plans[i];
        PlanDev &d = host[i];
        std::memset(&d, 0, sizeof d);
        if (p.channels != 1 && p.channels != 2) return fail(h, B200M_ERR_INVALID, "plan %d: channels must be 1 or 2", i);
        if (p.n_eq < 0 || p.n_eq > 4) return fail(h, B200M_ERR_INVALID, "plan %d: n_eq out of range", i);
        d.rate = p.sample_rate; d.channels = p.channels; d.sat_on = p.sat_on; d.n_eq = p.n_eq;
        d.width_on = p.width_on && p.channels == 2; d.multiband = p.multiband; d.has_lufs = p.has_lufs;
        d.sat_clean = p.sat_clean; d.sat_mix = p.sat_mix; d.sat_drive = p.sat_drive;
        d.width = p.width; d.lufs = p.lufs;
        if (p.sat_on) { int rc = get_sat_lut(h, p, &d.sat_lut, &d.sat_sym); if (rc) return rc; }
        for (int s = 0; s < p.n_eq; ++s) build_sectab(p.eq[s], d.eq[s]);
        build_sectab(p.kw[0], d.kw[0]);
        build_sectab(p.kw[1], d.kw[1]);
        if (p.has_lufs) {
            // full block length: u - l of pyloudnorm's block 0 = int(0.4 * 1.0 * rate) - 0
            int fl = 0;
            int rc = get_pw_tree(h, (int)(int64_t)(0.4 * (0.0 * 0.25 + 1.0) * (double)p.sample_rate), &d.ptree, &fl);
            if (rc) return rc;
            blocks_smem = std::max(blocks_smem, fl);
            // numpy's tree over n elements splits at n2 = n/2 - (n/2) % 8: when n is a multiple of 32 the
            // first two levels cut at n/2 and n/4, i.e. a block is four hop trees (k_hops)
            const int nblk = (int)(int64_t)(0.4 * (0.0 * 0.25 + 1.0) * (double)p.sample_rate);
            if (d.ptree && nblk % 32 == 0 && nblk > 512) {
                int hf = 0;
                rc = get_pw_tree(h, nblk / 4, &d.htree, &hf);
                if (rc) return rc;
                if (d.htree) { d.hop = nblk / 4; hops_smem = std::max(hops_smem, hf); hops_stage = std::max(hops_stage, d.hop); }
            }
        }
        if (p.multiband) {
            for (int s = 0; s < 2; ++s) { build_sectab(p.lp[s], d.lp[s]); build_sectab(p.hp[s], d.hp[s]); }
            for (int b = 0; b < 3; ++b) {
                const b200m_band &bb = p.band[b];
                if (bb.look_frames < 0 || bb.look_frames > 8192)
                    return fail(h, B200M_ERR_INVALID, "plan %d band %d: look_frames %d unsupported", i, b, bb.look_frames);
                std::vector<double> tab;
                int rc = get_curve(h, bb, &d.curve[b], &tab);
                if (rc) return rc;
                bool trick = div_trick_exact(tab, bb.attack_frames) && div_trick_exact(tab, bb.release_frames);
                // the fast path of the recurrence also assumes finite values >= +0 and positive attack / release
                for (double m : tab) if (!(m >= 0.0) || !std::isfinite(m) || std::signbit(m)) { trick = false; break; }
                if (!(bb.attack_frames > 0) || !(bb.release_frames > 0)) trick = false;
                // curve[r] == 0 for r <= hold_max and != 0 above it (the curve is monotone); if that
                // does not hold for some odd parameter set, nothing is ever flagged "held"
                int hold_max = -1;
                {
                    int r = 0;
                    while (r < CURVE_N && tab[r] == 0.0) ++r;
                    hold_max = r - 1;
                    for (; r < CURVE_N; ++r) if (tab[r] == 0.0) { hold_max = -1; break; }
                }
                // the recurrence keeps 0 <= att <= max(curve) when every curve value is finite and >= +0
                bool bounded = true;
                for (double m : tab) if (!(m >= 0.0) || !(m <= 5000.0) || std::signbit(m)) { bounded = false; break; }
                d.band[b] = {bb.thresh_rms, bb.attack_frames, bb.release_frames, bb.slope,
                             1.0 / bb.attack_frames, 1.0 / bb.release_frames, bb.look_frames, trick ? 1 : 0, hold_max, bounded ? 1 : 0};
            }
        }
    }
    if ((size_t)n > h->d_plans_cap) {
        if (h->d_plans) cudaFree(h->d_plans);
        h->d_plans = nullptr; h->d_plans_cap = 0;
        CK(cudaMalloc(&h->d_plans, n * sizeof(PlanDev)));
        h->d_plans_cap = n;
    }
    CK(cudaMemcpyAsync(h->d_plans, host.data(), n * sizeof(PlanDev), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->plans_host.swap(host);
    h->blocks_smem_floats = blocks_smem;
    h->hops_smem_floats = hops_smem;
    h->hops_stage_floats = ((size_t)(hops_smem + 4 + hops_stage) * 4 + 16 <= HOPS_SMEM_MAX) ? hops_stage : 0;
    h->plans_key.assign(plans, plans + n);
    return B200M_OK;
}

// ------------------------------------------------------------------------------------
// lifetime
// ------------------------------------------------------------------------------------
template <typename K> static cudaError_t allow_smem(K kernel, size_t bytes)
{
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

static size_t detect_smem_bytes(int H)
{
    const int R = detect_run(H);
    return (size_t)(R * DNT + 4) * 8 + (size_t)R * DNT * 4;
}

extern "C" int b200m_abi_version(void) { return B200M_ABI_VERSION; }

extern "C" int b200m_create(int device, b200m_handle **out)
{
    b200m_handle *h = nullptr;
    if (!out) return B200M_ERR_INVALID;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(nullptr, B200M_ERR_CUDA, "no CUDA device: libb200master has no CPU fallback");
    }
    if (device < 0 || device >= count) return fail(nullptr, B200M_ERR_INVALID, "device %d out of range (0..%d)", device, count - 1);
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail(nullptr, B200M_ERR_CUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
    h = new b200m_handle();
    h->device = device;
    if (const char *ck = std::getenv("B200M_CHAIN_KERNEL")) h->chain_kernel = std::max(0, std::min(2, std::atoi(ck)));   // test / experiment override
    if (const char *ck = std::getenv("B200M_KWEIGHT_KERNEL")) h->kweight_kernel = std::max(0, std::min(2, std::atoi(ck)));
    if (const char *ck = std::getenv("B200M_COMP_SPRINT")) h->comp_sprint = std::atoi(ck) != 0;
    if (const char *ck = std::getenv("B200M_CHAIN_SLUT")) h->chain_slut_ok = std::atoi(ck) != 0;
    if (const char *ck = std::getenv("B200M_DETECT_KERNEL")) h->detect_kernel = std::atoi(ck);
    if (const char *ck = std::getenv("B200M_CHAIN_WAVES")) h->chain_waves = std::max(1, std::min(16, std::atoi(ck)));
    if (const char *ck = std::getenv("B200M_PIPE_MAX_FRAMES")) h->pipe_max_frames = std::max(8e6, std::atof(ck));
    e = allow_smem(k_chain<1, true>, chain_smem_bytes<1>());
    if (e == cudaSuccess) e = allow_smem(k_chain<2, true>, chain_smem_bytes<2>());
    if (e == cudaSuccess) e = allow_smem(k_chain<1, false>, chain_smem_bytes<1>());
    if (e == cudaSuccess) e = allow_smem(k_chain<2, false>, chain_smem_bytes<2>());
    if (e == cudaSuccess) e = allow_smem(k_chainw<1, true, true>, ChainW<1>::SMEM);
    if (e == cudaSuccess) e = allow_smem(k_chainw<1, false, true>, ChainW<1>::SMEM);
    if (e == cudaSuccess) e = allow_smem(k_chainw<2, true, true>, ChainW<2>::SMEM);
    if (e == cudaSuccess) e = allow_smem(k_chainw<2, false, true>, ChainW<2>::SMEM);
    if (e == cudaSuccess) e = allow_smem(k_chainw<1, true, false>, ChainW<1>::SMEM);
    if (e == cudaSuccess) e = allow_smem(k_chainw<1, false, false>, ChainW<1>::SMEM);
    if (e == cudaSuccess) e = allow_smem(k_chainw<2, true, false>, ChainW<2>::SMEM);
    if (e == cudaSuccess) e = allow_smem(k_chainw<2, false, false>, ChainW<2>::SMEM);
    if (e == cudaSuccess) e = allow_smem(k_chainw<1, true, true, CW_SLUT, true>, ChainW<1>::SMEM_SLUT);
    if (e == cudaSuccess) e = allow_smem(k_chainw<1, false, true, CW_SLUT, true>, ChainW<1>::SMEM_SLUT);
    if (e == cudaSuccess) e = allow_smem(k_chainw<2, true, true, CW_SLUT, true>, ChainW<2>::SMEM_SLUT);
    if (e == cudaSuccess) e = allow_smem(k_chainw<2, false, true, CW_SLUT, true>, ChainW<2>::SMEM_SLUT);
    if (e == cudaSuccess) e = allow_smem(k_chainw<1, true, false, CW_SLUT, true>, ChainW<1>::SMEM_SLUT_MP);
    if (e == cudaSuccess) e = allow_smem(k_chainw<1, false, false, CW_SLUT, true>, ChainW<1>::SMEM_SLUT_MP);
    if (e == cudaSuccess) e = allow_smem(k_chainw<2, true, false, CW_SLUT, true>, ChainW<2>::SMEM_SLUT_MP);
    if (e == cudaSuccess) e = allow_smem(k_chainw<2, false, false, CW_SLUT, true>, ChainW<2>::SMEM_SLUT_MP);
    if (e == cudaSuccess) e = allow_smem(k_hops, HOPS_SMEM_MAX);
    if (e == cudaSuccess) e = allow_smem(k_detect<1>, detect_smem_bytes(8192));
    if (e == cudaSuccess) e = allow_smem(k_detect<2>, detect_smem_bytes(8192));
    if (e == cudaSuccess) e = allow_smem(k_comp<1, 1, true>, recur_smem_bytes(1));
    if (e == cudaSuccess) e = allow_smem(k_comp<2, 1, true>, recur_smem_bytes(1));
    if (e == cudaSuccess) e = allow_smem(k_comp<1, 3, false>, recur_smem_bytes(3));
    if (e == cudaSuccess) e = allow_smem(k_comp<2, 3, false>, recur_smem_bytes(3));
    if (const char *ck = std::getenv("B200M_COMP_CARVEOUT")) {          // experiment: shared-memory carve-out (percent) of the compressor kernel
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_comp<2, 3, false>, cudaFuncAttributePreferredSharedMemoryCarveout, std::atoi(ck));
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_comp<1, 3, false>, cudaFuncAttributePreferredSharedMemoryCarveout, std::atoi(ck));
    }
    if (e == cudaSuccess) e = allow_smem(k_kweight<1, int16_t, true>, kweight_smem_bytes());
    if (e == cudaSuccess) e = allow_smem(k_kweight<2, int16_t, true>, kweight_smem_bytes());
    if (e == cudaSuccess) e = allow_smem(k_kweight<1, float, true>, kweight_smem_bytes());
    if (e == cudaSuccess) e = allow_smem(k_kweight<1, int16_t, false>, kweight_smem_bytes());
    if (e == cudaSuccess) e = allow_smem(k_kweight<2, int16_t, false>, kweight_smem_bytes());
    if (e == cudaSuccess) e = allow_smem(k_kweight<1, float, false>, kweight_smem_bytes());
    if (e == cudaSuccess) e = allow_smem(k_sosfilt<float>, sosfilt_smem_bytes());
    if (e == cudaSuccess) e = allow_smem(k_sosfilt<double>, sosfilt_smem_bytes());
    if (e != cudaSuccess) {
        fail(nullptr, B200M_ERR_CUDA, "kernel attribute set-up failed (is this an sm_100a device?): %s", cudaGetErrorString(e));
        delete h;
        return B200M_ERR_CUDA;
    }
    if (cudaMalloc(&h->d_counters, 16 * sizeof(unsigned long long)) != cudaSuccess ||
        cudaMemset(h->d_counters, 0, 16 * sizeof(unsigned long long)) != cudaSuccess) {
        fail(nullptr, B200M_ERR_CUDA, "counter allocation failed");
        delete h;
        return B200M_ERR_CUDA;
    }
    *out = h;
    return B200M_OK;
}

extern "C" void b200m_destroy(b200m_handle *h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    if (h->ws) cudaFree(h->ws);
    if (h->d_plans) cudaFree(h->d_plans);
    if (h->pin) cudaFreeHost(h->pin);
    if (h->pinB) cudaFreeHost(h->pinB);
    for (cudaEvent_t e : h->pin_ev) if (e) cudaEventDestroy(e);
    if (h->d_counters) cudaFree(h->d_counters);
    for (auto &kv : h->curves) cudaFree(kv.second.ptr);
    for (auto &kv : h->sat_luts) cudaFree(kv.second.ptr);
    for (auto &kv : h->pw_trees) if (kv.second.first) cudaFree(kv.second.first);
    for (auto &r : h->prof_recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    for (auto e : h->ev_pool) cudaEventDestroy(e);
    for (auto e : h->sync_events) cudaEventDestroy(e);
    if (h->s_in) cudaStreamDestroy(h->s_in);
    if (h->s_out) cudaStreamDestroy(h->s_out);
    if (h->s_comp2) cudaStreamDestroy(h->s_comp2);
    delete h;
}

extern "C" const char *b200m_last_error(const b200m_handle *h) { return h ? h->err.c_str() : g_create_err.c_str(); }

extern "C" int b200m_set_stream(b200m_handle *h, void *s)
{
    if (!h) return B200M_ERR_INVALID;
    CK(cudaStreamSynchronize(h->stream));
    h->stream = (cudaStream_t)s;
    return B200M_OK;
}

extern "C" int b200m_synchronize(b200m_handle *h)
{
    if (!h) return B200M_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    return B200M_OK;
}

extern "C" int b200m_set_workspace_limit(b200m_handle *h, int64_t bytes)
{
    if (!h || bytes < (1 << 20)) return B200M_ERR_INVALID;
    h->ws_limit = bytes;
    return B200M_OK;
}

extern "C" int64_t b200m_launch_count(const b200m_handle *h) { return h ? h->launches : 0; }

extern "C" int b200m_set_profiling(b200m_handle *h, int on)
{
    if (!h) return B200M_ERR_INVALID;
    int rc = prof_collect(h);
    h->profiling = on != 0;
    return rc;
}

extern "C" int b200m_kernel_time_ms(b200m_handle *h, const char *kernel, double *total_ms, int64_t *launches)
{
    if (!h || !kernel) return B200M_ERR_INVALID;
    int rc = prof_collect(h);
    if (rc) return rc;
    auto it = h->prof_acc.find(kernel);
    if (total_ms) *total_ms = it == h->prof_acc.end() ? 0.0 : it->second.first;
    if (launches) *launches = it == h->prof_acc.end() ? 0 : it->second.second;
    return B200M_OK;
}

extern "C" int b200m_set_recur_tiling(b200m_handle *h, int tile_frames, int warm_frames, int rounds)
{
    if (!h || tile_frames < 0 || warm_frames < 0 || rounds > 64) return B200M_ERR_INVALID;
    h->recur_tile = tile_frames;
    h->recur_warm = warm_frames > 0 ? warm_frames : 0;
    h->recur_rounds = rounds < 0 ? -1 : rounds;
    return B200M_OK;
}

extern "C" int b200m_set_segment_tiles(b200m_handle *h, int chain_tiles, int kweight_tiles)
{
    if (!h) return B200M_ERR_INVALID;
    h->seg_chain = chain_tiles;
    h->seg_kweight = kweight_tiles;
    return B200M_OK;
}

extern "C" int b200m_set_pipeline(b200m_handle *h, int on)
{
    if (!h) return B200M_ERR_INVALID;
    h->pipeline = on != 0;
    return B200M_OK;
}

extern "C" int b200m_set_chain_kernel(b200m_handle *h, int mode)
{
    if (!h || mode < 0 || mode > 2) return B200M_ERR_INVALID;
    h->chain_kernel = mode;
    h->kweight_kernel = mode;
    return B200M_OK;
}

extern "C" int b200m_set_pipeline_shape(b200m_handle *h, int groups, int compute_streams)
{
    if (!h) return B200M_ERR_INVALID;
    if (groups < 0 || groups > 4096 || compute_streams < 0 || compute_streams > 2)
        return fail(h, B200M_ERR_INVALID, "b200m_set_pipeline_shape: groups 0..4096, compute_streams 0..2");
    h->pipe_groups = groups == 0 ? 8 : groups;
    h->pipe_streams = compute_streams == 0 ? 1 : compute_streams;
    return B200M_OK;
}

extern "C" int b200m_recur_stats(b200m_handle *h, int64_t *wrong_tiles, int64_t *rerun_frames, int64_t *round_repairs, int reset)
{
    if (!h) return B200M_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    unsigned long long c[4];
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaMemcpy(c, h->d_counters, sizeof c, cudaMemcpyDeviceToHost));
    if (wrong_tiles) *wrong_tiles = (int64_t)c[0];
    if (rerun_frames) *rerun_frames = (int64_t)c[1];
    if (round_repairs) *round_repairs = (int64_t)c[2];
    if (reset) CK(cudaMemset(h->d_counters, 0, sizeof c));
    return B200M_OK;
}

// build experiments (-DB200M_RECUR_TIMING): the raw counter block, counters[8 ..] = per-phase cycles
extern "C" int b200m_debug_counters(b200m_handle *h, unsigned long long *out16, int reset)
{
    if (!h || !out16) return B200M_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaMemcpy(out16, h->d_counters, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    if (reset) CK(cudaMemset(h->d_counters, 0, 16 * sizeof(unsigned long long)));
    return B200M_OK;
}

extern "C" int b200m_reset_profile(b200m_handle *h)
{
    if (!h) return B200M_ERR_INVALID;
    int rc = prof_collect(h);
    h->prof_acc.clear();
    return rc;
}

static void fill_tabc(SecTabC &d, const SecTab &s)
{
    d.b0 = s.b0; d.b1 = s.b1; d.b2 = s.b2; d.a1 = s.a1; d.a2 = s.a2;
    std::memcpy(d.g, s.g, sizeof d.g);
    std::memcpy(d.P, s.P, sizeof d.P);
    std::memcpy(d.PW, s.PW, sizeof d.PW);
    std::memcpy(d.AH, s.AH, sizeof d.AH);
}

// ------------------------------------------------------------------------------------
// The chain on device buffers (shared by master_batch and the stage entry points)
// ------------------------------------------------------------------------------------
struct Group {
    int ch = 2;
    int n_streams = 0, n_tracks = 0;
    int max_stream_frames = 0;
    int64_t max_track_frames = 0;
    int max_blocks = 0;
    int max_look = 0;
    int64_t total_blocks = 0;        // hold-flag WORDS: sum over streams of ceil(out_frames / 1024)
    bool any_multiband = false, any_lufs = false;
    bool chain_stable = true;        // every EQ / crossover section of every plan has its poles inside the unit circle
    const StreamDesc *d_streams = nullptr;
    const TrackDesc *d_tracks = nullptr;
    const SegDesc *d_csegs = nullptr, *d_ksegs = nullptr;   // k_chain / k_kweight segments
    int n_csegs = 0, n_ksegs = 0;
    bool chain_warps = false;        // csegs are warp segments (k_chainw)
    bool kw_warps = false;           // ksegs are warp segments (k_kweightw)
    bool chain_slut = false;         // ... in its 16-warp shape around a shared-memory exciter table (padded to whole CTAs, d_cta_iters)
    const int32_t *d_cta_iters = nullptr;
    int single_plan = -1;            // >= 0: every stream of the group uses this plan (its tables travel as a kernel parameter)
};

#ifndef B200M_COMP_CTAS
#define B200M_COMP_CTAS B200M_COMP_MINB   // resident CTAs per SM (42 KB of shared memory each) k_comp's automatic tile length aims at
#endif
static RecurParams recur_params(const b200m_handle *h, const Group &g, int nbands, int band_base)
{
    RecurParams P;
    P.nbands = nbands; P.band_base = band_base; P.n_streams = g.n_streams;
    P.warm = h->recur_warm > 0 ? std::max(32, (h->recur_warm + 31) & ~31) : 0;      // 0: automatic, per band (k_comp)
    if (h->recur_tile > 0) {
        P.tile_len = std::max(32, (h->recur_tile + 31) & ~31);
    } else {
        // One lane per (stream, tile), one warp per band: about one resident wave of lanes (148 SMs x
        // B200M_COMP_CTAS CTAs x 32); tiles of 4096 .. 262144 frames.  Every lane also runs the warm-up
        // (`warm` active frames, level detector -> curve -> recurrence only), so long tiles waste less
        // work and short tiles give a small batch enough lanes.
        // (the tile count is rounded DOWN: a few CTAs beyond the resident wave would cost a whole second pass)
        const double want_tiles = std::max(1.0, std::floor(148.0 * B200M_COMP_CTAS * 32 / std::max(1, g.n_streams)));
        const double len = std::min(262144.0, std::max(4096.0, std::ceil(g.max_stream_frames / want_tiles)));
        P.tile_len = ((int)len + 1023) & ~1023;
    }
    P.tiles = std::max(1, (g.max_stream_frames + P.tile_len - 1) / P.tile_len);
    return P;
}

constexpr int MAX_REPAIR_ROUNDS = 64;
static size_t recur_spec_doubles(const b200m_handle *h, const Group &g, int nbands)
{
    const RecurParams P = recur_params(h, g, nbands, 0);
    const size_t lanes = (size_t)g.n_streams * P.tiles;
    const size_t pieces = (size_t)g.n_streams * (size_t)((g.max_stream_frames + 1023) / 1024);      // k_comp_sprint's 1024-frame pieces
    return 4 * lanes * nbands /* ss / se ping-pong */ + (lanes + 1) / 2 /* dirty list (uint32) */ + MAX_REPAIR_ROUNDS / 2 + 2 /* one counter per round */ +
           (pieces + 1) / 2 /* list of pieces (uint32) */ + (pieces + 63) / 64 + 2 /* their bitmap */;
}

// debug_att: the per-frame attenuation trajectory is materialised too (b200m_compress_dynamic_range's att_out)
static size_t compressor_ws_bytes(const b200m_handle *h, const Group &g, int64_t F, int nbands, bool debug_att = false)
{
    const size_t per_band = (size_t)F * 2 /*rms*/ + (debug_att ? (size_t)F * 8 : 0) + (size_t)(g.total_blocks + 1) * (4 /*hold*/ + 32 * 8 /*bend*/) + 4 * 256;
    return nbands * per_band + recur_spec_doubles(h, g, nbands) * 8 + 256;
}

// rms / hold / bend (/ att) for `nbands` bands starting at band_base, and the speculation scratch
static double *take_compressor_ws(b200m_handle *h, Arena &A, const Group &g, int64_t F, int nbands, int band_base, BandPtrs &bp, bool debug_att = false)
{
    for (int b = band_base; b < band_base + nbands; ++b) bp.rms[b] = A.take<uint16_t>(F);
    for (int b = band_base; b < band_base + nbands; ++b) bp.att[b] = debug_att ? A.take<double>(F) : nullptr;
    for (int b = band_base; b < band_base + nbands; ++b) bp.hold[b] = A.take<uint32_t>(g.total_blocks + 1);
    for (int b = band_base; b < band_base + nbands; ++b) bp.bend[b] = A.take<double>((size_t)(g.total_blocks + 1) * 32);
    return A.take<double>(recur_spec_doubles(h, g, nbands));
}

static int launch_compressor(b200m_handle *h, const Group &g, const BandPtrs &bp, int nbands, int band_base, int16_t *d_proc, double *d_spec)
{
    if (!((nbands == 3 && band_base == 0) || nbands == 1))
        return fail(h, B200M_ERR_INVALID, "internal: compressor runs on one band or on all three");
    if (g.max_look <= DW_MAX_LOOK && h->detect_kernel != 1) {
        const dim3 gw((g.max_stream_frames + DW_SPAN * DW_WARPS - 1) / (DW_SPAN * DW_WARPS), g.n_streams, nbands);
        if (g.ch == 2) LAUNCH("k_detect", k_detectw<2><<<gw, 32 * DW_WARPS, 0, h->stream>>>(g.d_streams, h->d_plans, bp, band_base));
        else           LAUNCH("k_detect", k_detectw<1><<<gw, 32 * DW_WARPS, 0, h->stream>>>(g.d_streams, h->d_plans, bp, band_base));
    } else {
        const dim3 gd((g.max_stream_frames + DT - 1) / DT, g.n_streams, nbands);
        const size_t smem = detect_smem_bytes(g.max_look);
        if (g.ch == 2) LAUNCH("k_detect", k_detect<2><<<gd, DNT, smem, h->stream>>>(g.d_streams, h->d_plans, bp, band_base));
        else           LAUNCH("k_detect", k_detect<1><<<gd, DNT, smem, h->stream>>>(g.d_streams, h->d_plans, bp, band_base));
    }
    const RecurParams P = recur_params(h, g, nbands, band_base);
    const size_t lanes = (size_t)g.n_streams * P.tiles;              // one lane per (stream, tile), all bands
    const size_t slots = lanes * nbands;
    RecurParams P0 = P;
    P0.mode = 0;
    double *ss[2] = {d_spec, d_spec + 2 * slots}, *se[2] = {d_spec + slots, d_spec + 3 * slots};
    const unsigned gr = (unsigned)((lanes + 31) / 32);
#define LAUNCH_COMP(NAME, PP, SSI, SEI, SSO, SEO, DL, DC) do { \
        if (nbands == 3) { if (g.ch == 2) LAUNCH(NAME, k_comp<2, 3, false><<<gr, 96, recur_smem_bytes(3), h->stream>>>(g.d_streams, h->d_plans, PP, bp, d_proc, SSI, SEI, SSO, SEO, DL, DC)); \
                           else           LAUNCH(NAME, k_comp<1, 3, false><<<gr, 96, recur_smem_bytes(3), h->stream>>>(g.d_streams, h->d_plans, PP, bp, d_proc, SSI, SEI, SSO, SEO, DL, DC)); } \
        else             { if (g.ch == 2) LAUNCH(NAME, k_comp<2, 1, true><<<gr, 32, recur_smem_bytes(1), h->stream>>>(g.d_streams, h->d_plans, PP, bp, d_proc, SSI, SEI, SSO, SEO, DL, DC)); \
                           else           LAUNCH(NAME, k_comp<1, 1, true><<<gr, 32, recur_smem_bytes(1), h->stream>>>(g.d_streams, h->d_plans, PP, bp, d_proc, SSI, SEI, SSO, SEO, DL, DC)); } } while (0)
    unsigned *d_list = reinterpret_cast<unsigned *>(d_spec + 4 * slots);
    unsigned *d_counts = d_list + ((lanes + 1) / 2) * 2;             // [MAX_REPAIR_ROUNDS]
    LAUNCH_COMP("k_comp", P0, nullptr, nullptr, ss[0], se[0], nullptr, nullptr);
    int cur = 0;
    if (P.tiles > 1 && h->recur_rounds != 0) {
        RecurParams P1 = P;
        P1.mode = 1;
        // a stretch in which trajectories from different pasts stay apart (a slow release: up to ~2 release times)
        // may span several short tiles, and a Jacobi round carries the truth one tile further: short tiles get more
        // rounds (an empty round costs two empty launches)
        const int auto_rounds = std::max(6, std::min(24, 2 + 40000 / std::max(1, P.tile_len)));
        const int rounds = std::min(h->recur_rounds >= 0 ? h->recur_rounds : auto_rounds, MAX_REPAIR_ROUNDS);
        CK(cudaMemsetAsync(d_counts, 0, MAX_REPAIR_ROUNDS * sizeof(unsigned), h->stream));
        const unsigned gdirty = (unsigned)((lanes + 255) / 256);
        if (h->recur_rounds < 0 && h->comp_sprint) {
            // The true state first, the samples after: k_comp_sprint walks the stretches whose guess was wrong with the
            // recurrence alone (sequential per stream and band, tiles whose guess was right are skipped), leaves the
            // true state at the end of every 32-frame block and lists the 1024-frame pieces whose samples change;
            // k_comp (mode 2) then recomputes those pieces, all at once, each from its own true state.
            const int fine_tiles = (g.max_stream_frames + 1023) / 1024;
            const size_t pieces = (size_t)g.n_streams * fine_tiles;
            unsigned *d_pieces = d_counts + MAX_REPAIR_ROUNDS + 2;
            unsigned *d_bitmap = d_pieces + ((pieces + 1) / 2) * 2;
            CK(cudaMemsetAsync(d_bitmap, 0, ((pieces + 63) / 64) * 8, h->stream));
            const unsigned gs = (unsigned)(((size_t)g.n_streams * nbands + 3) / 4);
            if (nbands == 3) LAUNCH("k_comp_sprint", k_comp_sprint<3><<<gs, 128, 0, h->stream>>>(g.d_streams, h->d_plans, P1, bp, ss[0], se[0], ss[1], se[1], d_pieces, d_counts, d_bitmap, fine_tiles, h->d_counters));
            else             LAUNCH("k_comp_sprint", k_comp_sprint<1><<<gs, 128, 0, h->stream>>>(g.d_streams, h->d_plans, P1, bp, ss[0], se[0], ss[1], se[1], d_pieces, d_counts, d_bitmap, fine_tiles, h->d_counters));
            RecurParams P2 = P;
            P2.mode = 2; P2.tile_len = 1024; P2.tiles = fine_tiles;
            {
                const unsigned gr_keep = gr;
                const unsigned gr = (unsigned)((pieces + 31) / 32);      // upper bound: CTAs beyond the list leave at once
                (void)gr_keep;
                LAUNCH_COMP("k_comp_repair", P2, nullptr, nullptr, nullptr, nullptr, d_pieces, d_counts);
            }
            cur = 1;                                                     // the joints are consistent by construction: (ss[1], se[1])
        } else
        for (int round = 0; round < rounds; ++round) {               // parallel repair rounds (Jacobi), dirty tiles only
            if (nbands == 3) LAUNCH("k_comp_dirty", k_comp_dirty<3><<<gdirty, 256, 0, h->stream>>>(g.d_streams, h->d_plans, P1, ss[cur], se[cur], ss[cur ^ 1], se[cur ^ 1], d_list, d_counts + round, h->d_counters));
            else             LAUNCH("k_comp_dirty", k_comp_dirty<1><<<gdirty, 256, 0, h->stream>>>(g.d_streams, h->d_plans, P1, ss[cur], se[cur], ss[cur ^ 1], se[cur ^ 1], d_list, d_counts + round, h->d_counters));
            LAUNCH_COMP("k_comp_repair", P1, ss[cur], se[cur], ss[cur ^ 1], se[cur ^ 1], d_list, d_counts + round);
            cur ^= 1;
        }
    }
#undef LAUNCH_COMP
    const unsigned gfx = (unsigned)((g.n_streams + 3) / 4);
    if (nbands == 3) { if (g.ch == 2) LAUNCH("k_comp_fix", k_comp_fix<2, 3><<<gfx, 128, 0, h->stream>>>(g.d_streams, h->d_plans, P, bp, d_proc, ss[cur], se[cur], h->d_counters));
                       else           LAUNCH("k_comp_fix", k_comp_fix<1, 3><<<gfx, 128, 0, h->stream>>>(g.d_streams, h->d_plans, P, bp, d_proc, ss[cur], se[cur], h->d_counters)); }
    else             { if (g.ch == 2) LAUNCH("k_comp_fix", k_comp_fix<2, 1><<<gfx, 128, 0, h->stream>>>(g.d_streams, h->d_plans, P, bp, d_proc, ss[cur], se[cur], h->d_counters));
                       else           LAUNCH("k_comp_fix", k_comp_fix<1, 1><<<gfx, 128, 0, h->stream>>>(g.d_streams, h->d_plans, P, bp, d_proc, ss[cur], se[cur], h->d_counters)); }
    CK(cudaGetLastError());
    return B200M_OK;
}

// The K-weighting tables of the current plans, by value, if every plan with a loudness target has the same ones
// (they depend on the rate alone, and a batch has one rate).
static bool kw_tables(const b200m_handle *h, KwTabsC &kt)
{
    const PlanDev *first = nullptr;
    for (const PlanDev &p : h->plans_host) {
        if (!p.has_lufs) continue;
        if (!first) first = &p;
        else if (std::memcmp(first->kw, p.kw, sizeof p.kw) != 0) return false;
    }
    if (!first) return false;
    fill_tabc(kt.sec[0], first->kw[0]);
    fill_tabc(kt.sec[1], first->kw[1]);
    return true;
}

static int launch_loudness(b200m_handle *h, const Group &g, const int16_t *d_proc, const float *d_mono,
                           float *d_kw, double *d_z, double *d_zsel, double2 *d_loud)
{
    if (g.any_lufs) {
        KwTabsC kt;
        const bool pt = kw_tables(h, kt);
#define LAUNCH_KW(CHN, INT, SRC) do { \
        if (pt) LAUNCH("k_kweight", k_kweight<CHN, INT, true><<<g.n_ksegs, KNT, kweight_smem_bytes(), h->stream>>>(SRC, g.d_tracks, g.d_ksegs, h->d_plans, d_kw, kt)); \
        else    LAUNCH("k_kweight", k_kweight<CHN, INT, false><<<g.n_ksegs, KNT, kweight_smem_bytes(), h->stream>>>(SRC, g.d_tracks, g.d_ksegs, h->d_plans, d_kw, kt)); } while (0)
        if (g.kw_warps && pt && !d_mono) {
            if (g.ch == 2) LAUNCH("k_kweight", k_kweightw<2><<<g.n_ksegs, 32, KwW<2>::SMEM, h->stream>>>(d_proc, g.d_tracks, g.d_ksegs, h->d_plans, d_kw, kt));
            else           LAUNCH("k_kweight", k_kweightw<1><<<g.n_ksegs, 32, KwW<1>::SMEM, h->stream>>>(d_proc, g.d_tracks, g.d_ksegs, h->d_plans, d_kw, kt));
        }
        else if (d_mono)    LAUNCH_KW(1, float, d_mono);
        else if (g.ch == 2) LAUNCH_KW(2, int16_t, d_proc);
        else                LAUNCH_KW(1, int16_t, d_proc);
#undef LAUNCH_KW
        const dim3 gb(g.max_blocks, g.n_tracks);
        // hop sums go to d_zsel (k_gate's scratch, free until then): nblocks + 3 floats in nblocks doubles per track
        const bool hops = h->hops_smem_floats > 0 && g.max_blocks >= 3;
        const dim3 gh(g.max_blocks + 3, g.n_tracks);
        if (hops) LAUNCH("k_hops", k_hops<<<gh, BNT, hops_smem_bytes(h), h->stream>>>(d_kw, g.d_tracks, h->d_plans, d_zsel, h->hops_stage_floats));
        if (g.max_blocks > 0) LAUNCH("k_blocks", k_blocks<<<gb, BNT, (size_t)h->blocks_smem_floats * 4 + 16, h->stream>>>(d_kw, g.d_tracks, h->d_plans, hops ? d_zsel : nullptr, d_z));
    }
    LAUNCH("k_gate", k_gate<<<g.n_tracks, GNT, 0, h->stream>>>(g.d_tracks, h->d_plans, d_z, d_zsel, d_loud));
    CK(cudaGetLastError());
    return B200M_OK;
}

static int num_blocks(int64_t frames, int rate)
{
    // pyloudnorm: T = numSamples / rate; numBlocks = int(np.round((T - T_g) / (T_g * step)) + 1)
    const double T = (double)frames / (double)rate, Tg = 0.4, step = 0.25;
    const double nb = std::nearbyint((T - Tg) / (Tg * step)) + 1;
    return nb < 0 ? 0 : (int)nb;
}

// ------------------------------------------------------------------------------------
// b200m_master_batch
//
// The batch is cut into *groups* of whole tracks.  A group is planned on the host
// (plan_group: stream / track / segment descriptors and its workspace need) and executed
// from a workspace slot (exec_group).  With device-resident PCM there is one slot and
// everything runs on the handle's stream.  With HOST buffers the groups are pipelined over
// three slots and three streams: the H2D copy of group g+1 and the D2H copy of group g-1
// overlap the kernels of group g (copy engines vs SMs), ordered by events.
// ------------------------------------------------------------------------------------
struct GroupPlan {
    int t_begin = 0, t_end = 0;
    Group g;
    std::vector<StreamDesc> streams;
    std::vector<TrackDesc> tracks;
    std::vector<SegDesc> csegs, ksegs;
    std::vector<int32_t> cta_iters;  // k_chainw's 16-warp shape: tiles every warp of CTA i walks
    int64_t F = 0, Fp = 0, in_total = 0, zoff = 0, out_base = 0;      // F: workspace frames (aligned track starts), Fp: packed output frames
    size_t desc_bytes = 0, res_off = 0, pin_bytes = 0, need = 0;
    int fmt = B200M_FMT_S16;         // PCM format of the caller's input (s24 / f32: staged to the 16-bit domain first, declared extension)
    int n_targets = 0;               // > 0: loudness sweep (b200m_master_batch_targets): that many outputs per track
    bool wav = false;                // b200m_master_batch_wav: a 44-byte RIFF header ahead of every track's samples
};

static void plan_group(const b200m_handle *h, GroupPlan &gp, bool in_dev, bool out_dev, int t_begin, int t_end,
                       const int64_t *in_offsets, const int64_t *in_frames, const int64_t *out_frames,
                       const b200m_plan *plans, const int32_t *plan_index, int64_t out_base, int n_targets = 0,
                       const int64_t *out_offsets = nullptr, int fmt = B200M_FMT_S16)
{
    gp.n_targets = n_targets;
    gp.fmt = fmt;
    const bool packed_in = !in_dev || fmt != B200M_FMT_S16;      // the kernels read a packed copy of the group's tracks (H2D staging or format staging)
    const int bps = fmt == B200M_FMT_S16 ? 2 : fmt == B200M_FMT_S24 ? 3 : 4;
    gp.wav = out_offsets != nullptr;
    // WAV images: the group's span of the output starts at the header of its first track
    const int64_t hdr = out_offsets ? 44 / (plans[0].channels * 2) : 0;
    if (out_offsets) out_base = out_offsets[t_begin] - hdr;
    const size_t outs = (size_t)std::max(1, n_targets);
    const int ch = plans[0].channels, rate = plans[0].sample_rate;
    // ENG:48-54 `audio[start_ms:start_ms + 30000]`: pydub turns a millisecond position into a frame index as
    // int(ms * (rate / 1000.0)) -- for about one integer rate in seven (11 024, 18 900, 37 800 ... Hz) the product for some
    // chunk k lands one frame short of 30 * rate * k, and filters and compressors restart there, so chunk k starts at
    // the frame pydub computes, operation for operation (all the usual rates are exact either way).
    const double frames_per_ms = (double)rate / 1000.0;
    auto chunk_start = [&](int64_t k) { return (int64_t)((double)(30000 * k) * frames_per_ms); };
    Group &g = gp.g;
    gp.t_begin = t_begin; gp.t_end = t_end; gp.out_base = out_base;
    g.ch = ch;
    g.n_tracks = t_end - t_begin;
    gp.tracks.resize(g.n_tracks);
    int64_t F = 0, Fp = 0, in_total = 0, zoff = 0;
    for (int t = t_begin; t < t_end; ++t) {
        const b200m_plan &p = plans[plan_index[t]];
        TrackDesc &td = gp.tracks[t - t_begin];
        F = (F + 31) & ~(int64_t)31;            // workspace rows of the compressor are moved in aligned 4-byte pieces
        td.off = F; td.dst_off = out_offsets ? out_offsets[t] - out_base : Fp; td.frames = out_frames[t]; td.plan = plan_index[t];
        td.abs0 = 0; td.total_frames = out_frames[t]; td.j0 = 0; td.pad_ = 0;
        td.nblocks = p.has_lufs ? num_blocks(out_frames[t], rate) : 0;
        td.zoff = zoff; zoff += td.nblocks;
        g.max_blocks = std::max(g.max_blocks, td.nblocks);
        g.max_track_frames = std::max(g.max_track_frames, td.frames);
        g.any_multiband |= p.multiband != 0;
        g.any_lufs |= p.has_lufs != 0;
        if (!(chain_warm_frames(p) < 1e29)) g.chain_stable = false;
        if (p.multiband) for (int b = 0; b < 3; ++b) g.max_look = std::max(g.max_look, p.band[b].look_frames);
        const int64_t in_base = packed_in ? in_total : in_offsets[t];
        for (int64_t k = 0; chunk_start(k) < out_frames[t]; ++k) {
            const int64_t s = chunk_start(k);
            StreamDesc sd;
            sd.in_off = in_base + s; sd.out_off = F + s;
            sd.out_frames = (int32_t)(std::min(chunk_start(k + 1), out_frames[t]) - s);
            sd.in_frames = (int32_t)std::max<int64_t>(0, std::min<int64_t>(sd.out_frames, in_frames[t] - s));
            sd.plan = plan_index[t]; sd.track = t - t_begin;
            sd.blk_off = (int32_t)g.total_blocks; sd.pad_ = 0;
            g.total_blocks += (sd.out_frames + 1023) / 1024;
            g.max_stream_frames = std::max(g.max_stream_frames, sd.out_frames);
            gp.streams.push_back(sd);
        }
        F += out_frames[t];
        Fp = out_offsets ? out_offsets[t] - out_base + out_frames[t] : Fp + out_frames[t];
        in_total += in_frames[t];
    }
    g.n_streams = (int)gp.streams.size();
    g.single_plan = gp.streams.empty() ? -1 : gp.streams[0].plan;
    for (auto &sd : gp.streams) if (sd.plan != g.single_plan) g.single_plan = -1;
    {
        int64_t ctiles = 0, ktiles = 0;
        for (auto &sd : gp.streams) ctiles += (sd.out_frames + TILE - 1) / TILE;
        for (auto &td : gp.tracks) ktiles += (td.frames + KTILE - 1) / KTILE;
        const int cs = h->seg_chain == 0 ? auto_seg_tiles(ctiles, 8, 48) : h->seg_chain;
        const int ks = h->seg_kweight == 0 ? auto_seg_tiles(ktiles, 8, 32) : h->seg_kweight;
        // k_chainw: every warp walks its own segment; 148 SMs x 16 warps are resident at a time and all segments of a
        // group have (nearly) the same length, so the kernel's time is waves x (segment + warm-up).  The cut is chosen
        // to minimise exactly that: for w = 1 .. chain_waves waves, the longest segments whose count still fits w
        // resident waves (a few segments beyond a wave cost a whole extra pass: 600 CTAs of 16 warps on 148 SMs ran
        // 9.4 ms per 64 tracks, 144 CTAs run 7.3 ms), never shorter than four warm-ups; smaller groups stay with
        // k_chain, whose eight warps share one warm-up.
        double max_warm = 0;
        bool bounded = true;
        for (auto &sd : gp.streams) {
            const double w = chain_warm_frames(plans[sd.plan]);
            if (!(w < 1e6)) bounded = false;
            max_warm = std::max(max_warm, w);
        }
        const int wt = ch == 2 ? ChainW<2>::WT : ChainW<1>::WT;
        std::vector<int64_t> wtiles(gp.streams.size());
        int64_t total_wtiles = 0;
        for (size_t i = 0; i < gp.streams.size(); ++i) { wtiles[i] = (gp.streams[i].out_frames + wt - 1) / wt; total_wtiles += wtiles[i]; }
        const double warm_tiles = std::ceil(std::max(max_warm, 1.0) / wt);
        const double min_len = std::max(8.0, 4.0 * std::max(warm_tiles, 1.0));      // segment length in warp tiles
        double best_len = bounded ? plan_warp_segments(wtiles, warm_tiles, min_len, h->chain_waves, 148.0 * 16) : 0;
        const bool enough = (double)total_wtiles / min_len >= 148.0 * 16;     // a resident wave of warps, each with a run of >= 4 warm-ups
        g.chain_warps = bounded && h->seg_chain >= 0 && (h->chain_kernel == 2 || (h->chain_kernel == 0 && enough));
        if (best_len == 0) best_len = min_len;
        if (g.chain_warps) {
            const int wseg = h->seg_chain > 0 ? h->seg_chain * (TILE / wt) : (int)std::min<double>(1 << 20, best_len);
            for (size_t i = 0; i < gp.streams.size(); ++i)
                make_segments_w(gp.csegs, (int)i, gp.streams[i].out_frames, wt, chain_warm_frames(plans[gp.streams[i].plan]), wseg);
            // every plan has the exciter on with the SAME odd table: the 16-warp shape keeps the table in shared memory
            // (one plan: its filter tables travel as a kernel parameter; several: the segments are grouped by plan, a
            // CTA's sixteen belong to one plan and its tables sit in shared memory next to the exciter table)
            g.chain_slut = h->chain_slut_ok && !gp.streams.empty();
            const float *lut0 = nullptr;
            for (auto &sd : gp.streams) {
                if ((size_t)sd.plan >= h->plans_host.size()) { g.chain_slut = false; break; }
                const PlanDev &pd = h->plans_host[sd.plan];
                if (!pd.sat_on || !pd.sat_sym || (lut0 && pd.sat_lut != lut0)) { g.chain_slut = false; break; }
                lut0 = pd.sat_lut;
            }
            if (g.chain_slut) {
                if (g.single_plan < 0) {
                    std::vector<SegDesc> grouped;
                    std::vector<int> plan_ids;
                    for (auto &sd : gp.streams) if (std::find(plan_ids.begin(), plan_ids.end(), sd.plan) == plan_ids.end()) plan_ids.push_back(sd.plan);
                    for (int pid : plan_ids) {
                        for (auto &sg : gp.csegs) if (gp.streams[sg.owner].plan == pid) grouped.push_back(sg);
                        int owner0 = 0;
                        for (size_t i = 0; i < gp.streams.size(); ++i) if (gp.streams[i].plan == pid) { owner0 = (int)i; break; }
                        while (grouped.size() % CW_SLUT != 0) grouped.push_back({0, 0, owner0, 0});     // padding names a stream of the same plan
                    }
                    gp.csegs.swap(grouped);
                }
                while (gp.csegs.size() % CW_SLUT != 0) gp.csegs.push_back({0, 0, gp.csegs.empty() ? 0 : gp.csegs.back().owner, 0});
                for (size_t c0 = 0; c0 < gp.csegs.size(); c0 += CW_SLUT) {
                    int64_t it = 0;
                    for (int w = 0; w < CW_SLUT; ++w) {
                        const SegDesc &sg = gp.csegs[c0 + w];
                        const int64_t first = std::max<int64_t>(0, sg.begin - sg.warm);
                        it = std::max(it, (sg.end - first + wt - 1) / wt);
                    }
                    gp.cta_iters.push_back((int32_t)it);
                }
            }
        } else {
            for (size_t i = 0; i < gp.streams.size(); ++i)
                make_segments(gp.csegs, (int)i, gp.streams[i].out_frames, TILE, chain_warm_frames(plans[gp.streams[i].plan]), cs);
        }
        // k_kweightw: the same idea for the K-weighting (one warp per segment of a track, 148 x B200M_KWW_OCC resident)
        {
            constexpr int kwt = KwW<2>::WT;
            std::vector<int64_t> ktl(gp.tracks.size(), 0);
            int64_t ktotal = 0;
            double kwarm = 0;
            bool kbounded = true;
            for (size_t i = 0; i < gp.tracks.size(); ++i) {
                if (!plans[gp.tracks[i].plan].has_lufs) continue;
                ktl[i] = (gp.tracks[i].frames + kwt - 1) / kwt; ktotal += ktl[i];
                const double w = kweight_warm_frames(plans[gp.tracks[i].plan]);
                if (!(w < 1e6)) kbounded = false;
                kwarm = std::max(kwarm, w);
            }
            const double kwarm_tiles = std::ceil(std::max(kwarm, 1.0) / kwt);
            const double kmin_len = std::max(8.0, 4.0 * kwarm_tiles);
            const double resident = 148.0 * B200M_KWW_OCC;
            const bool kenough = (double)ktotal / kmin_len >= resident;
            g.kw_warps = kbounded && h->seg_kweight >= 0 && (h->kweight_kernel == 2 || (h->kweight_kernel == 0 && kenough));
            if (g.kw_warps) {
                double len = h->seg_kweight > 0 ? (double)h->seg_kweight * (KTILE / kwt) : plan_warp_segments(ktl, kwarm_tiles, kmin_len, h->chain_waves, resident);
                if (len == 0) len = kmin_len;
                for (size_t i = 0; i < gp.tracks.size(); ++i)
                    if (plans[gp.tracks[i].plan].has_lufs && gp.tracks[i].frames > 0)
                        make_segments_w(gp.ksegs, (int)i, gp.tracks[i].frames, kwt, kweight_warm_frames(plans[gp.tracks[i].plan]), (int)std::min<double>(1 << 20, len));
            } else {
                for (size_t i = 0; i < gp.tracks.size(); ++i)
                    if (plans[gp.tracks[i].plan].has_lufs)
                        make_segments(gp.ksegs, (int)i, gp.tracks[i].frames, KTILE, kweight_warm_frames(plans[gp.tracks[i].plan]), ks);
            }
        }
        g.n_csegs = (int)gp.csegs.size(); g.n_ksegs = (int)gp.ksegs.size();
    }
    gp.F = F; gp.Fp = Fp; gp.in_total = in_total; gp.zoff = zoff;
    gp.desc_bytes = gp.streams.size() * sizeof(StreamDesc) + gp.tracks.size() * sizeof(TrackDesc) +
                    (gp.csegs.size() + gp.ksegs.size()) * sizeof(SegDesc) + gp.cta_iters.size() * sizeof(int32_t);
    gp.res_off = (gp.desc_bytes + 63) & ~(size_t)63;       // double2 results: 16-byte aligned slot after the descriptors
    gp.pin_bytes = (gp.res_off + (size_t)g.n_tracks * 16 * (1 + (size_t)n_targets) + 255) & ~(size_t)255;
    size_t need = 16384 + gp.desc_bytes + (size_t)g.n_tracks * 16 * (1 + (size_t)n_targets) + (size_t)F * ch * 2 /*proc*/ + (size_t)F * 4 /*kw*/ +
                  (size_t)zoff * 16 + 25 * 256;
    if (!in_dev) need += (size_t)in_total * ch * bps + 256;
    if (fmt != B200M_FMT_S16) need += (size_t)in_total * ch * 2 + 256;
    if (!out_dev) need += (size_t)Fp * ch * 2 * outs + 256 * outs + 256;
    if (g.any_multiband) need += (size_t)F * 3 * ch * 2 + 3 * 256 + compressor_ws_bytes(h, g, F, 3);
    gp.need = (need + 1023) & ~(size_t)1023;
}

struct ExecStreams {
    cudaStream_t in, comp, out;             // equal when the group is not pipelined
    cudaEvent_t slot_free, h2d_done, comp_done, d2h_done;   // null when not pipelined
};

// ext_proc != NULL: stop after the chunk-wise part (ENG:48-80) and leave `processed_audio` there
// (time slices of a long track: loudness is measured across slices, b200m_slice_*)
static int exec_group(b200m_handle *h, GroupPlan &gp, char *ws, char *pin, const ExecStreams &X,
                      const int16_t *pcm_in, bool in_dev, const int64_t *in_offsets, const int64_t *in_frames,
                      int16_t *pcm_out, bool out_dev, int16_t *ext_proc = nullptr,
                      const double *targets = nullptr, int64_t out_total = 0)
{
    // targets != NULL (gp.n_targets of them): k_final runs once per target; output copy k of the whole batch
    // starts out_total frames after copy k - 1
    const int n_out = targets ? gp.n_targets : 1;
    Group &g = gp.g;
    const int ch = g.ch;
    const int64_t F = gp.F;
    if (F == 0) return B200M_OK;
    Arena A(ws);
    StreamDesc *d_streams = A.take<StreamDesc>(gp.streams.size());
    TrackDesc *d_tracks = A.take<TrackDesc>(gp.tracks.size());
    SegDesc *d_csegs = A.take<SegDesc>(gp.csegs.size() + 1);
    SegDesc *d_ksegs = A.take<SegDesc>(gp.ksegs.size() + 1);
    int32_t *d_cta_iters = A.take<int32_t>(gp.cta_iters.size() + 1);
    double2 *d_loud = A.take<double2>((size_t)g.n_tracks * (1 + (size_t)gp.n_targets));     // [0]: k_gate's, [1 + k]: target k
    int16_t *d_proc = ext_proc ? ext_proc : A.take<int16_t>((size_t)F * ch);
    float *d_kw = A.take<float>(F);
    double *d_z = A.take<double>(gp.zoff + 1);
    double *d_zsel = A.take<double>(gp.zoff + 1);
    int16_t *d_in = nullptr, *d_out = nullptr;
    const int fmt = gp.fmt, bps = fmt == B200M_FMT_S16 ? 2 : fmt == B200M_FMT_S24 ? 3 : 4;
    unsigned char *d_raw = nullptr;                      // host input: the group's tracks in the caller's format, packed
    if (!in_dev) d_raw = A.take<unsigned char>((size_t)gp.in_total * ch * bps + 16);
    if (fmt != B200M_FMT_S16) d_in = A.take<int16_t>((size_t)gp.in_total * ch);     // ... and staged to the 16-bit domain
    else d_in = reinterpret_cast<int16_t *>(d_raw);
    const size_t out_stride = ((size_t)gp.Fp * ch + 127) & ~(size_t)127;                    // samples between the staged copies (256-byte aligned)
    if (!out_dev) {
        d_out = A.take<int16_t>(out_stride * n_out + 8);
        // WAV images: the staged span starts at a header, i.e. anywhere; keep the device address congruent to the
        // host offset modulo 16 bytes so that samples the caller aligned stay aligned for k_final's 16-byte stores
        if (gp.wav) d_out += ((gp.out_base * ch * 2) & 15) / 2;
    }
    BandPtrs bp;
    double *d_spec = nullptr;
    std::memset(&bp, 0, sizeof bp);
    if (g.any_multiband) {
        for (int b = 0; b < 3; ++b) bp.band[b] = A.take<int16_t>((size_t)F * ch);
        d_spec = take_compressor_ws(h, A, g, F, 3, 0, bp);
    }
    if (A.used > gp.need) return fail(h, B200M_ERR_NOMEM, "internal: workspace estimate too small (%zu > %zu)", A.used, gp.need);

    // ---- input stream: descriptors (pinned staging) and PCM -> device ----------------------
    if (X.slot_free) CK(cudaStreamWaitEvent(X.in, X.slot_free, 0));    // the slot's previous group has left it
    {
        char *pp = pin;
        auto up = [&](void *dst, const void *src, size_t bytes) -> cudaError_t {
            if (!bytes) return cudaSuccess;
            std::memcpy(pp, src, bytes);
            cudaError_t e = cudaMemcpyAsync(dst, pp, bytes, cudaMemcpyHostToDevice, X.in);
            pp += bytes;
            return e;
        };
        CK(up(d_streams, gp.streams.data(), gp.streams.size() * sizeof(StreamDesc)));
        CK(up(d_tracks, gp.tracks.data(), gp.tracks.size() * sizeof(TrackDesc)));
        CK(up(d_csegs, gp.csegs.data(), gp.csegs.size() * sizeof(SegDesc)));
        CK(up(d_ksegs, gp.ksegs.data(), gp.ksegs.size() * sizeof(SegDesc)));
        CK(up(d_cta_iters, gp.cta_iters.data(), gp.cta_iters.size() * sizeof(int32_t)));
    }
    g.d_streams = d_streams; g.d_tracks = d_tracks; g.d_csegs = d_csegs; g.d_ksegs = d_ksegs;
    const int16_t *d_src = pcm_in;
    if (!in_dev) {
        int64_t pos = 0;
        for (int t = gp.t_begin; t < gp.t_end;) {       // one copy per run of tracks that are contiguous in the caller's buffer
            int t2 = t + 1;
            int64_t run = in_frames[t];
            while (t2 < gp.t_end && in_offsets[t2] == in_offsets[t2 - 1] + in_frames[t2 - 1]) { run += in_frames[t2]; ++t2; }
            if (run > 0)
                CK(cudaMemcpyAsync(d_raw + (size_t)pos * ch * bps, reinterpret_cast<const unsigned char *>(pcm_in) + (size_t)in_offsets[t] * ch * bps,
                                   (size_t)run * ch * bps, cudaMemcpyHostToDevice, X.in));
            pos += run;
            t = t2;
        }
        d_src = d_in;
    }
    if (X.h2d_done) { CK(cudaEventRecord(X.h2d_done, X.in)); CK(cudaStreamWaitEvent(X.comp, X.h2d_done, 0)); }
    if (fmt != B200M_FMT_S16) {
        // declared extension (b200m_stage_pcm): s24 keeps its high-order 16 bits, f32 goes through ENG:123-126
        auto stage = [&](const unsigned char *src, int64_t n_samples, int16_t *dst) {
            if (n_samples <= 0) return;
            const int grid = (int)std::min<int64_t>((n_samples / 4 + 255) / 256 + 1, 148 * 16);
            if (fmt == B200M_FMT_S24) LAUNCH("k_stage_s24", k_stage_s24<<<grid, 256, 0, h->stream>>>(src, n_samples, dst));
            else                      LAUNCH("k_stage_f32", k_stage_f32<<<grid, 256, 0, h->stream>>>(reinterpret_cast<const float *>(src), n_samples, dst));
        };
        if (!in_dev) stage(d_raw, gp.in_total * ch, d_in);
        else {
            int64_t pos = 0;
            for (int t = gp.t_begin; t < gp.t_end; ++t) {
                stage(reinterpret_cast<const unsigned char *>(pcm_in) + (size_t)in_offsets[t] * ch * bps, in_frames[t] * ch, d_in + pos * ch);
                pos += in_frames[t];
            }
        }
        CK(cudaGetLastError());
        d_src = d_in;
    }
    int16_t *d_dst = out_dev ? pcm_out + gp.out_base * ch : d_out;

    // ---- kernels (all on the handle's stream) ------------------------------------------------
    if (g.chain_warps) {
        ChainTabsC ct;                           // 3.6 KB, copied into the launch parameter buffer by the runtime
        const bool pt = g.single_plan >= 0;
        if (pt) {
            const PlanDev &pd = h->plans_host[g.single_plan];
            const SecTab *src[8] = {&pd.eq[0], &pd.eq[1], &pd.eq[2], &pd.eq[3], &pd.lp[0], &pd.lp[1], &pd.hp[0], &pd.hp[1]};
            for (int s8 = 0; s8 < 8; ++s8) fill_tabc(ct.sec[s8], *src[s8]);
        }
        const bool nanchk = !g.chain_stable;
        if (g.chain_slut) {
            const int nb = g.n_csegs / CW_SLUT;
#define LAUNCH_CHAINS(CHN, NAN_, PT_) \
            LAUNCH("k_chain", k_chainw<CHN, NAN_, PT_, CW_SLUT, true><<<nb, 32 * CW_SLUT, (PT_) ? ChainW<CHN>::SMEM_SLUT : ChainW<CHN>::SMEM_SLUT_MP, h->stream>>>(d_src, d_streams, d_csegs, g.n_csegs, h->d_plans, d_proc, bp, ct, d_cta_iters))
            if (ch == 2) {
                if (pt) { if (nanchk) LAUNCH_CHAINS(2, true, true); else LAUNCH_CHAINS(2, false, true); }
                else    { if (nanchk) LAUNCH_CHAINS(2, true, false); else LAUNCH_CHAINS(2, false, false); }
            } else {
                if (pt) { if (nanchk) LAUNCH_CHAINS(1, true, true); else LAUNCH_CHAINS(1, false, true); }
                else    { if (nanchk) LAUNCH_CHAINS(1, true, false); else LAUNCH_CHAINS(1, false, false); }
            }
#undef LAUNCH_CHAINS
        } else {
            const int nb = g.n_csegs;
#define LAUNCH_CHAINW(CHN, NAN_, PT_) \
            LAUNCH("k_chain", k_chainw<CHN, NAN_, PT_><<<nb, 32, (PT_) ? ChainW<CHN>::SMEM_PT : ChainW<CHN>::SMEM, h->stream>>>(d_src, d_streams, d_csegs, g.n_csegs, h->d_plans, d_proc, bp, ct, nullptr))
            if (ch == 2) {
                if (pt) { if (nanchk) LAUNCH_CHAINW(2, true, true); else LAUNCH_CHAINW(2, false, true); }
                else    { if (nanchk) LAUNCH_CHAINW(2, true, false); else LAUNCH_CHAINW(2, false, false); }
            } else {
                if (pt) { if (nanchk) LAUNCH_CHAINW(1, true, true); else LAUNCH_CHAINW(1, false, true); }
                else    { if (nanchk) LAUNCH_CHAINW(1, true, false); else LAUNCH_CHAINW(1, false, false); }
            }
#undef LAUNCH_CHAINW
        }
    } else if (ch == 2) {
        if (g.chain_stable) LAUNCH("k_chain", k_chain<2, false><<<g.n_csegs, NSEG * 2, chain_smem_bytes<2>(), h->stream>>>(d_src, d_streams, d_csegs, h->d_plans, d_proc, bp));
        else                LAUNCH("k_chain", k_chain<2, true><<<g.n_csegs, NSEG * 2, chain_smem_bytes<2>(), h->stream>>>(d_src, d_streams, d_csegs, h->d_plans, d_proc, bp));
    } else {
        if (g.chain_stable) LAUNCH("k_chain", k_chain<1, false><<<g.n_csegs, NSEG * 1, chain_smem_bytes<1>(), h->stream>>>(d_src, d_streams, d_csegs, h->d_plans, d_proc, bp));
        else                LAUNCH("k_chain", k_chain<1, true><<<g.n_csegs, NSEG * 1, chain_smem_bytes<1>(), h->stream>>>(d_src, d_streams, d_csegs, h->d_plans, d_proc, bp));
    }
    CK(cudaGetLastError());
    int rc;
    if (g.any_multiband) { rc = launch_compressor(h, g, bp, 3, 0, d_proc, d_spec); if (rc) return rc; }
    if (ext_proc) return B200M_OK;
    rc = launch_loudness(h, g, d_proc, nullptr, d_kw, d_z, d_zsel, d_loud);
    if (rc) return rc;
    const dim3 gf((unsigned)std::min<int64_t>((g.max_track_frames + 2047) / 2048, 8192), g.n_tracks);   // two vectors of four frames per thread
    for (int k = 0; k < n_out; ++k) {
        const double2 *loud_k = d_loud;
        int16_t *dst_k = d_dst;
        if (targets) {
            double2 *lk = d_loud + (size_t)(1 + k) * g.n_tracks;
            LAUNCH("k_regain", k_regain<<<(g.n_tracks + 127) / 128, 128, 0, h->stream>>>(d_loud, g.n_tracks, targets[k], lk));
            loud_k = lk;
            dst_k = out_dev ? pcm_out + ((int64_t)k * out_total + gp.out_base) * ch : d_out + (size_t)k * out_stride;
        }
        if (ch == 2) LAUNCH("k_final", k_final<2><<<gf, 256, 0, h->stream>>>(d_proc, d_tracks, h->d_plans, loud_k, dst_k));
        else         LAUNCH("k_final", k_final<1><<<gf, 256, 0, h->stream>>>(d_proc, d_tracks, h->d_plans, loud_k, dst_k));
    }
    if (gp.wav) LAUNCH("k_wav_headers", k_wav_headers<<<(g.n_tracks + 127) / 128, 128, 0, h->stream>>>(d_tracks, h->d_plans, g.n_tracks, ch, d_dst));
    CK(cudaGetLastError());
    if (X.comp_done) { CK(cudaEventRecord(X.comp_done, X.comp)); CK(cudaStreamWaitEvent(X.out, X.comp_done, 0)); }

    // ---- output stream: PCM and {loudness, gain} -> host --------------------------------------
    if (!out_dev)
        for (int k = 0; k < n_out; ++k)
            CK(cudaMemcpyAsync(pcm_out + ((int64_t)k * out_total + gp.out_base) * ch, d_out + (size_t)k * out_stride, (size_t)gp.Fp * ch * 2,
                               cudaMemcpyDeviceToHost, X.out));
    CK(cudaMemcpyAsync(pin + gp.res_off, d_loud, (size_t)g.n_tracks * 16 * (1 + (size_t)gp.n_targets), cudaMemcpyDeviceToHost, X.out));
    if (X.d2h_done) CK(cudaEventRecord(X.d2h_done, X.out));
    return B200M_OK;
}

static cudaEvent_t sync_event(b200m_handle *h, size_t i)
{
    while (h->sync_events.size() <= i) {
        cudaEvent_t e = nullptr;
        cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
        h->sync_events.push_back(e);
    }
    return h->sync_events[i];
}

static int master_batch_impl(b200m_handle *h, const void *pcm_in, int in_on_device, int fmt, int n_tracks,
                             const int64_t *in_offsets, const int64_t *in_frames, const int64_t *out_frames,
                             const b200m_plan *plans, int n_plans, const int32_t *plan_index,
                             const double *targets, int n_targets, const int64_t *out_offsets,
                             void *pcm_out, int out_on_device, double *loudness_out, double *gain_out)
{
    if (!h) return B200M_ERR_INVALID;
    struct Range { Range() { nvtxRangePushA("b200m_master_batch"); } ~Range() { nvtxRangePop(); } } nvtx_range;
    if (!pcm_in || !pcm_out || n_tracks <= 0 || !in_offsets || !in_frames || !out_frames || !plans || n_plans <= 0 || !plan_index)
        return fail(h, B200M_ERR_INVALID, "b200m_master_batch: null or empty argument");
    if (fmt != B200M_FMT_S16 && fmt != B200M_FMT_S24 && fmt != B200M_FMT_F32) return fail(h, B200M_ERR_INVALID, "b200m_master_batch: unknown PCM format %d", fmt);
    CK(cudaSetDevice(h->device));
    const int ch = plans[0].channels, rate = plans[0].sample_rate;
    for (int i = 0; i < n_plans; ++i)
        if (plans[i].channels != ch || plans[i].sample_rate != rate || rate <= 0)
            return fail(h, B200M_ERR_INVALID, "all plans of a batch must share sample rate and channel count");
    int64_t total_frames = 0;
    for (int t = 0; t < n_tracks; ++t) {
        if (plan_index[t] < 0 || plan_index[t] >= n_plans) return fail(h, B200M_ERR_INVALID, "plan_index[%d] out of range", t);
        if (in_frames[t] < 0 || out_frames[t] < 0 || in_offsets[t] < 0) return fail(h, B200M_ERR_INVALID, "negative frame count for track %d", t);
        if (out_frames[t] >= ((int64_t)1 << 40)) return fail(h, B200M_ERR_INVALID, "track %d too long", t);
        // pyloudnorm valid_audio: data.shape[0] < block_size * rate -> ValueError
        if (plans[plan_index[t]].has_lufs && (double)out_frames[t] < 0.4 * rate)
            return fail(h, B200M_ERR_TOO_SHORT, "track %d: audio must have length greater than the block size (400 ms)", t);
        total_frames += std::max(out_frames[t], in_frames[t]);
        if (plans[plan_index[t]].multiband) {
            // ENG:210: pydub's overlay re-derives a chunk's length from its ROUNDED millisecond length.  At the usual
            // rates that is the chunk itself; at the few integer rates where 30 000 ms is not a whole number of frames as
            // pydub computes it (11 024, 18 900, 37 800 ... Hz) the reference drops or inserts a frame at every chunk seam,
            // shifting everything behind it.  That is not reproduced: such a track is refused, not mastered differently.
            const double fpm = (double)rate / 1000.0;
            for (int64_t k = 0;; ++k) {
                const int64_t s0 = (int64_t)((double)(30000 * k) * fpm);
                if (s0 >= out_frames[t]) break;
                const int64_t n = std::min((int64_t)((double)(30000 * (k + 1)) * fpm), out_frames[t]) - s0;
                const int64_t e = (int64_t)(std::nearbyint(1000.0 * ((double)n / (double)rate)) * fpm);
                if (e != n)
                    return fail(h, B200M_ERR_INVALID, "track %d: at %d Hz chunk %lld (%lld frames) is not a whole number of milliseconds as pydub counts them "
                                "(its overlay would re-frame it to %lld); the multiband stage is not supported at this sample rate -- resample first",
                                t, rate, (long long)k, (long long)n, (long long)e);
            }
        }
        if (targets && !plans[plan_index[t]].has_lufs)
            return fail(h, B200M_ERR_INVALID, "track %d: a loudness sweep needs a plan with a loudness target (has_lufs)", t);
    }
    int64_t out_total = 0;                      // frames of one packed copy of the batch output
    for (int t = 0; t < n_tracks; ++t) out_total += out_frames[t];
    if (out_offsets) {                          // WAV images: room for a header ahead of every track, tracks in ascending order
        const int64_t hdr = 44 / (ch * 2);
        for (int t = 0; t < n_tracks; ++t) {
            const int64_t prev_end = t ? out_offsets[t - 1] + out_frames[t - 1] : 0;
            if (out_offsets[t] < prev_end + hdr)
                return fail(h, B200M_ERR_INVALID, "out_offsets[%d]: needs 44 free bytes after the end of the previous track", t);
            if ((out_offsets[t] * ch * 2) % 4 != 0)
                return fail(h, B200M_ERR_INVALID, "out_offsets[%d]: samples (and with them the header) must start at a multiple of 4 bytes", t);
            if ((uint64_t)out_frames[t] * ch * 2 > 0xffffffffull - 36)
                return fail(h, B200M_ERR_INVALID, "track %d does not fit a RIFF file (4 GiB)", t);
        }
    }
    int rc = ensure_plans(h, plans, n_plans);
    if (rc) return rc;
    const bool in_dev = in_on_device != 0, out_dev = out_on_device != 0;
    const bool pipelined = h->pipeline && (!in_dev || !out_dev) && n_tracks > 1;
    const int slots = pipelined ? 3 : 1;        // a slot is busy for H2D + kernels + D2H of its group: three keep all engines fed

    // ---- cut the batch into groups ---------------------------------------------------------
    // a group must fit one workspace slot; with host buffers it is also at most ~1/8 of the batch
    // (and at least ~8 M frames) so that copies and kernels of neighbouring groups overlap
    const double per_frame = ch * 2 * (2 + std::max(1, n_targets)) + 4 + 3 * (ch * 2 + 2 + 0.26) + 2;
    const double slot_limit = (double)h->ws_limit / slots;
    // host buffers: ~1/8 of the batch per group, but no more than `pipe_max_frames` (about 16 three-minute tracks): the
    // first group's H2D and the last group's D2H are not overlapped by anything, so on a large batch smaller groups win.
    // device buffers: as few groups as the workspace limit allows, of equal size.
    double pipe_frames = 1e300;
    if (pipelined) pipe_frames = std::max<double>(8e6, std::min<double>((double)total_frames / (double)h->pipe_groups, h->pipe_max_frames));
    else {
        const double ngroups = std::ceil(per_frame * (double)total_frames / slot_limit);
        if (ngroups > 1) pipe_frames = (double)total_frames / ngroups * 1.02 + 1;
    }
    std::vector<GroupPlan> gps;
    {
        int t0 = 0;
        int64_t out_base = 0;
        while (t0 < n_tracks) {
            int t1 = t0;
            double bytes = 0, fr = 0;
            int64_t frames = 0;
            while (t1 < n_tracks) {
                const double f = (double)std::max(out_frames[t1], in_frames[t1]);
                if (t1 > t0 && (bytes + per_frame * f > slot_limit || fr + 0.5 * f > pipe_frames)) break;
                bytes += per_frame * f; fr += f; frames += out_frames[t1]; ++t1;
            }
            gps.emplace_back();
            plan_group(h, gps.back(), in_dev, out_dev, t0, t1, in_offsets, in_frames, out_frames, plans, plan_index, out_base, targets ? n_targets : 0,
                       out_offsets, fmt);
            out_base += frames;
            t0 = t1;
        }
    }
    size_t max_need = 0, pin_total = 0;
    for (auto &gp : gps) { max_need = std::max(max_need, gp.need); pin_total += gp.pin_bytes; }
    rc = ws_reserve(h, max_need * slots);
    if (rc) return rc;
    const int pb = (h->pin_turn ^= 1);
    if (!h->pin_ev[pb]) CK(cudaEventCreateWithFlags(&h->pin_ev[pb], cudaEventDisableTiming));
    CK(cudaEventSynchronize(h->pin_ev[pb]));    // pinned staging: the call that used this buffer last (two calls ago) is through
    if (pb == 0) {
        rc = pin_reserve(h, pin_total);
        if (rc) return rc;
    } else if (pin_total > h->pinB_cap) {
        CK(cudaStreamSynchronize(h->stream));
        if (h->pinB) cudaFreeHost(h->pinB);
        h->pinB = nullptr; h->pinB_cap = 0;
        CK(cudaMallocHost(&h->pinB, pin_total + 4096));
        h->pinB_cap = pin_total + 4096;
    }
    char *const pbuf = pb ? h->pinB : h->pin;
    if (pipelined && !h->s_in) {
        CK(cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&h->s_comp2, cudaStreamNonBlocking));
    }
    cudaEvent_t start_ev = nullptr;
    if (pipelined) {                            // the side streams start after whatever the caller queued on the handle's stream
        start_ev = sync_event(h, 0);
        CK(cudaEventRecord(start_ev, h->stream));
        CK(cudaStreamWaitEvent(h->s_in, start_ev, 0));
        CK(cudaStreamWaitEvent(h->s_out, start_ev, 0));
        CK(cudaStreamWaitEvent(h->s_comp2, start_ev, 0));
    }
    size_t pin_off = 0;
    std::vector<size_t> pin_offs(gps.size());
    for (size_t i = 0; i < gps.size(); ++i) {
        GroupPlan &gp = gps[i];
        pin_offs[i] = pin_off;
        ExecStreams X;
        // Neighbouring groups alternate between two compute streams: the latency-bound kernels of one
        // group (k_recur_*, the tails of k_chain) share the SMs with the next group's throughput kernels,
        // which a small group cannot fill on its own.  Every launch of exec_group goes to h->stream.
        cudaStream_t const own = h->stream;
        if (pipelined && h->pipe_streams > 1 && (i & 1)) h->stream = h->s_comp2;
        if (pipelined) {
            X.in = h->s_in; X.comp = h->stream; X.out = h->s_out;
            X.h2d_done = sync_event(h, 1 + 3 * i); X.comp_done = sync_event(h, 2 + 3 * i); X.d2h_done = sync_event(h, 3 + 3 * i);
            X.slot_free = i >= (size_t)slots ? sync_event(h, 3 + 3 * (i - slots)) : nullptr;      // D2H of the slot's previous group
        } else {
            X.in = X.comp = X.out = h->stream;
            X.slot_free = X.h2d_done = X.comp_done = X.d2h_done = nullptr;
        }
        rc = exec_group(h, gp, h->ws + (i % slots) * max_need, pbuf + pin_off, X, (const int16_t *)pcm_in, in_dev, in_offsets,
                        in_frames, (int16_t *)pcm_out, out_dev, nullptr, targets, out_total);
        h->stream = own;
        if (rc) break;
        pin_off += gp.pin_bytes;
    }
    if (pipelined) {                            // the handle's stream observes the completion of everything
        cudaEvent_t done = sync_event(h, 1 + 3 * gps.size());
        cudaEventRecord(done, h->s_out);
        cudaStreamWaitEvent(h->stream, done, 0);
        cudaEvent_t done_in = sync_event(h, 2 + 3 * gps.size());
        cudaEventRecord(done_in, h->s_in);
        cudaStreamWaitEvent(h->stream, done_in, 0);
        cudaEvent_t done_c2 = sync_event(h, 3 + 3 * gps.size());
        cudaEventRecord(done_c2, h->s_comp2);
        cudaStreamWaitEvent(h->stream, done_c2, 0);
    }
    cudaEventRecord(h->pin_ev[pb], h->stream);
    if (rc) { cudaStreamSynchronize(h->stream); return rc; }
    if (loudness_out || gain_out || !out_dev) {
        CK(cudaStreamSynchronize(h->stream));
        for (size_t i = 0; i < gps.size(); ++i) {
            const GroupPlan &gp = gps[i];
            const double2 *hl = reinterpret_cast<const double2 *>(pbuf + pin_offs[i] + gp.res_off);
            for (int t = gp.t_begin; t < gp.t_end; ++t) {
                const bool has = gp.F > 0;
                if (loudness_out) loudness_out[t] = has ? hl[t - gp.t_begin].x : NAN;
                if (gain_out && !targets) gain_out[t] = has ? hl[t - gp.t_begin].y : 1.0;
                if (gain_out && targets)
                    for (int k = 0; k < n_targets; ++k)
                        gain_out[(size_t)k * n_tracks + t] = has ? hl[(size_t)(1 + k) * gp.g.n_tracks + (t - gp.t_begin)].y : 1.0;
            }
        }
    }
    return B200M_OK;
}

extern "C" int b200m_master_batch(b200m_handle *h, const void *pcm_in, int in_on_device, int fmt, int n_tracks,
                                  const int64_t *in_offsets, const int64_t *in_frames, const int64_t *out_frames,
                                  const b200m_plan *plans, int n_plans, const int32_t *plan_index,
                                  void *pcm_out, int out_on_device, double *loudness_out, double *gain_out)
{
    return master_batch_impl(h, pcm_in, in_on_device, fmt, n_tracks, in_offsets, in_frames, out_frames, plans, n_plans, plan_index,
                             nullptr, 0, nullptr, pcm_out, out_on_device, loudness_out, gain_out);
}

extern "C" int b200m_master_batch_targets(b200m_handle *h, const void *pcm_in, int in_on_device, int fmt, int n_tracks,
                                          const int64_t *in_offsets, const int64_t *in_frames, const int64_t *out_frames,
                                          const b200m_plan *plans, int n_plans, const int32_t *plan_index,
                                          const double *targets, int n_targets,
                                          void *pcm_out, int out_on_device, double *loudness_out, double *gain_out)
{
    if (!h) return B200M_ERR_INVALID;
    if (!targets || n_targets <= 0 || n_targets > 64) return fail(h, B200M_ERR_INVALID, "b200m_master_batch_targets: 1..64 loudness targets");
    return master_batch_impl(h, pcm_in, in_on_device, fmt, n_tracks, in_offsets, in_frames, out_frames, plans, n_plans, plan_index,
                             targets, n_targets, nullptr, pcm_out, out_on_device, loudness_out, gain_out);
}

extern "C" int b200m_master_batch_wav(b200m_handle *h, const void *pcm_in, int in_on_device, int fmt, int n_tracks,
                                      const int64_t *in_offsets, const int64_t *in_frames, const int64_t *out_frames,
                                      const b200m_plan *plans, int n_plans, const int32_t *plan_index,
                                      const int64_t *out_offsets, void *out, int out_on_device,
                                      double *loudness_out, double *gain_out)
{
    if (!h) return B200M_ERR_INVALID;
    if (!out_offsets) return fail(h, B200M_ERR_INVALID, "b200m_master_batch_wav: out_offsets is required");
    return master_batch_impl(h, pcm_in, in_on_device, fmt, n_tracks, in_offsets, in_frames, out_frames, plans, n_plans, plan_index,
                             nullptr, 0, out_offsets, out, out_on_device, loudness_out, gain_out);
}

extern "C" int b200m_wav_header(int sample_rate, int channels, int64_t frames, unsigned char *out44)
{
    if (!out44 || sample_rate <= 0 || (channels != 1 && channels != 2) || frames < 0 || (uint64_t)frames * channels * 2 > 0xffffffffull - 36)
        return B200M_ERR_INVALID;
    const uint32_t bytes = (uint32_t)(frames * channels * 2);
    const uint32_t w[11] = {0x46464952u, 36u + bytes, 0x45564157u, 0x20746d66u, 16u, 1u | ((uint32_t)channels << 16), (uint32_t)sample_rate,
                            (uint32_t)sample_rate * channels * 2u, ((uint32_t)channels * 2u) | (16u << 16), 0x61746164u, bytes};
    std::memcpy(out44, w, 44);
    return B200M_OK;
}

// ------------------------------------------------------------------------------------
// Time slices of ONE long track (BASELINE config 4: a track split along time over several
// GPUs).  Slices are cut at 30-s chunk boundaries, so everything up to `processed_audio`
// (ENG:48-80) is local to a slice; only the loudness measurement (ENG:82-86, 212-222) couples
// them: the K-weighting filter state at the slice start (provided as a halo of the previous
// slice's processed samples, joined by overlap-discard exactly like the segments inside one
// GPU) and the 400 ms block energies (every block is computed by the slice that holds its first
// frame, from a halo of the next slice; the per-rank z arrays are disjoint, so a SUM all-reduce
// assembles them exactly).  The host (b200master/longtrack.py) moves the halos and the z array
// with NCCL; these entry points do the arithmetic.  All buffers are DEVICE pointers.
// ------------------------------------------------------------------------------------
extern "C" int b200m_stage_pcm(b200m_handle *h, const void *pcm_dev, int fmt, int64_t n_samples, int16_t *out_dev)
{
    if (!h) return B200M_ERR_INVALID;
    if (n_samples < 0 || (n_samples && (!pcm_dev || !out_dev))) return fail(h, B200M_ERR_INVALID, "stage_pcm: bad argument");
    CK(cudaSetDevice(h->device));
    if (n_samples == 0) return B200M_OK;
    const int grid = (int)std::min<int64_t>((n_samples / 4 + 255) / 256 + 1, 148 * 16);
    if (fmt == B200M_FMT_S16) CK(cudaMemcpyAsync(out_dev, pcm_dev, (size_t)n_samples * 2, cudaMemcpyDeviceToDevice, h->stream));
    else if (fmt == B200M_FMT_S24) LAUNCH("k_stage_s24", k_stage_s24<<<grid, 256, 0, h->stream>>>((const unsigned char *)pcm_dev, n_samples, out_dev));
    else if (fmt == B200M_FMT_F32) LAUNCH("k_stage_f32", k_stage_f32<<<grid, 256, 0, h->stream>>>((const float *)pcm_dev, n_samples, out_dev));
    else return fail(h, B200M_ERR_INVALID, "stage_pcm: unknown format %d", fmt);
    CK(cudaGetLastError());
    return B200M_OK;
}

extern "C" int b200m_slice_halo(const b200m_plan *plan, int64_t abs_offset, int64_t *halo_before, int64_t *halo_after)
{
    if (!plan || abs_offset < 0 || plan->sample_rate <= 0) return B200M_ERR_INVALID;
    const double wt = std::ceil(kweight_warm_frames(*plan) / KTILE);
    if (!(wt <= 4096)) return B200M_ERR_INVALID;                 // unstable K-weighting: cannot be joined by overlap-discard
    // the slice's buffer starts at an absolute K-weighting tile boundary, `wt` whole tiles ahead
    // of the tile that holds the slice's first frame (the first slice starts at the track start)
    if (halo_before) *halo_before = abs_offset == 0 ? 0 : std::min<int64_t>(abs_offset, (int64_t)wt * KTILE + abs_offset % KTILE);
    if (halo_after) *halo_after = (int64_t)(0.4 * plan->sample_rate) + 1;     // one 400 ms block beyond the last block start
    return B200M_OK;
}

extern "C" int b200m_slice_chain(b200m_handle *h, const int16_t *pcm_dev, int64_t in_frames, int64_t out_frames,
                                 const b200m_plan *plan, int16_t *proc_dev)
{
    if (!h) return B200M_ERR_INVALID;
    if (!plan || in_frames < 0 || out_frames < 0 || (out_frames && (!pcm_dev || !proc_dev)))
        return fail(h, B200M_ERR_INVALID, "slice_chain: bad argument");
    if (out_frames == 0) return B200M_OK;
    CK(cudaSetDevice(h->device));
    b200m_plan p = *plan;
    p.has_lufs = 0;
    int rc = ensure_plans(h, &p, 1);
    if (rc) return rc;
    const int64_t zero = 0;
    const int32_t pi = 0;
    GroupPlan gp;
    plan_group(h, gp, true, true, 0, 1, &zero, &in_frames, &out_frames, &p, &pi, 0);
    rc = ws_reserve(h, gp.need);
    if (rc) return rc;
    rc = pin_reserve(h, gp.pin_bytes);
    if (rc) return rc;
    CK(cudaStreamSynchronize(h->stream));       // pinned staging of an earlier call may still be in flight
    ExecStreams X;
    X.in = X.comp = X.out = h->stream;
    X.slot_free = X.h2d_done = X.comp_done = X.d2h_done = nullptr;
    return exec_group(h, gp, h->ws, h->pin, X, pcm_dev, true, &zero, &in_frames, proc_dev, true, proc_dev);
}

extern "C" int b200m_slice_energies(b200m_handle *h, const int16_t *proc_ext_dev, int64_t ext_frames, int64_t halo_before,
                                    int64_t local_frames, int64_t abs_offset, int64_t track_frames, const b200m_plan *plan,
                                    double *z_dev, int32_t *first_block_out, int32_t *n_blocks_out)
{
    if (!h) return B200M_ERR_INVALID;
    if (!plan || !proc_ext_dev || !z_dev || ext_frames < 0 || halo_before < 0 || local_frames < 0 || abs_offset < halo_before ||
        halo_before + local_frames > ext_frames || abs_offset + local_frames > track_frames)
        return fail(h, B200M_ERR_INVALID, "slice_energies: bad argument");
    CK(cudaSetDevice(h->device));
    const int rate = plan->sample_rate, ch = plan->channels;
    if ((double)track_frames < 0.4 * rate) return fail(h, B200M_ERR_TOO_SHORT, "audio must have length greater than the block size (400 ms)");
    b200m_plan p = *plan;
    p.has_lufs = 1;
    int rc = ensure_plans(h, &p, 1);
    if (rc) return rc;
    const int64_t abs0 = abs_offset - halo_before;
    if (abs0 % KTILE != 0) return fail(h, B200M_ERR_INVALID, "slice_energies: the buffer must start at a K-weighting tile boundary (use b200m_slice_halo)");
    // blocks of the TRACK whose first frame lies in this slice: l_j = int(0.4 * (j * 0.25) * rate)
    const int nb_total = num_blocks(track_frames, rate);
    auto lj = [&](int64_t j) { return (int64_t)(0.4 * ((double)j * 0.25) * (double)rate); };
    int64_t j0 = (int64_t)std::floor((double)abs_offset / (0.1 * rate)) - 2;
    if (j0 < 0) j0 = 0;
    while (j0 < nb_total && lj(j0) < abs_offset) ++j0;
    int64_t j1 = j0;
    while (j1 < nb_total && lj(j1) < abs_offset + local_frames) ++j1;
    if (first_block_out) *first_block_out = (int32_t)j0;
    if (n_blocks_out) *n_blocks_out = (int32_t)(j1 - j0);
    if (j1 == j0) return B200M_OK;
    // the last block must end inside the buffer (or at the track end)
    {
        int64_t u = (int64_t)(0.4 * ((double)(j1 - 1) * 0.25 + 1.0) * (double)rate);
        if (u > track_frames) u = track_frames;
        if (u - abs0 > ext_frames) return fail(h, B200M_ERR_INVALID, "slice_energies: halo after the slice is too short for its last block");
    }
    // K-weighting segments over the buffer: the tile holding the slice's first frame onwards
    const int64_t begin = (halo_before / KTILE) * KTILE;
    const double wt = std::ceil(kweight_warm_frames(p) / KTILE);
    const int warm = (int)std::min<double>((double)begin, wt * KTILE);
    if (abs_offset != 0 && begin < (int64_t)(wt * KTILE)) return fail(h, B200M_ERR_INVALID, "slice_energies: halo before the slice is too short (use b200m_slice_halo)");
    std::vector<SegDesc> segs;
    {
        const int64_t ntiles = (ext_frames - begin + KTILE - 1) / KTILE;
        const int64_t seg = std::max<int64_t>(std::max<int64_t>(8, 2 * (int64_t)wt), std::min<int64_t>(32, ntiles / (148 * 6)));
        for (int64_t t = 0; t < ntiles; t += seg) {
            const int64_t b = begin + t * KTILE, e = std::min<int64_t>(ext_frames, begin + (t + seg) * KTILE);
            segs.push_back({b, e, 0, (int32_t)(t == 0 ? warm : (int)(wt * KTILE))});
        }
    }
    TrackDesc td = {0, 0, ext_frames, j0, (int32_t)(j1 - j0), 0, abs0, track_frames, (int32_t)j0, 0};
    rc = ws_reserve(h, 65536 + (size_t)ext_frames * 4 + segs.size() * sizeof(SegDesc) + (size_t)(j1 + 4) * 8);
    if (rc) return rc;
    rc = pin_reserve(h, sizeof td + segs.size() * sizeof(SegDesc));
    if (rc) return rc;
    CK(cudaStreamSynchronize(h->stream));
    Arena A(h->ws);
    TrackDesc *d_tracks = A.take<TrackDesc>(1);
    SegDesc *d_segs = A.take<SegDesc>(segs.size());
    float *d_kw = A.take<float>(ext_frames);
    double *d_hops = A.take<double>(j1 + 4);         // indexed like z_dev (td.zoff = j0): hop sums of blocks j0 .. j1 + 2
    std::memcpy(h->pin, &td, sizeof td);
    std::memcpy(h->pin + sizeof td, segs.data(), segs.size() * sizeof(SegDesc));
    CK(cudaMemcpyAsync(d_tracks, h->pin, sizeof td, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(d_segs, h->pin + sizeof td, segs.size() * sizeof(SegDesc), cudaMemcpyHostToDevice, h->stream));
    Group g;
    g.ch = ch; g.n_tracks = 1; g.d_tracks = d_tracks; g.d_ksegs = d_segs; g.n_ksegs = (int)segs.size();
    KwTabsC kt;
    kw_tables(h, kt);                            // one plan: its tables
    if (ch == 2) LAUNCH("k_kweight", k_kweight<2, int16_t, true><<<g.n_ksegs, KNT, kweight_smem_bytes(), h->stream>>>(proc_ext_dev, d_tracks, d_segs, h->d_plans, d_kw, kt));
    else         LAUNCH("k_kweight", k_kweight<1, int16_t, true><<<g.n_ksegs, KNT, kweight_smem_bytes(), h->stream>>>(proc_ext_dev, d_tracks, d_segs, h->d_plans, d_kw, kt));
    const dim3 gb((unsigned)(j1 - j0), 1);
    const bool hops = h->hops_smem_floats > 0 && j1 - j0 >= 3;
    if (hops) LAUNCH("k_hops", k_hops<<<dim3((unsigned)(j1 - j0 + 3), 1), BNT, hops_smem_bytes(h), h->stream>>>(d_kw, d_tracks, h->d_plans, d_hops, h->hops_stage_floats));
    LAUNCH("k_blocks", k_blocks<<<gb, BNT, (size_t)h->blocks_smem_floats * 4 + 16, h->stream>>>(d_kw, d_tracks, h->d_plans, hops ? d_hops : nullptr, z_dev));
    CK(cudaGetLastError());
    return B200M_OK;
}

extern "C" int b200m_track_blocks(int64_t track_frames, int rate)
{
    return (rate > 0 && (double)track_frames >= 0.4 * rate) ? num_blocks(track_frames, rate) : 0;
}

extern "C" int b200m_gate(b200m_handle *h, const double *z_dev, int32_t n_blocks, const b200m_plan *plan, double *loudness_out, double *gain_out)
{
    if (!h) return B200M_ERR_INVALID;
    if (!plan || n_blocks < 0 || (n_blocks && !z_dev)) return fail(h, B200M_ERR_INVALID, "gate: bad argument");
    CK(cudaSetDevice(h->device));
    b200m_plan p = *plan;
    p.has_lufs = 1;
    int rc = ensure_plans(h, &p, 1);
    if (rc) return rc;
    rc = ws_reserve(h, 4096 + (size_t)(n_blocks + 2) * 8);
    if (rc) return rc;
    Arena A(h->ws);
    TrackDesc *d_tracks = A.take<TrackDesc>(1);
    double2 *d_loud = A.take<double2>(1);
    double *d_zsel = A.take<double>(n_blocks + 1);
    TrackDesc td = {0, 0, 0, 0, n_blocks, 0, 0, 0, 0, 0};
    CK(cudaMemcpyAsync(d_tracks, &td, sizeof td, cudaMemcpyHostToDevice, h->stream));
    LAUNCH("k_gate", k_gate<<<1, GNT, 0, h->stream>>>(d_tracks, h->d_plans, z_dev, d_zsel, d_loud));
    CK(cudaGetLastError());
    double2 res;
    CK(cudaMemcpyAsync(&res, d_loud, sizeof res, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (loudness_out) *loudness_out = res.x;
    if (gain_out) *gain_out = res.y;
    return B200M_OK;
}

extern "C" int b200m_slice_final(b200m_handle *h, const int16_t *proc_dev, int64_t frames, const b200m_plan *plan,
                                 int has_gain, double gain, int16_t *out_dev)
{
    if (!h) return B200M_ERR_INVALID;
    if (!plan || frames < 0 || (frames && (!proc_dev || !out_dev))) return fail(h, B200M_ERR_INVALID, "slice_final: bad argument");
    if (frames == 0) return B200M_OK;
    CK(cudaSetDevice(h->device));
    b200m_plan p = *plan;
    p.has_lufs = has_gain != 0;                 // without a loudness target the limiter runs in float32 (ENG:88 on float32 samples)
    int rc = ensure_plans(h, &p, 1);
    if (rc) return rc;
    rc = ws_reserve(h, 4096);
    if (rc) return rc;
    Arena A(h->ws);
    TrackDesc *d_tracks = A.take<TrackDesc>(1);
    double2 *d_loud = A.take<double2>(1);
    TrackDesc td = {0, 0, frames, 0, 0, 0, 0, frames, 0, 0};
    const double2 lg = make_double2(0.0, gain);
    CK(cudaMemcpyAsync(d_tracks, &td, sizeof td, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(d_loud, &lg, sizeof lg, cudaMemcpyHostToDevice, h->stream));
    const dim3 gf((unsigned)std::min<int64_t>((frames + 2047) / 2048, 8192 * 4), 1);
    if (p.channels == 2) LAUNCH("k_final", k_final<2><<<gf, 256, 0, h->stream>>>(proc_dev, d_tracks, h->d_plans, d_loud, out_dev));
    else                 LAUNCH("k_final", k_final<1><<<gf, 256, 0, h->stream>>>(proc_dev, d_tracks, h->d_plans, d_loud, out_dev));
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(h->stream));       // td / lg live on this stack frame
    return B200M_OK;
}

// ------------------------------------------------------------------------------------
// Stage-level entry points (host arrays in, host arrays out)
// ------------------------------------------------------------------------------------
static int grid_for(int64_t n) { return (int)std::min<int64_t>((n + 255) / 256, 148 * 16); }

template <typename F>
static int elementwise(b200m_handle *h, const void *x, size_t in_bytes, void *out, size_t out_bytes, F launch)
{
    if (!h) return B200M_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    if (in_bytes == 0) return B200M_OK;
    int rc = ws_reserve(h, in_bytes + out_bytes + 1024);
    if (rc) return rc;
    Arena A(h->ws);
    char *d_in = A.take<char>(in_bytes);
    char *d_out = A.take<char>(out_bytes);
    CK(cudaMemcpyAsync(d_in, x, in_bytes, cudaMemcpyHostToDevice, h->stream));
    launch(d_in, d_out);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, d_out, out_bytes, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return B200M_OK;
}

extern "C" int b200m_pcm16_to_float(b200m_handle *h, const int16_t *pcm, int64_t n, float *out)
{
    if (!h || n < 0 || (n && (!pcm || !out))) return h ? fail(h, B200M_ERR_INVALID, "pcm16_to_float: bad argument") : B200M_ERR_INVALID;
    return elementwise(h, pcm, n * 2, out, n * 4, [&](char *di, char *dn) {
        LAUNCH("k_pcm16_to_float", k_pcm16_to_float<<<grid_for(n), 256, 0, h->stream>>>((const int16_t *)di, n, (float *)dn));
    });
}

extern "C" int b200m_float_to_pcm16(b200m_handle *h, const void *x, int is_f64, int64_t n, int16_t *out)
{
    if (!h || n < 0 || (n && (!x || !out))) return h ? fail(h, B200M_ERR_INVALID, "float_to_pcm16: bad argument") : B200M_ERR_INVALID;
    return elementwise(h, x, n * (is_f64 ? 8 : 4), out, n * 2, [&](char *di, char *dn) {
        if (is_f64) LAUNCH("k_float_to_pcm16", k_float_to_pcm16<double><<<grid_for(n), 256, 0, h->stream>>>((const double *)di, n, (int16_t *)dn));
        else        LAUNCH("k_float_to_pcm16", k_float_to_pcm16<float><<<grid_for(n), 256, 0, h->stream>>>((const float *)di, n, (int16_t *)dn));
    });
}

extern "C" int b200m_saturation(b200m_handle *h, const float *x, int64_t n, double pct, float *out)
{
    if (!h || n < 0 || (n && (!x || !out))) return h ? fail(h, B200M_ERR_INVALID, "saturation: bad argument") : B200M_ERR_INVALID;
    const double mix = (pct / 100.0) * (pct / 100.0);
    if (pct == 0) { if (out != x) std::memcpy(out, x, n * 4); return B200M_OK; }      // ENG:129 bypass
    return elementwise(h, x, n * 4, out, n * 4, [&](char *di, char *dn) {
        LAUNCH("k_saturation", k_saturation<<<grid_for(n), 256, 0, h->stream>>>((const float *)di, n, (float)(1 - mix), (float)mix,
                                                                                  (float)(1 + mix * 4), (float *)dn));
    });
}

extern "C" int b200m_saturation_pcm(b200m_handle *h, const int16_t *pcm, int64_t n, const float *sat_lut, uint64_t key, float *out)
{
    if (!h || n < 0 || !sat_lut || (n && (!pcm || !out))) return h ? fail(h, B200M_ERR_INVALID, "saturation_pcm: bad argument") : B200M_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));       // get_sat_lut may evict a table: nothing in flight may read it
    b200m_plan p;
    std::memset(&p, 0, sizeof p);
    p.sat_lut = sat_lut; p.sat_lut_key = key;
    ++h->tab_tick;
    const float *d_lut = nullptr;
    int rc = get_sat_lut(h, p, &d_lut);
    if (rc) return rc;
    return elementwise(h, pcm, n * 2, out, n * 4, [&](char *di, char *dn) {
        LAUNCH("k_saturation_pcm", k_saturation_pcm<<<grid_for(n), 256, 0, h->stream>>>((const int16_t *)di, n, d_lut, (float *)dn));
    });
}

extern "C" int b200m_stereo_width(b200m_handle *h, const void *x, int is_f64, int64_t nframes, double width, void *out)
{
    if (!h || nframes < 0 || (nframes && (!x || !out))) return h ? fail(h, B200M_ERR_INVALID, "stereo_width: bad argument") : B200M_ERR_INVALID;
    const size_t bytes = (size_t)nframes * 2 * (is_f64 ? 8 : 4);
    return elementwise(h, x, bytes, out, bytes, [&](char *di, char *dn) {
        if (is_f64) LAUNCH("k_width", k_width<double><<<grid_for(nframes), 256, 0, h->stream>>>((const double *)di, nframes, width, (double *)dn));
        else        LAUNCH("k_width", k_width<float><<<grid_for(nframes), 256, 0, h->stream>>>((const float *)di, nframes, width, (float *)dn));
    });
}

extern "C" int b200m_soft_limiter(b200m_handle *h, const void *x, int is_f64, int64_t n, double thr, void *out)
{
    if (!h || n < 0 || (n && (!x || !out))) return h ? fail(h, B200M_ERR_INVALID, "soft_limiter: bad argument") : B200M_ERR_INVALID;
    const size_t bytes = (size_t)n * (is_f64 ? 8 : 4);
    return elementwise(h, x, bytes, out, bytes, [&](char *di, char *dn) {
        if (is_f64) LAUNCH("k_limiter", k_limiter<double><<<grid_for(n), 256, 0, h->stream>>>((const double *)di, n, thr, (double *)dn));
        else        LAUNCH("k_limiter", k_limiter<float><<<grid_for(n), 256, 0, h->stream>>>((const float *)di, n, thr, (float *)dn));
    });
}

extern "C" int b200m_sosfilt(b200m_handle *h, const b200m_biquad *sections, int nsec, const void *x, int is_f64,
                             int64_t nframes, int channels, double *out)
{
    if (!h) return B200M_ERR_INVALID;
    if (nsec < 0 || nsec > 8 || nframes < 0 || channels < 1 || (nsec && !sections) || (nframes && (!x || !out)))
        return fail(h, B200M_ERR_INVALID, "sosfilt: bad argument (1..8 sections supported)");
    CK(cudaSetDevice(h->device));
    const size_t n = (size_t)nframes * channels, esz = is_f64 ? 8 : 4;
    if (n == 0) return B200M_OK;
    std::vector<SecTab> tabs(std::max(nsec, 1));
    for (int s = 0; s < nsec; ++s) build_sectab(sections[s], tabs[s]);
    int rc = ws_reserve(h, n * esz + n * 8 + tabs.size() * sizeof(SecTab) + 1024);
    if (rc) return rc;
    Arena A(h->ws);
    SecTab *d_tabs = A.take<SecTab>(tabs.size());
    char *d_in = A.take<char>(n * esz);
    double *d_out = A.take<double>(n);
    CK(cudaMemcpyAsync(d_tabs, tabs.data(), tabs.size() * sizeof(SecTab), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(d_in, x, n * esz, cudaMemcpyHostToDevice, h->stream));
    if (is_f64) LAUNCH("k_sosfilt", k_sosfilt<double><<<channels, KNT, sosfilt_smem_bytes(), h->stream>>>((const double *)d_in, nframes, channels, d_tabs, nsec, d_out));
    else        LAUNCH("k_sosfilt", k_sosfilt<float><<<channels, KNT, sosfilt_smem_bytes(), h->stream>>>((const float *)d_in, nframes, channels, d_tabs, nsec, d_out));
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, d_out, n * 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return B200M_OK;
}

// One int16 chunk through crossover + 3 compressors + overlay (zero state): the batch
// path with a plan stripped down to the multiband stage.
extern "C" int b200m_multiband(b200m_handle *h, const b200m_plan *plan, const int16_t *pcm, int64_t nframes, int16_t *out)
{
    if (!h) return B200M_ERR_INVALID;
    if (!plan || nframes < 0 || nframes > 0x7fffffff || (nframes && (!pcm || !out)))
        return fail(h, B200M_ERR_INVALID, "multiband: bad argument");
    if (nframes == 0) return B200M_OK;
    CK(cudaSetDevice(h->device));
    b200m_plan p = *plan;
    p.sat_on = 0; p.n_eq = 0; p.width_on = 0; p.has_lufs = 0; p.multiband = 1;
    int rc = ensure_plans(h, &p, 1);
    if (rc) return rc;
    const int ch = p.channels;
    const size_t F = (size_t)nframes;
    Group g;
    g.ch = ch; g.n_streams = 1; g.n_tracks = 1; g.max_stream_frames = (int)nframes;
    for (int b = 0; b < 3; ++b) g.max_look = std::max(g.max_look, p.band[b].look_frames);
    g.total_blocks = (nframes + 1023) / 1024;
    rc = ws_reserve(h, 16384 + F * ch * 2 * 5 + compressor_ws_bytes(h, g, (int64_t)F, 3));
    if (rc) return rc;
    Arena A(h->ws);
    StreamDesc *d_streams = A.take<StreamDesc>(1);
    SegDesc *d_csegs = A.take<SegDesc>(1);
    int16_t *d_in = A.take<int16_t>(F * ch);
    int16_t *d_proc = A.take<int16_t>(F * ch);
    BandPtrs bp;
    std::memset(&bp, 0, sizeof bp);
    for (int b = 0; b < 3; ++b) bp.band[b] = A.take<int16_t>(F * ch);
    double *d_spec = take_compressor_ws(h, A, g, (int64_t)F, 3, 0, bp);
    StreamDesc sd = {0, 0, (int32_t)nframes, (int32_t)nframes, 0, 0, 0, 0};
    CK(cudaMemcpyAsync(d_streams, &sd, sizeof sd, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(d_in, pcm, F * ch * 2, cudaMemcpyHostToDevice, h->stream));
    SegDesc sg = {0, (int64_t)nframes, 0, 0};
    CK(cudaMemcpyAsync(d_csegs, &sg, sizeof sg, cudaMemcpyHostToDevice, h->stream));
    g.d_streams = d_streams;
    if (ch == 2) LAUNCH("k_chain", k_chain<2, true><<<1, NSEG * 2, chain_smem_bytes<2>(), h->stream>>>(d_in, d_streams, d_csegs, h->d_plans, d_proc, bp));
    else         LAUNCH("k_chain", k_chain<1, true><<<1, NSEG * 1, chain_smem_bytes<1>(), h->stream>>>(d_in, d_streams, d_csegs, h->d_plans, d_proc, bp));
    CK(cudaGetLastError());
    rc = launch_compressor(h, g, bp, 3, 0, d_proc, d_spec);
    if (rc) return rc;
    CK(cudaMemcpyAsync(out, d_proc, F * ch * 2, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return B200M_OK;
}

extern "C" int b200m_compress_dynamic_range(b200m_handle *h, const int16_t *pcm, int64_t nframes, int channels,
                                            const b200m_band *band, int16_t *out, double *att_out, uint32_t *rms_out)
{
    if (!h) return B200M_ERR_INVALID;
    if (!band || nframes < 0 || nframes > 0x7fffffff || (channels != 1 && channels != 2) || (nframes && (!pcm || !out)))
        return fail(h, B200M_ERR_INVALID, "compress_dynamic_range: bad argument");
    if (nframes == 0) return B200M_OK;
    CK(cudaSetDevice(h->device));
    b200m_plan p;
    std::memset(&p, 0, sizeof p);
    p.sample_rate = 1; p.channels = channels; p.multiband = 1;
    p.lp[0] = p.lp[1] = p.hp[0] = p.hp[1] = p.kw[0] = p.kw[1] = {1, 0, 0, 0, 0};
    p.band[0] = p.band[1] = p.band[2] = *band;
    int rc = ensure_plans(h, &p, 1);
    if (rc) return rc;
    const size_t F = (size_t)nframes;
    Group g;
    g.ch = channels; g.n_streams = 1; g.n_tracks = 1; g.max_stream_frames = (int)nframes;
    g.max_look = band->look_frames;
    g.total_blocks = (nframes + 1023) / 1024;
    rc = ws_reserve(h, 16384 + F * channels * 2 * 2 + compressor_ws_bytes(h, g, (int64_t)F, 1, att_out != nullptr));
    if (rc) return rc;
    Arena A(h->ws);
    StreamDesc *d_streams = A.take<StreamDesc>(1);
    int16_t *d_in = A.take<int16_t>(F * channels);
    int16_t *d_proc = A.take<int16_t>(F * channels);
    BandPtrs bp;
    std::memset(&bp, 0, sizeof bp);
    bp.band[0] = d_in;
    double *d_spec = take_compressor_ws(h, A, g, (int64_t)F, 1, 0, bp, att_out != nullptr);
    StreamDesc sd = {0, 0, (int32_t)nframes, (int32_t)nframes, 0, 0, 0, 0};
    CK(cudaMemcpyAsync(d_streams, &sd, sizeof sd, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(d_in, pcm, F * channels * 2, cudaMemcpyHostToDevice, h->stream));
    g.d_streams = d_streams;
    rc = launch_compressor(h, g, bp, 1, 0, d_proc, d_spec);
    if (rc) return rc;
    CK(cudaMemcpyAsync(out, d_proc, F * channels * 2, cudaMemcpyDeviceToHost, h->stream));
    if (att_out) CK(cudaMemcpyAsync(att_out, bp.att[0], F * 8, cudaMemcpyDeviceToHost, h->stream));
    std::vector<uint16_t> r16;
    if (rms_out) { r16.resize(F); CK(cudaMemcpyAsync(r16.data(), bp.rms[0], F * 2, cudaMemcpyDeviceToHost, h->stream)); }
    CK(cudaStreamSynchronize(h->stream));
    if (rms_out) for (size_t i = 0; i < F; ++i) rms_out[i] = r16[i];
    return B200M_OK;
}

static int loudness_core(b200m_handle *h, const b200m_biquad *kw, const float *x, int64_t n, int channels, int rate,
                         double target, double *scaled_out, double *lufs_out, double *gain_out)
{
    if ((double)n < 0.4 * rate) return fail(h, B200M_ERR_TOO_SHORT, "audio must have length greater than the block size (400 ms)");
    CK(cudaSetDevice(h->device));
    b200m_plan p;
    std::memset(&p, 0, sizeof p);
    p.sample_rate = rate; p.channels = 1; p.has_lufs = 1; p.lufs = target;
    p.kw[0] = kw[0]; p.kw[1] = kw[1];
    int rc = ensure_plans(h, &p, 1);
    if (rc) return rc;
    TrackDesc td = {0, 0, n, 0, num_blocks(n, rate), 0, 0, n, 0, 0};
    const size_t ns = (size_t)n * channels;
    rc = ws_reserve(h, 65536 + ns * 4 + (size_t)n * 8 + (scaled_out ? ns * 8 : 0) + (size_t)(td.nblocks + 2) * 16 + (size_t)(n / KTILE + 2) * sizeof(SegDesc));
    if (rc) return rc;
    Arena A(h->ws);
    TrackDesc *d_tracks = A.take<TrackDesc>(1);
    std::vector<SegDesc> ksegs;
    make_segments(ksegs, 0, n, KTILE, kweight_warm_frames(p), h->seg_kweight == 0 ? auto_seg_tiles((n + KTILE - 1) / KTILE, 8, 32) : h->seg_kweight);
    SegDesc *d_ksegs = A.take<SegDesc>(ksegs.size());
    double2 *d_loud = A.take<double2>(1);
    float *d_x = A.take<float>(ns);
    float *d_mono = channels == 2 ? A.take<float>(n) : d_x;
    float *d_kw = A.take<float>(n);
    double *d_z = A.take<double>(td.nblocks + 1);
    double *d_zsel = A.take<double>(td.nblocks + 1);
    double *d_scaled = scaled_out ? A.take<double>(ns) : nullptr;
    CK(cudaMemcpyAsync(d_tracks, &td, sizeof td, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(d_ksegs, ksegs.data(), ksegs.size() * sizeof(SegDesc), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(d_x, x, ns * 4, cudaMemcpyHostToDevice, h->stream));
    if (channels == 2) LAUNCH("k_mono_mean", k_mono_mean<<<grid_for(n), 256, 0, h->stream>>>(d_x, n, d_mono));
    Group g;
    g.ch = 1; g.n_tracks = 1; g.d_tracks = d_tracks; g.max_blocks = td.nblocks; g.any_lufs = true; g.max_track_frames = n;
    g.d_ksegs = d_ksegs; g.n_ksegs = (int)ksegs.size();
    rc = launch_loudness(h, g, nullptr, d_mono, d_kw, d_z, d_zsel, d_loud);
    if (rc) return rc;
    if (scaled_out) {
        LAUNCH("k_scale", k_scale<<<grid_for((int64_t)ns), 256, 0, h->stream>>>(d_x, (int64_t)ns, d_loud, d_scaled));
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(scaled_out, d_scaled, ns * 8, cudaMemcpyDeviceToHost, h->stream));
    }
    double2 res;
    CK(cudaMemcpyAsync(&res, d_loud, sizeof res, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (lufs_out) *lufs_out = res.x;
    if (gain_out) *gain_out = res.y;
    return B200M_OK;
}

extern "C" int b200m_integrated_loudness(b200m_handle *h, const b200m_biquad *kw, const float *mono, int64_t n, int rate, double *lufs_out)
{
    if (!h) return B200M_ERR_INVALID;
    if (!kw || !mono || !lufs_out || n < 0 || rate <= 0) return fail(h, B200M_ERR_INVALID, "integrated_loudness: bad argument");
    return loudness_core(h, kw, mono, n, 1, rate, 0.0, nullptr, lufs_out, nullptr);
}

extern "C" int b200m_normalize_to_lufs(b200m_handle *h, const b200m_biquad *kw, const float *x, int64_t n_frames, int channels,
                                       int rate, double target_lufs, double *out, double *loudness_out, double *gain_out)
{
    if (!h) return B200M_ERR_INVALID;
    if (!kw || !x || !out || n_frames < 0 || rate <= 0 || (channels != 1 && channels != 2))
        return fail(h, B200M_ERR_INVALID, "normalize_to_lufs: bad argument");
    return loudness_core(h, kw, x, n_frames, channels, rate, target_lufs, out, loudness_out, gain_out);
}
