// FA loss, POSITION semantics (the paper's N x N position affinity; opt-in), sm_100a.
//
//   F  = avgpool_k(X) viewed as (C, N), N = h*w                    (pool step as FALoss.py:23-24)
//   Fh = F / max(||F_i||_2 over channels, 1e-12)                   per position i
//   S  = Fh^T Fh  (N x N)                                          (the contraction of FALoss.py:11 on a (C, N) view)
//   L  = reduce_{i,j} | S1 - S2 |_{ij},  diagonal forced to 0      (mean over B*N*N, or sum)
//
// The N x N affinity never exists in memory.  With Fcat = [Fh1 ; Fh2] (C1+C2 channels per position) the difference
// is ONE contraction,  D = S1 - S2 = Fcat^T diag(+1..+1,-1..-1) Fcat,  so a 128 x 128 tile of D is accumulated
// directly in tensor memory by tcgen05.mma (kind::tf32, the branch-2 channel chunks issued with the negate-A bit),
// and the gradient  dL/dFcat_i = (2/Z) sum_j sign(D_ij) Fcat_j  is a second contraction whose A operand -- the sign
// tile -- is written back into the SAME tensor-memory columns by the epilogue warps and consumed from there
// (A-from-TMEM), like P = softmax(S) in an attention kernel.
//
// Kernels:
//   fa_pos_pack     pool + per-position L2 normalise, TF32 rounding (or FP16 copies only, precision 'f16'), two layouts:
//                     Fpm (B, Npad, Kc) position-major  -> K-major operand tiles of the D contraction
//                     Fcm (B, Kc, Npad) channel-major   -> K-major B tiles of the gradient contraction
//   fa_pos_tiles    one CTA per (128-row tile i, channel group, sample): warp 0 = TMA producer, warp 1 = MMA issuer,
//                   warps 2..9 = epilogue (|D| sum, sign tile, final normalisation Jacobian).  TMEM: columns
//                   [0,256) gradient accumulator, [256,512) two D / sign tiles (double buffered).
//   fa_pos_tiles_pair  the same for a cluster of two CTAs (two row tiles) with tcgen05 cta_group::2, M = 256: each CTA
//                   supplies half of every B tile, which halves the B-operand shared-memory traffic (the hot variant)
//                   With kHalf the operands are FP16 (kind::f16, same 11-bit significand as TF32 on unit-norm features):
//                   twice the tensor rate, half the operand bytes, own rows always resident (precision 'f16')
//   fa_pos_jacobian sums the partial accumulators when the column range of a row tile is split over several CTAs
//   fa_pos_unpool   backward proper: dX = grad_out / k^2 * unpool(dP)   (not launched by the fused forward + backward call
//                   without pooling: there the gradient kernel / fa_pos_jacobian write dX themselves)
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include <mutex>

#include "common.cuh"
#include "tc05.cuh"

namespace dsrl {
namespace {

using namespace tc;

constexpr int kTile = 128;          // positions per tile (MMA M and N)
constexpr int kChunk = 32;          // fp32 elements per 128-byte swizzle row = K extent of one operand box
constexpr int kBoxBytes = kTile * kChunk * 4;    // 16 KB: one TMA box (128 rows x 32 fp32, 128-byte swizzle)
// a pipeline stage holds up to kSB boxes behind one full/empty barrier pair: 2 when the CTA's own operand rows are
// resident in shared memory (little room left), 4 when everything is streamed (so one barrier round trip feeds >= 8 MMAs)
__host__ __device__ constexpr int stage_boxes(bool resident) { return resident ? 2 : 4; }
constexpr int kEpiWarps = 8;        // two warps per TMEM lane quarter, each converting half of a tile's columns
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kThreads = 64 + kEpiThreads;   // warp 0 = TMA producer, warp 1 = MMA issuer, warps 2.. = epilogue
constexpr int kMaxGroupCh = 256;    // gradient accumulator columns per CTA
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kColD = 256;     // first column of the two D / sign tiles
constexpr size_t kSmemBudget = 227 * 1024;
constexpr size_t kSmemAux = 1536;   // barriers, TMEM pointer, reduction scratch, projection halves
// every tile kernel allocates all 512 tensor-memory columns, so two of its CTAs must never share an SM (the second would
// spin in tcgen05.alloc; with CTA pairs that can deadlock): each requests more than half of an SM's shared memory
constexpr size_t kOneCtaSmem = 116 * 1024;
constexpr size_t kSignBlockBytes = 2 * 128 * 16;              // two-pass form: sign planes, bytes per (row tile, column tile)
constexpr size_t kSignPlaneCapBytes = (size_t)6 << 30;        // largest sign-plane footprint the two-pass form may use

struct PosGeom {
    int B, C1, C2, H, W, k, h, w, N, Npad, C1p, C2p, Kc, G;
    int gbeg[2], gcnt[2];   // channel range of group g on the concatenated channel axis
    int tiles, nkc;         // Npad / 128, Kc / 32
    int split;              // 1: 3xTF32 -- operands carried as hi + lo TF32 parts, D = hi*hi + hi*lo + lo*hi
    int q_resident, stages;
    int jsplit;             // gradient variant: the column tiles of one row tile are spread over jsplit CTAs (small grids)
    int pair, pair_stages;  // CTA-pair form of the gradient variant usable for this geometry; its ring depth
    int half, nkh, half_stages;   // FP16 operands (kind::f16) requested; Kc / 64 chunks; ring depth of the pair form
    int half_pair, half1_stages;  // pair form usable for this geometry; ring depth of the single-CTA form
    int exact, fnsub, fsub, fcap; // exact signs: near-tie entries of D are listed per row -- 2 * jsplit private sub-lists (one per
                                  // epilogue thread that converts part of the row) of fsub entries, fcap in all -- and re-decided in FP64
    int raw_o;                    // the tile kernel stores raw accumulator rows, fa_pos_finish completes them (jsplit > 1 or exact)
    int ab;                       // FP16 form as two symmetric passes (fa_position_ab.cuh): D tiles j >= i -> sign planes -> gradient
    int a_chunk, a_units;         // pass A: column tiles per work unit (tiles: no split), work units per sample
    int a_stages;                 // pass A: depth of the K ring (stages of four 64-channel chunks of this CTA's 64 rows = 32 KB)
    size_t a_smem_bytes;
    int b_stages;                 // pass B: depth of the V ring
    size_t b_smem_bytes, sb_bytes;   // pass B shared memory; sign planes of all samples
    size_t smem_bytes, pair_smem_bytes, half_smem_bytes, half1_smem_bytes;
};

struct PosWs { size_t Fpm, Fcm, FpmH, FcmH, nrm, partials, opart, Ppm, inv64, tau, fcnt, fent, sb, total; };
struct PosSaved { size_t dP, total; };

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

inline bool make_geom(int B, int C1, int C2, int H, int W, int k, int split, PosGeom &g, int half = 0, int exact = 0) {
    if (B < 1 || C1 < 1 || C2 < 1 || H < 1 || W < 1 || k < 1) return false;
    g.B = B; g.C1 = C1; g.C2 = C2; g.H = H; g.W = W; g.k = k;
    g.h = H / k; g.w = W / k;
    if (g.h < 1 || g.w < 1) return false;
    const long long N = (long long)g.h * g.w;
    if (N > (1 << 20)) return false;
    g.N = (int)N;
    g.Npad = (int)align_up((size_t)N, kTile);
    g.C1p = (int)align_up((size_t)C1, kChunk);
    g.C2p = (int)align_up((size_t)C2, kChunk);
    if (g.C1p > kMaxGroupCh || g.C2p > kMaxGroupCh) return false;
    g.Kc = g.C1p + g.C2p;
    if (g.Kc <= kMaxGroupCh) { g.G = 1; g.gbeg[0] = 0; g.gcnt[0] = g.Kc; g.gbeg[1] = 0; g.gcnt[1] = 0; }
    else { g.G = 2; g.gbeg[0] = 0; g.gcnt[0] = g.C1p; g.gbeg[1] = g.C1p; g.gcnt[1] = g.C2p; }
    g.tiles = g.Npad / kTile;
    g.nkc = g.Kc / kChunk;
    if ((long long)B * g.Npad > 0x7fffffffLL / 2 || (long long)B * g.Kc + kTile > 0x7fffffffLL / 2) return false;
    const int avail = (int)((kSmemBudget - 1024 - kSmemAux) / kBoxBytes);    // 1024: alignment slack of the dynamic base
    g.split = split ? 1 : 0;
    const int qboxes = g.nkc * (1 + g.split);
    g.q_resident = qboxes <= 8;
    const int sb = stage_boxes(g.q_resident);
    g.stages = (avail - (g.q_resident ? qboxes : 0)) / sb;
    if (g.stages > 6) g.stages = 6;
    g.smem_bytes = 1024 + (size_t)(g.q_resident ? qboxes : 0) * kBoxBytes + (size_t)g.stages * sb * kBoxBytes + kSmemAux;
    if (g.smem_bytes < kOneCtaSmem) g.smem_bytes = kOneCtaSmem;
    // Wave quantisation: every CTA of the gradient variant does the same work (all column tiles of one row tile), and
    // only one CTA fits per SM, so a grid of e.g. 512 CTAs takes ceil(512/148) = 4 rounds instead of 3.46.  Splitting
    // the column range over `jsplit` CTAs (partial accumulators summed by fa_pos_jacobian) makes the rounds shorter.
    g.jsplit = 1;
    const long long ctas = (long long)g.tiles * g.G * B, sms = device_sm_count();
    double best = (double)((ctas + sms - 1) / sms);
    for (int s = 2; s <= 4; s *= 2) {
        if (g.tiles % s) break;
        const double cost = (double)((ctas * s + sms - 1) / sms) / s;
        if (cost < 0.96 * best) { best = cost; g.jsplit = s; }
    }
    // CTA-pair form: needs an even number of row tiles
    g.pair = g.tiles % 2 == 0;
    if (const char *e = getenv("DSRL_POS_PAIR")) { if (atoi(e) == 0) g.pair = 0; }
    {
        const int stage = g.q_resident ? 4 * (kBoxBytes / 2) : 2 * (kBoxBytes + kBoxBytes / 2);
        const size_t qbytes = g.q_resident ? (size_t)qboxes * kBoxBytes : 0;
        g.pair_stages = (int)((kSmemBudget - 1024 - kSmemAux - qbytes) / stage);
        if (g.pair_stages > 6) g.pair_stages = 6;
        if (const char *e = getenv("DSRL_POS_STAGES")) { const int v = atoi(e); if (v >= 2 && v < g.pair_stages) g.pair_stages = v; }   // tuning hook
        g.pair_smem_bytes = 1024 + qbytes + (size_t)g.pair_stages * stage + kSmemAux;
        if (g.pair_smem_bytes < kOneCtaSmem) g.pair_smem_bytes = kOneCtaSmem;
    }
    // FP16 operands: 64 channels per 128-byte row, so the CTA's own rows (<= 8 boxes) are always resident
    g.nkh = (g.Kc + 63) / 64;
    g.half = half && !g.split;
    g.half_pair = g.half && g.pair && (g.G == 1 || g.gcnt[0] == g.gcnt[1]);
    {
        g.half1_stages = (avail - g.nkh) / 2;
        if (g.half1_stages > 6) g.half1_stages = 6;
        g.half1_smem_bytes = 1024 + (size_t)g.nkh * kBoxBytes + (size_t)g.half1_stages * 2 * kBoxBytes + kSmemAux;
        if (g.half1_smem_bytes < kOneCtaSmem) g.half1_smem_bytes = kOneCtaSmem;
    }
    {
        const size_t qbytes = (size_t)g.nkh * kBoxBytes;
        g.half_stages = (int)((kSmemBudget - 1024 - kSmemAux - qbytes) / (2 * kBoxBytes));
        if (g.half_stages > 6) g.half_stages = 6;
        if (const char *e = getenv("DSRL_POS_STAGES")) { const int v = atoi(e); if (v >= 2 && v < g.half_stages) g.half_stages = v; }
        g.half_smem_bytes = 1024 + qbytes + (size_t)g.half_stages * 2 * kBoxBytes + kSmemAux;
        if (g.half_smem_bytes < kOneCtaSmem) g.half_smem_bytes = kOneCtaSmem;
        if (g.half_stages < 2) g.half_pair = 0;
    }
    if (const char *force = getenv("DSRL_POS_JSPLIT")) {          // test hook: force 1, 2 or 4 (when it divides the tile count)
        const int s = atoi(force);
        if ((s == 1 || s == 2 || s == 4) && g.tiles % s == 0) g.jsplit = s;
    }
    // FP16 form as two symmetric passes over CTA pairs (fa_position_ab.cuh): whenever the pair form applies and the sign
    // planes (N^2 / 8 bytes per sample) stay below the cap
    {
        g.sb_bytes = (size_t)B * ((size_t)g.tiles * (g.tiles + 1) / 2) * kSignBlockBytes;      // packed upper triangle of blocks
        // (with tensor-core signs and few channels both passes are bound by their conversion warps and the fused pair kernel,
        // which converts every tile once, is ahead: 3.0 vs 3.5 ms at C = 32 per branch, batch 8)
        g.ab = g.half_pair && g.sb_bytes <= kSignPlaneCapBytes && (exact || g.Kc > 128);
        if (const char *e = getenv("DSRL_POS_AB")) { if (atoi(e) == 0) g.ab = 0; }
        if (!g.ab) g.sb_bytes = 0;
        // pass A: a pair of row tiles against its column tiles j >= 2p is one work unit when there are plenty of them; small
        // grids cut the column range into chunks (no accumulator: a chunk only re-loads the pair's own rows) so that the
        // triangle balances over the SMs.  Near-tie sub-lists are per (chunk, column half): at most 8 chunks.
        // (a unit pays ~4 tiles of time for loading its own rows, so whole rows win as soon as every SM pair has one: list
        // scheduling of the 128 rows of one N = 32768 sample on 74 pairs is 86 % efficient unsplit, 81 % in chunks of 32 tiles)
        g.a_chunk = g.tiles;
        if ((long long)B * (g.tiles / 2) < sms / 2) g.a_chunk = (g.tiles + 7) / 8 < 8 ? (g.tiles < 8 ? g.tiles : 8) : (g.tiles + 7) / 8;
        if (const char *e = getenv("DSRL_POS_ACHUNK")) {          // test hook
            const int v = atoi(e);
            if (v >= 1 && (g.tiles + v - 1) / v <= 8) g.a_chunk = v < g.tiles ? v : g.tiles;
        }
        g.a_units = 0;
        for (int p = 0; p < g.tiles / 2; ++p) g.a_units += (g.tiles - 2 * p + g.a_chunk - 1) / g.a_chunk;
        {
            const long long room = (long long)kSmemBudget - 1024 - (long long)kSmemAux - (long long)g.nkh * kBoxBytes;
            g.a_stages = room > 0 ? (int)(room / (2 * kBoxBytes)) : 0;
            if (g.a_stages > 6) g.a_stages = 6;
            g.a_smem_bytes = 1024 + (size_t)g.nkh * kBoxBytes + (size_t)g.a_stages * 2 * kBoxBytes + kSmemAux;
            if (g.a_smem_bytes < kOneCtaSmem) g.a_smem_bytes = kOneCtaSmem;
            if (g.a_stages < 2) { g.ab = 0; g.sb_bytes = 0; }
        }
        g.b_stages = 8;                                           // 8 x 16 KB V boxes (4 tiles of MMA time ahead) + 4 boxes of own feature rows
        g.b_smem_bytes = 1024 + (size_t)(g.b_stages + 4) * kBoxBytes + 6144;      // barriers, reduction scratch, projection parts [4][2][128]
    }
    // exact signs: about 1e-3 of a row's entries are near ties (|D| below ~3.5 sigma of the operand-rounding error, whatever
    // C is: threshold and spread of D both scale like 1/sqrt(C)); the per-row list holds 4x that, at least 32 entries
    g.exact = exact ? 1 : 0;
    {
        // sub-lists per row: column half x column share (x 2 when the two channel groups split the column tiles between them);
        // expected N * 1.1e-3 / nsub entries each, room for 4x that + 16.  Two-pass form: column half x column chunk of pass A
        // (upper triangle only, so a row holds at most what a full row would): one sub-list per 32-column strip of a tile and chunk.
        const int nsub = g.fnsub = g.ab ? 4 * ((g.tiles + g.a_chunk - 1) / g.a_chunk) : 2 * g.jsplit;
        g.fsub = (int)align_up((size_t)(N / (200 * nsub)) + 16, 8);
        if (g.fsub > 2048) g.fsub = 2048;
        g.fcap = g.fsub * nsub;
    }
    g.raw_o = g.jsplit > 1 || (g.exact && !g.ab);      // two-pass form: the resolve pass fixes the sign planes, not the accumulators
    return true;
}

// Only the operand copies the chosen precision reads are reserved: fp32 (TF32-rounded) layouts for the tf32 kinds, FP16
// layouts for kind::f16 (at BASELINE configs[3] that is 0.54 GB instead of the 2.2 GB a precision-blind maximum needs).
inline PosWs make_ws(const PosGeom &g) {
    PosWs w;
    size_t off = 0;
    w.Fpm = off;      off = align_up(off + (g.half ? 0 : (size_t)(1 + g.split) * g.B * g.Npad * g.Kc * 4), 1024);   // hi rows, then lo rows
    w.Fcm = off;      off = align_up(off + (g.half ? 0 : ((size_t)g.B * g.Kc + kTile) * g.Npad * 4), 1024);   // + one box of slack rows
    w.FpmH = off;     off = align_up(off + (g.half ? (size_t)g.B * g.Npad * g.Kc * 2 : 0), 1024);           // FP16 copies of both layouts
    w.FcmH = off;     off = align_up(off + (g.half ? ((size_t)g.B * g.Kc + kTile) * g.Npad * 2 : 0), 1024);
    w.nrm = off;      off = align_up(off + (size_t)g.B * 2 * g.Npad * 4, 256);
    {
        size_t np = (size_t)g.B * g.tiles * g.jsplit * g.G;
        if (g.ab && (size_t)g.B * 2 * g.a_units > np) np = (size_t)g.B * 2 * g.a_units;       // pass A: one loss partial per CTA
        w.partials = off; off = align_up(off + np * 8 + 16384, 256);   // + debug timing area
    }
    w.opart = off;    off = align_up(off + (g.raw_o ? (size_t)g.jsplit * g.B * g.Npad * g.Kc * 4 : 0), 256);
    // exact signs: raw pooled features (position-major fp32), FP64 inverse norms, per-sample tie threshold, per-row tie lists
    w.Ppm = off;      off = align_up(off + (g.exact ? (size_t)g.B * g.Npad * g.Kc * 4 : 0), 256);
    w.inv64 = off;    off = align_up(off + (g.exact ? (size_t)g.B * 2 * g.Npad * 8 : 0), 256);
    w.tau = off;      off = align_up(off + (g.exact ? (size_t)g.B * 4 : 0), 256);
    w.fcnt = off;     off = align_up(off + (g.exact ? (size_t)g.B * g.Npad * g.fnsub * 4 : 0), 256);
    w.fent = off;     off = align_up(off + (g.exact ? (size_t)g.B * g.Npad * g.fcap * 4 : 0), 256);
    w.sb = off;       off = align_up(off + g.sb_bytes, 256);
    w.total = off;
    return w;
}

inline PosSaved make_saved(const PosGeom &g) {
    PosSaved s;
    s.dP = 256;   // [0,8) double: local sum of |D|; exact signs: [8,16) ties listed, [16,24) signs corrected, [24,32) ties dropped
                  // (list full), [32,36) float: largest |D_exact| / threshold among the corrected entries
    s.total = s.dP + (size_t)g.B * g.Kc * g.Npad * 4;
    return s;
}

// ---------------------------------------------------------------------------------------------------------------
// pack: pool, normalise over channels, round to TF32, write both operand layouts
// ---------------------------------------------------------------------------------------------------------------
// kExact (exact signs): the norm is summed in FP64, and the kernel also leaves what fa_pos_finish needs to re-decide near
// ties: the raw pooled features position-major (Ppm), FP64 inverse norms and zeroed statistics.
struct PackExact { float *Ppm; double *inv64; unsigned long long *stats; };

template <bool kExact>
__global__ void __launch_bounds__(256) fa_pos_pack(const float *__restrict__ x1, const float *__restrict__ x2, PosGeom g,
                                                  float *__restrict__ Fpm, float *__restrict__ Fcm, float *__restrict__ nrm,
                                                  __half *__restrict__ FpmH, __half *__restrict__ FcmH, PackExact ex) {
    // One CTA = one 32-position strip of ONE branch (blockIdx.z): the norm is per branch, so nothing couples the two, and half
    // the shared memory doubles the resident CTAs (ncu r02j: 23 warps per SM and 45 % issue activity at 4.5 TB/s -- the kernel
    // wanted more loads in flight, not fewer instructions).
    extern __shared__ float T[];                     // [Cp][33] pooled values of the strip, channels of this branch
    __shared__ float s_inv[32];
    __shared__ double s_part[kExact ? 8 : 1][kExact ? 32 : 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.y, br = blockIdx.z, p0 = blockIdx.x * 32, p = p0 + lane;
    const bool valid = p < g.N;
    const int Cr = br ? g.C2 : g.C1, Cp = br ? g.C2p : g.C1p, c0 = br ? g.C1p : 0;      // real / padded channels, first channel in Kc
    const float *xb = (br ? x2 : x1) + (size_t)b * Cr * g.H * g.W;

    if (g.k == 1) {
        // no pooling: one coalesced 128-byte row per (channel, strip); eight independent loads in flight per warp, running pointers
        const size_t hw = (size_t)g.H * g.W;
        const float *xr = xb + (size_t)warp * 8 * hw + (valid ? p : 0);
        float *tr = T + warp * 8 * 33 + lane;
        for (int c8 = warp * 8; c8 < Cp; c8 += 64, xr += 64 * hw, tr += 64 * 33) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = (valid && c8 + u < Cr) ? ldg_stream_f32(xr + u * hw) : 0.f;
#pragma unroll
            for (int u = 0; u < 8; ++u) tr[u * 33] = v[u];
        }
    } else {
        // k x k mean (FALoss.py:23-24).  With k % 4 == 0 every window row is k/4 aligned 16-byte vectors and a warp's 32
        // neighbouring windows form one contiguous run per input row; all vectors of up to 8 rows are issued before summing.
        const int py = valid ? p / g.w : 0, px = valid ? p - py * g.w : 0;
        const float inv_kk = 1.f / (float)(g.k * g.k);
        const bool vec4 = (g.k % 4 == 0) && (g.W % 4 == 0) && ((reinterpret_cast<uintptr_t>(x1) | reinterpret_cast<uintptr_t>(x2)) & 15) == 0;
        for (int c = warp; c < Cp; c += 8) {
            float v = 0.f;
            if (valid && c < Cr) {
                const float *x = xb + ((size_t)c * g.H + (size_t)py * g.k) * g.W + (size_t)px * g.k;
                float s = 0.f;
                if (vec4 && g.k == 8) {
                    float4 r[16];
#pragma unroll
                    for (int dy = 0; dy < 8; ++dy) {
                        const uint4 a = ldg_stream_u4(x + (size_t)dy * g.W), bq = ldg_stream_u4(x + (size_t)dy * g.W + 4);
                        r[2 * dy] = make_float4(__uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(a.z), __uint_as_float(a.w));
                        r[2 * dy + 1] = make_float4(__uint_as_float(bq.x), __uint_as_float(bq.y), __uint_as_float(bq.z), __uint_as_float(bq.w));
                    }
#pragma unroll
                    for (int q = 0; q < 16; ++q) { s += r[q].x; s += r[q].y; s += r[q].z; s += r[q].w; }
                } else if (vec4) {
                    for (int dy = 0; dy < g.k; ++dy)
                        for (int q = 0; q < g.k / 4; ++q) {
                            const uint4 a = ldg_stream_u4(x + (size_t)dy * g.W + 4 * q);
                            s += __uint_as_float(a.x); s += __uint_as_float(a.y); s += __uint_as_float(a.z); s += __uint_as_float(a.w);
                        }
                } else {
                    for (int dy = 0; dy < g.k; ++dy)
                        for (int dx = 0; dx < g.k; ++dx) s += __ldg(x + (size_t)dy * g.W + dx);
                }
                v = s * inv_kk;
            }
            T[c * 33 + lane] = v;
        }
    }
    __syncthreads();
    if (kExact) {
        // FP64 sum of squares, the channels spread over the eight warps (fixed order: deterministic)
        double s1 = 0.0;
        {
            const float *tr = T + warp * 33 + lane;
#pragma unroll 4
            for (int c = warp; c < Cp; c += 8, tr += 8 * 33) { const double t = (double)*tr; s1 = fma(t, t, s1); }
        }
        s_part[warp][lane] = s1;
        __syncthreads();
        if (warp == 0) {
            double sq = 0.0;
            for (int i = 0; i < 8; ++i) sq += s_part[i][lane];
            const double n = sqrt(sq), inv = 1.0 / fmax(n, 1e-12);
            s_inv[lane] = (float)inv;
            nrm[((size_t)b * 2 + br) * g.Npad + p] = (float)n;
            ex.inv64[((size_t)b * 2 + br) * g.Npad + p] = inv;
        }
        if (blockIdx.x == 0 && blockIdx.y == 0 && br == 0 && threadIdx.x < 4) ex.stats[threadIdx.x] = 0ull;
    } else if (warp == 0) {                           // per-position L2 norm over the channels of the branch
        float s = 0.f;
        for (int c = 0; c < Cp; ++c) { const float t = T[c * 33 + lane]; s = fmaf(t, t, s); }
        const float n = sqrtf(s);
        s_inv[lane] = 1.f / fmaxf(n, 1e-12f);
        nrm[((size_t)b * 2 + br) * g.Npad + p] = n;
    }
    __syncthreads();
    // Stores with running pointers (ncu r02g: two thirds of this kernel's instructions were index arithmetic and branch selects
    // around the loads / stores of these loops).
    {                                                 // channel-major rows: 128 (fp32) / 64 (FP16) contiguous bytes per warp store
        const size_t o0 = ((size_t)b * g.Kc + c0 + warp) * g.Npad + p, ostep = (size_t)8 * g.Npad;
        const float *tr = T + warp * 33 + lane;
        const float il = s_inv[lane];
        if (Fcm) {
            float *pc = Fcm + o0;
            for (int c = warp; c < Cp; c += 8, tr += 8 * 33, pc += ostep) *pc = round_tf32(*tr * il);
            tr = T + warp * 33 + lane;
        }
        if (FcmH) {
            __half *ph = FcmH + o0;
#pragma unroll 4
            for (int c = warp; c < Cp; c += 8, tr += 8 * 33, ph += ostep) *ph = __float2half_rn(*tr * il);
        }
    }
#pragma unroll 1
    for (int q = warp; q < 32; q += 8) {              // position-major rows: the branch's Cp contiguous channels per position
        const float iq = s_inv[q];
        const float *tq = T + q;
        const size_t ro = ((size_t)b * g.Npad + p0 + q) * g.Kc + c0;
        if (Fpm) {
            float *dst = Fpm + ro, *dst_lo = dst + (size_t)g.B * g.Npad * g.Kc;
            for (int c = lane; c < Cp; c += 32) {
                const float f = tq[c * 33] * iq, hi = round_tf32(f);
                dst[c] = hi;
                if (g.split) dst_lo[c] = round_tf32(f - hi);
            }
        }
        if (FpmH) {                                   // FP16 copy: two channels per lane, 128 bytes per warp store
            __half2 *dh = reinterpret_cast<__half2 *>(FpmH + ro) + lane;
            const float *t2 = tq + 2 * lane * 33;
#pragma unroll 4
            for (int c = 2 * lane; c < Cp; c += 64, dh += 32, t2 += 64 * 33) *dh = __floats2half2_rn(t2[0] * iq, t2[33] * iq);
        }
        if (kExact) {                                 // raw pooled features, position-major: what the FP64 re-decision reads
            float *dp = ex.Ppm + ro + lane;
            const float *t1 = tq + lane * 33;
#pragma unroll 4
            for (int c = lane; c < Cp; c += 32, dp += 32, t1 += 32 * 33) *dp = *t1;
        }
    }
}

// Exact signs: the tie threshold of one sample.  A single tensor-core pass rounds both operands of every product to an 11-bit
// significand (relative error r ~ 2.1e-4 rms), so D_ij carries an error of variance 2 r^2 sum_c (f_ic f_jc)^2, which for
// positions with independent features is 2 r^2 sum_c mu_c^2, mu_c = mean_i f_ic^2 (summed over both branches; e.g. 2/C for
// relu(randn) features).  mu_c is estimated from up to 64 evenly spaced positions in a fixed order (deterministic);
// `floor2` is the variance of the FP32 accumulation noise, the only term left for the 3xTF32 split.
// tau = ksigma * sqrt(2 r^2 sum_c mu_c^2 + floor2).
__global__ void __launch_bounds__(1024) fa_pos_tau(PosGeom g, const float *__restrict__ Ppm, const double *__restrict__ inv64,
                                                   float r2, float floor2, float ksigma, float *__restrict__ tau) {
    __shared__ float scratch[33];
    __shared__ float s_mu[512];
    const int b = blockIdx.x, c = threadIdx.x & 511, grp = threadIdx.x >> 9;         // two position groups x one thread per channel
    const int M = g.N < 64 ? g.N : 64, step = g.N / M;
    float mu = 0.f;
    if (c < g.Kc) {
        const double *inv = inv64 + ((size_t)b * 2 + (c >= g.C1p)) * g.Npad;
#pragma unroll 8
        for (int m = grp; m < M; m += 2) {
            const int i = m * step;
            const float f = Ppm[((size_t)b * g.Npad + i) * g.Kc + c] * (float)inv[i];
            mu = fmaf(f, f, mu);
        }
    }
    if (grp == 1) s_mu[c] = mu;
    __syncthreads();
    if (grp == 0) mu = (mu + s_mu[c]) / (float)M; else mu = 0.f;
    const float v = block_sum(mu * mu, scratch);
    if (threadIdx.x == 0) tau[b] = ksigma * sqrtf(2.f * r2 * v + floor2);
}

// ---------------------------------------------------------------------------------------------------------------
// the tile engine
// ---------------------------------------------------------------------------------------------------------------
struct PosArgs {
    const float *Fpm;
    const __half *FpmH;  // FP16 form: the only position-major copy that exists
    const float *nrm;
    float *dP;
    float *opart;       // jsplit > 1: raw partial accumulators (jsplit, B, Npad, Kc)
    double *partials;
    unsigned *ticket;
    double *sum_out;
    float *loss_out;
    double loss_div;
    float grad_scale;   // 2 / Z
    // fused forward + backward without pooling (k == 1): the finished gradient goes straight to dX, scaled by *go
    float *dx[2];
    const float *go;
    int direct;
    // exact signs: per-sample tie threshold, per-row tie counters and lists (column | sign bit of the tensor-core value)
    const float *tau;
    unsigned *fcnt, *fent;
    uint4 *sb;          // two-pass form: the sign planes
    // two-pass form: pass A derives the tie threshold itself while its operand rows load (the arithmetic of fa_pos_tau)
    const float *Ppm;
    const double *inv64;
    float *tau_out;
    float tau_r2, tau_floor2, tau_ksigma;
};

#ifdef DSRL_POS_TIMING
#define TWAIT(acc, stmt) do { const long long _t = clock64(); stmt; acc += clock64() - _t; } while (0)
#else
#define TWAIT(acc, stmt) do { stmt; } while (0)
#endif

// Epilogue role, shared by the single-CTA and the CTA-pair tile kernels (warps 2..9 of a CTA): per column tile turn D into
// |D| (loss partial) and sign(D) (in place, operand of the gradient MMAs); after the last tile finish the gradient
// accumulator (normalisation Jacobian, or the raw partial rows when the column range is split); finally the loss.
// kPair: the barrier the MMA issuer waits on lives in the leader CTA of the pair.
struct EpiCtx {
    const unsigned char *fsm = nullptr;   // FP16 form: the CTA's own feature rows of its channel group in shared memory (TMA boxes of 128 rows
                                          // x 64 channels, 128-byte swizzle), or nullptr: read them from global memory
    uint64_t *d_full, *p_full, *o_full;
    double *red;          // [16]
    int *flag;
    float *proj;          // [2][128] partial projections of the two column halves
    uint32_t tmem;
    int itile, js, grp, b, j0, nt, gN, gbeg, T;
    int part_index, nparts;     // this CTA's slot among the loss partials (-1: it has none), number of slots
    int sub, nsub;              // exact signs: this CTA's sub-list pair (sub + column half) among a row's nsub sub-lists
};

// four consecutive channels (cl = channel - first channel of the group, a multiple of 4) of row r from the swizzled boxes
__device__ __forceinline__ float4 load_f4_smem(const unsigned char *fsm, int r, int cl) {
    const int box = cl >> 6, k = (cl & 63) >> 3;                     // 16-byte chunk k of the 128-byte row, stored at chunk k ^ (r & 7)
    const uint2 q = *reinterpret_cast<const uint2 *>(fsm + (size_t)box * kBoxBytes + r * 128 + ((k ^ (r & 7)) << 4) + ((cl & 4) << 1));
    const float2 lo = __half22float2(*reinterpret_cast<const __half2 *>(&q.x)), hi = __half22float2(*reinterpret_cast<const __half2 *>(&q.y));
    return make_float4(lo.x, lo.y, hi.x, hi.y);
}
// four consecutive channels of the normalised feature row at channel c (fp32 copy, or the FP16 copy of the FP16 form)
template <bool kHalf>
__device__ __forceinline__ float4 load_f4(const float *frow, const __half *frow_h, int c) {
    if (!kHalf) return __ldg(reinterpret_cast<const float4 *>(frow + c));
    const uint2 q = __ldg(reinterpret_cast<const uint2 *>(frow_h + c));
    const float2 lo = __half22float2(*reinterpret_cast<const __half2 *>(&q.x)), hi = __half22float2(*reinterpret_cast<const __half2 *>(&q.y));
    return make_float4(lo.x, lo.y, hi.x, hi.y);
}

// After the last column tile: the gradient accumulator -> raw rows (finished by fa_pos_finish) or the normalisation Jacobian
// in place; then this CTA's loss partial, the whole loss being finished in a fixed order by the last CTA to arrive.
template <bool kGrad, bool kHalf>
__device__ __forceinline__ void epilogue_finish(const PosGeom &g, const PosArgs &a, const EpiCtx &c, double acc) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint64_t *o_full = c.o_full;
    double *red = c.red;
    int *flag = c.flag;
    float *projbuf = c.proj;
    const uint32_t tmem = c.tmem;
    const int itile = c.itile, js = c.js, b = c.b, gN = c.gN, gbeg = c.gbeg;
    const int q = warp & 3, r = q * 32 + lane;
    const int half = (warp - 2) >> 2;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    if (kGrad && g.raw_o) {
        // partial accumulator of this column share: raw rows to global memory, finished by fa_pos_finish
        mbar_wait(o_full, 0, 7);
        fence_after_sync();
        const int row = itile * kTile + r;
        float *orow = a.opart + (((size_t)js * g.B + b) * g.Npad + row) * g.Kc + gbeg;
        for (int c0 = 32 * half; c0 < gN; c0 += 64) {
            uint32_t v[32];
            tmem_ld32(tmem + lane_addr + (uint32_t)c0, v);
            tmem_ld_wait();
#pragma unroll
            for (int e4 = 0; e4 < 8; ++e4)
                reinterpret_cast<uint4 *>(orow + c0)[e4] = make_uint4(v[e4 * 4], v[e4 * 4 + 1], v[e4 * 4 + 2], v[e4 * 4 + 3]);
        }
    } else if (kGrad) {
        // normalisation Jacobian of the gradient accumulator, stored channel-major (coalesced along positions)
#ifdef DSRL_POS_TIMING
        const long long te0 = clock64();
#endif
        mbar_wait(o_full, 0, 7);
        fence_after_sync();
#ifdef DSRL_POS_TIMING
        const long long te1 = clock64();
        long long te2 = 0;
#endif
        const int row = itile * kTile + r;
        const float *frow = a.Fpm + ((size_t)b * g.Npad + row) * g.Kc;
        const __half *frow_h = a.FpmH + ((size_t)b * g.Npad + row) * g.Kc;
        for (int br = 0; br < 2; ++br) {
            const int cb = br ? g.C1p : 0, ce = br ? g.Kc : g.C1p;         // channel range of the branch
            if (cb < gbeg || ce > gbeg + gN) continue;                     // not in this CTA's group
            const float n = a.nrm[((size_t)b * 2 + br) * g.Npad + row];
            const float sgn = br ? -a.grad_scale : a.grad_scale;           // dL/dFh2 = -2 Fh2 Sigma
            float proj = 0.f;
            for (int c0 = cb + 32 * half; c0 < ce; c0 += 64) {
                uint32_t v[32];
                tmem_ld32(tmem + lane_addr + (uint32_t)(c0 - gbeg), v);
                tmem_ld_wait();
#pragma unroll
                for (int e4 = 0; e4 < 8; ++e4) {
                    const float4 f = (kHalf && c.fsm) ? load_f4_smem(c.fsm, r, c0 + 4 * e4 - gbeg) : load_f4<kHalf>(frow, frow_h, c0 + 4 * e4);
                    proj = fmaf(f.x, __uint_as_float(v[e4 * 4 + 0]), proj);
                    proj = fmaf(f.y, __uint_as_float(v[e4 * 4 + 1]), proj);
                    proj = fmaf(f.z, __uint_as_float(v[e4 * 4 + 2]), proj);
                    proj = fmaf(f.w, __uint_as_float(v[e4 * 4 + 3]), proj);
                }
            }
#ifdef DSRL_POS_TIMING
            te2 = clock64();
#endif
            // the two column halves of a row each hold part of the projection
            asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
            projbuf[half * kTile + r] = proj;
            asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
            proj = projbuf[r] + projbuf[kTile + r];
            const bool dead = !(n > 1e-12f);                                // F/eps branch of the clamp: no projection
            const float scale = sgn / fmaxf(n, 1e-12f);
            const float gmul = a.direct ? __ldg(a.go) : 1.f;            // applied as a second multiply: the bits fa_pos_unpool would produce
            if (dead) proj = 0.f;
            const int Cr = br ? g.C2 : g.C1;
            for (int c0 = cb + 32 * half; c0 < ce; c0 += 64) {
                uint32_t v[32];
                tmem_ld32(tmem + lane_addr + (uint32_t)(c0 - gbeg), v);
                tmem_ld_wait();
                // dP (channel-major, padded), or -- fused forward + backward without pooling -- dX itself: (B, C, N), real
                // channels and positions only
                const size_t pitch = a.direct ? (size_t)g.N : (size_t)g.Npad;
                float *dst = a.direct ? a.dx[br] + ((size_t)b * Cr + (c0 - cb)) * pitch + row
                                      : a.dP + ((size_t)b * g.Kc + c0) * pitch + row;
                const int nreal = a.direct ? (row < g.N ? Cr - (c0 - cb) : 0) : 32;      // channels of this strip that exist in dX
#pragma unroll
                for (int e4 = 0; e4 < 8; ++e4) {
                    const float4 f = (kHalf && c.fsm) ? load_f4_smem(c.fsm, r, c0 + 4 * e4 - gbeg) : load_f4<kHalf>(frow, frow_h, c0 + 4 * e4);
                    if (e4 * 4 + 0 < nreal) dst[(size_t)(e4 * 4 + 0) * pitch] = (__uint_as_float(v[e4 * 4 + 0]) - f.x * proj) * scale * gmul;
                    if (e4 * 4 + 1 < nreal) dst[(size_t)(e4 * 4 + 1) * pitch] = (__uint_as_float(v[e4 * 4 + 1]) - f.y * proj) * scale * gmul;
                    if (e4 * 4 + 2 < nreal) dst[(size_t)(e4 * 4 + 2) * pitch] = (__uint_as_float(v[e4 * 4 + 2]) - f.z * proj) * scale * gmul;
                    if (e4 * 4 + 3 < nreal) dst[(size_t)(e4 * 4 + 3) * pitch] = (__uint_as_float(v[e4 * 4 + 3]) - f.w * proj) * scale * gmul;
                }
            }
        }
#ifdef DSRL_POS_TIMING
        if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 64) {
            long long *tm = reinterpret_cast<long long *>(a.sum_out) + 8;
            tm[20] = te1 - te0; tm[21] = te2 - te1; tm[22] = clock64() - te2;
        }
#endif
    }

    // loss: per-CTA partial, finished in a fixed order by the last CTA to arrive (deterministic)
    if (c.part_index >= 0) {
        const int et = threadIdx.x - 64;                          // index among the epilogue threads
        double tot = warp_sum(acc);
        if (lane == 0) red[et >> 5] = tot;
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
        if (et == 0) {
            double s = 0.0;
            for (int i = 0; i < kEpiWarps; ++i) s += red[i];
            a.partials[c.part_index] = s;
            __threadfence();
            const unsigned nparts = (unsigned)c.nparts;
            *flag = atomicInc(a.ticket, nparts - 1) == nparts - 1;      // self-resetting
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
        if (*flag) {
            __threadfence();
            const int nparts = c.nparts;
            double s = 0.0;
            for (int i = et; i < nparts; i += kEpiThreads) s += __ldcg(a.partials + i);
            s = warp_sum(s);
            asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
            if (lane == 0) red[kEpiWarps + (et >> 5)] = s;
            asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
            if (et == 0) {
                double all = 0.0;
                for (int i = 0; i < kEpiWarps; ++i) all += red[kEpiWarps + i];
                *a.sum_out = all;
                *a.loss_out = (float)(all / a.loss_div);
            }
        }
    }
}

template <bool kGrad, bool kPair, bool kHalf = false>
__device__ __forceinline__ void epilogue_role(const PosGeom &g, const PosArgs &a, const EpiCtx &c) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint64_t *d_full = c.d_full, *p_full = c.p_full;
    const uint32_t tmem = c.tmem;
    const int itile = c.itile, grp = c.grp, b = c.b, j0 = c.j0, nt = c.nt;
    {
    // ===================================== epilogue warps =====================================
    const int q = warp & 3, r = q * 32 + lane;                 // TMEM lane quarter of this warp, row inside the tile
    const int half = (warp - 2) >> 2;                          // which half of the columns / channel chunks this warp converts
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    double acc = 0.0;
    float facc = 0.f;
    long long w_d = 0;
    const long long t_begin = clock64();
    // exact signs: entries with 0 < |D| < tau are listed for fa_pos_finish (once per entry: the channel groups of a row tile
    // compute the same D, group 0 lists them)
    // The list of a row is cut into private sub-lists, one per thread that converts part of the row (column half x column
    // share [x channel-group CTA when the groups split the column tiles]): no atomics, no round trip on the conversion path -- an appended entry is one fire-and-forget store.
    const bool listing = kGrad && g.exact && grp == 0;
    const float tau = listing ? __ldg(a.tau + b) : 0.f;                    // 0: nothing is ever listed
    const size_t gsub = ((size_t)b * g.Npad + (size_t)itile * kTile + r) * c.nsub + (size_t)(c.sub + half);
    unsigned nlisted = 0;
    // (the FP64 add of the tile sums sits in an outer loop: as `if ((jj & 7) == 7)` the compiler predicates it into a DADD per tile)
    for (int jj0 = 0; jj0 < nt; jj0 += 8) {
    const int jj1 = min(jj0 + 8, nt);
#pragma unroll 1
    for (int jj = jj0; jj < jj1; ++jj) {
        const int buf = jj & 1, j = j0 + jj;
        TWAIT(w_d, mbar_wait(&d_full[buf], (jj >> 1) & 1, 3));
        fence_after_sync();
        const bool diag = j == itile;
        float tsum = 0.f;
#pragma unroll 1
        for (int cg = 2 * half; cg < 2 * half + 2; ++cg) {
            uint32_t v[32];
            const uint32_t taddr = tmem + lane_addr + kColD + (uint32_t)(buf * kTile + cg * 32);
            tmem_ld32(taddr, v);
            tmem_ld_wait();
            if (diag && cg == q) {                               // S_ii = 1 in both branches: a structural tie
#pragma unroll
                for (int e = 0; e < 32; ++e) if (e == lane) v[e] = 0u;
            }
            // |D| sum; the smallest magnitude tells whether any entry is exactly zero (the forced diagonal, exact ties), the
            // only entries whose sign is 0: everything else takes the short path, sign bit OR 1.0 -- one logic op per entry
            // (FP16: per pair of entries) instead of a compare + select each
            float zmin = 3.0e38f;
#pragma unroll
            for (int e = 0; e < 32; ++e) {
                const float x = __uint_as_float(v[e]);
                tsum += fabsf(x);
                zmin = fminf(zmin, fabsf(x));
            }
            if (kGrad && zmin < tau) {
                // near ties of this strip (about one strip in forty per thread): exact zeros -- the forced diagonal, padding,
                // dead positions, identical branches -- keep sign 0 and are not listed
                uint32_t m = 0u, neg = 0u;
#pragma unroll
                for (int e = 0; e < 32; ++e) {
                    const float ax = fabsf(__uint_as_float(v[e]));
                    m |= (ax < tau && ax != 0.f) ? (1u << e) : 0u;
                    neg |= (v[e] >> 31) << e;
                }
                while (m) {
                    const int e = __ffs(m) - 1;
                    m &= m - 1;
                    if (nlisted < (unsigned)g.fsub) a.fent[gsub * g.fsub + nlisted] = (uint32_t)(j * kTile + cg * 32 + e) | (((neg >> e) & 1u) << 31);
                    ++nlisted;
                }
            }
            if (kGrad && !kHalf) {
                if (zmin != 0.f) {
#pragma unroll
                    for (int e = 0; e < 32; ++e) v[e] = (v[e] & 0x80000000u) | 0x3f800000u;
                } else {
#pragma unroll
                    for (int e = 0; e < 32; ++e) v[e] = (v[e] & 0x80000000u) | ((v[e] & 0x7fffffffu) ? 0x3f800000u : 0u);   // sign(x) as a TF32 value
                }
                tmem_st32(taddr, v);
            }
            if (kGrad && kHalf) {
                // FP16 sign tile, two positions per 32-bit column.  The columns a warp writes ([64*half, 64*half + 32) of
                // the tile) lie inside the column range it has already read, so the two halves never race.
                uint32_t pk[16];
                if (zmin != 0.f) {
#pragma unroll
                    for (int e = 0; e < 16; ++e) pk[e] = (__byte_perm(v[2 * e], v[2 * e + 1], 0x7632) & 0x80008000u) | 0x3c003c00u;
                } else {
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const uint32_t lo = v[2 * e], hi = v[2 * e + 1];
                        pk[e] = ((lo >> 16) & 0x8000u) | ((lo & 0x7fffffffu) ? 0x3c00u : 0u) |
                                ((((hi >> 16) & 0x8000u) | ((hi & 0x7fffffffu) ? 0x3c00u : 0u)) << 16);
                    }
                }
                tmem_st16(tmem + lane_addr + kColD + (uint32_t)(buf * kTile + half * 64 + (cg & 1) * 16), pk);
            }
        }
        if (kGrad) tmem_st_wait();
        fence_before_sync();
        __syncwarp();                     // every lane's tcgen05.st has completed and is fenced: one arrival per warp
        if (lane == 0) { if (kPair) mbar_arrive_leader(&p_full[buf]); else mbar_arrive(&p_full[buf]); }
        // tile sums are gathered in FP32 over eight tiles (<= 2 * 64 * 8 per thread: rounding ~1e-7 of the running sum) before
        // they enter the FP64 total -- a DADD per tile was a quarter of the epilogue warps' stall samples (FP64 pipe)
        facc += (kGrad || diag) ? tsum : 2.f * tsum;
    }
    acc += (double)facc;
    facc = 0.f;
    }
    if (listing) a.fcnt[gsub] = nlisted;
#ifdef DSRL_POS_TIMING
    if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 64) {
        long long *tm = reinterpret_cast<long long *>(a.sum_out) + 8;      // saved[64..192): role clocks of CTA (0,0,0)
        tm[6] = clock64() - t_begin; tm[7] = w_d;
    }
#else
    (void)t_begin; (void)w_d;
#endif

    epilogue_finish<kGrad, kHalf>(g, a, c, acc);
    }
}

// kGrad:     also accumulate the gradient contraction (otherwise loss only, tiles j >= i by symmetry)
// kSplit:    3xTF32 -- D = hi*hi + hi*lo + lo*hi with the lo parts as extra operand boxes
// kResident: the CTA's own operand rows Q_i stay in shared memory for all tiles (otherwise streamed with K_j)
// kHalf:     FP16 operands (kind::f16, K = 16; 64 elements per 128-byte operand row), see fa_pos_tiles_pair
template <bool kGrad, bool kSplit, bool kResident, bool kHalf = false>
__global__ void __launch_bounds__(kThreads, 1)
fa_pos_tiles(const __grid_constant__ CUtensorMap tm_pm, const __grid_constant__ CUtensorMap tm_cm, const PosGeom g, const PosArgs a) {
    extern __shared__ unsigned char smraw[];
    const uint32_t raw = smem_u32(smraw);
    unsigned char *sm = smraw + (((raw + 1023u) & ~1023u) - raw);     // 1024-byte aligned: swizzle-128B tiles
    static_assert(!kHalf || (kResident && !kSplit), "FP16 form: resident rows, single pass");
    constexpr int kElems = kHalf ? 64 : kChunk;                 // operand elements per 128-byte row
    constexpr int kSB = stage_boxes(kResident);                 // boxes per stage
    constexpr int kStageBytes = kSB * kBoxBytes;
    constexpr int kU = (kResident ? 0 : (kSplit ? 2 : 1)) + (kSplit ? 2 : 1);   // operand boxes per 32-channel chunk: [Q hi, Q lo,] K hi [, K lo]
    constexpr int kUPS = kSB / kU;                              // chunks per stage
    const int S = kHalf ? g.half1_stages : g.stages, nkc = kHalf ? g.nkh : g.nkc, nq = kResident ? nkc * (kSplit ? 2 : 1) : 0;
    unsigned char *qreg = sm;
    unsigned char *ring = sm + (size_t)nq * kBoxBytes;
    uint64_t *full = reinterpret_cast<uint64_t *>(ring + (size_t)S * kStageBytes);
    uint64_t *empty = full + S;
    uint64_t *q_full = empty + S;
    uint64_t *d_full = q_full + 1;      // [2] D tile complete in TMEM            (MMA -> epilogue)
    uint64_t *p_full = d_full + 2;      // [2] kGrad: sign tile written (epilogue -> MMA); else: D tile drained
    uint64_t *o_full = p_full + 2;      //     gradient accumulator complete      (MMA -> epilogue)
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(o_full + 1);
    double *red = reinterpret_cast<double *>(o_full + 2);               // [16]
    int *flag = reinterpret_cast<int *>(red + 2 * kEpiWarps);
    float *projbuf = reinterpret_cast<float *>(flag + 2);               // [2][128]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int T = g.tiles, js = kGrad ? blockIdx.x / T : 0;
    const int itile = blockIdx.x - js * T, grp = blockIdx.y, b = blockIdx.z;
    // column tiles of this CTA: gradient variant -> an equal share of all tiles; forward-only -> the symmetry
    // D_ij = D_ji: tiles j >= i with weight 2
    const int nt = kGrad ? T / g.jsplit : T - itile;
    const int j0 = kGrad ? js * nt : itile;
    const int gN = g.gcnt[grp], gbeg = g.gbeg[grp], nbox = (gN + kTile - 1) / kTile;
    const int row_q = b * g.Npad + itile * kTile, lo_rows = g.B * g.Npad;     // lo parts live lo_rows below the hi parts

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tm_pm);
        prefetch_tmap(&tm_cm);
        for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(q_full, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&d_full[i], 1); mbar_init(&p_full[i], kEpiWarps); }
        mbar_init(o_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = *tmem_slot;

    // Producer and issuer walk the ring of stages in the same order.  Both run their loops with the whole warp
    // (uniform control flow, operands in uniform registers); one elected lane issues the TMA / MMA / commit
    // instructions.  A stage carries one or two 16 KB boxes, so every barrier round trip feeds 4-12 MMAs.
    int slot = 0;
    uint32_t ph = 0;
#define RING_ADVANCE() do { if (++slot == S) { slot = 0; ph ^= 1; } } while (0)

    if (warp == 0) {
        // ===================================== TMA producer =====================================
        long long w_empty = 0;
        const long long t_begin = clock64();
        if (kResident) {
            if (elect_one()) {
                mbar_arrive_expect_tx(q_full, (uint32_t)nq * kBoxBytes);
                for (int kc = 0; kc < nq; ++kc)
                    tma_load_2d(qreg + (size_t)kc * kBoxBytes, &tm_pm, q_full, (kc % nkc) * kElems, row_q + (kc / nkc) * lo_rows);
            }
            __syncwarp();
        }
        // fills the current stage: waits for the slot, arms the barrier with the byte count, runs BODY (TMA issues into `dst`)
#define STAGE_FILL(nboxes, ...)                                                                          \
        do {                                                                                             \
            TWAIT(w_empty, mbar_wait(&empty[slot], ph ^ 1, 1));                                          \
            if (elect_one()) {                                                                           \
                unsigned char *dst = ring + (size_t)slot * kStageBytes;                                  \
                uint64_t *bar = &full[slot];                                                             \
                mbar_arrive_expect_tx(bar, (uint32_t)(nboxes) * kBoxBytes);                              \
                __VA_ARGS__                                                                              \
            }                                                                                            \
            __syncwarp();                                                                                \
            RING_ADVANCE();                                                                              \
        } while (0)
        // operand boxes of D(i, j): per 32-channel chunk [Q_i hi, Q_i lo,] K_j hi [, K_j lo]; kUPS chunks per stage
        auto load_k = [&](int j) {
            const int row_k = b * g.Npad + j * kTile;
            for (int kc0 = 0; kc0 < nkc; kc0 += kUPS) {
                const int nu = min(kUPS, nkc - kc0);
                STAGE_FILL(nu * kU, {
                    for (int u = 0; u < nu; ++u) {
                        const int c0 = (kc0 + u) * kElems;
                        unsigned char *d = dst + (size_t)u * kU * kBoxBytes;
                        if (!kResident) {
                            tma_load_2d(d, &tm_pm, bar, c0, row_q); d += kBoxBytes;
                            if (kSplit) { tma_load_2d(d, &tm_pm, bar, c0, row_q + lo_rows); d += kBoxBytes; }
                        }
                        tma_load_2d(d, &tm_pm, bar, c0, row_k); d += kBoxBytes;
                        if (kSplit) tma_load_2d(d, &tm_pm, bar, c0, row_k + lo_rows);
                    }
                });
            }
        };
        // B boxes of the gradient contraction, flattened over (32-position chunk jc, 128-channel box bx), kSB per stage
        auto load_v = [&](int j) {
            const int nb = (kTile / kElems) * nbox;          // 4 or 8 (FP16: 2 or 4)
            for (int i0 = 0; i0 < nb; i0 += kSB) {
                const int n = min(kSB, nb - i0);
                STAGE_FILL(n, {
                    for (int h = 0; h < n; ++h) {
                        const int jc = (i0 + h) / nbox, bx = (i0 + h) - jc * nbox;
                        tma_load_2d(dst + (size_t)h * kBoxBytes, &tm_cm, bar, j * kTile + jc * kElems, b * g.Kc + gbeg + bx * kTile);
                    }
                });
            }
        };
        if (kGrad) {
            load_k(j0);
            for (int jj = 0; jj < nt; ++jj) {
                if (jj + 1 < nt) load_k(j0 + jj + 1);
                load_v(j0 + jj);
            }
        } else {
            for (int jj = 0; jj < nt; ++jj) load_k(j0 + jj);
        }
#undef STAGE_FILL
#ifdef DSRL_POS_TIMING
        if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0) {
            long long *tm = reinterpret_cast<long long *>(a.sum_out) + 8;      // saved[64..192): role clocks of CTA (0,0,0)
            tm[0] = clock64() - t_begin; tm[1] = w_empty;
        }
#else
        (void)t_begin; (void)w_empty;
#endif
    } else if (warp == 1) {
        // ===================================== MMA issuer =====================================
        long long w_full = 0, w_p = 0, w_drain = 0;
        const long long t_begin = clock64();
        constexpr uint64_t kBoxDesc = kBoxBytes >> 4, kStageDesc = kStageBytes >> 4;    // in descriptor address units (16 B)
        const uint64_t ring_desc = smem_desc_sw128(smem_u32(ring)), q_desc = smem_desc_sw128(smem_u32(qreg));
        const uint32_t id_pos = kHalf ? idesc_f16(kTile, kTile, false) : idesc_tf32(kTile, kTile, false);
        const uint32_t id_neg = kHalf ? idesc_f16(kTile, kTile, true) : idesc_tf32(kTile, kTile, true);
        const int last_rows = gN - (nbox - 1) * kTile;                                   // last channel box may be narrower
        const uint32_t id_last = kHalf ? idesc_f16(kTile, last_rows, false) : idesc_tf32(kTile, last_rows, false);
        const uint32_t id_wide = kHalf ? idesc_f16(kTile, 2 * kTile, false) : idesc_tf32(kTile, 2 * kTile, false);
        const bool wide = nbox == 2 && last_rows == kTile;                               // 256 channels: N = 256 instructions
        const int q_neg = g.C1p / (kElems / 4);                                          // first K step of branch 2 (subtracted)
#define RING_TAKE(desc_out, slot_out)                                                      \
        do {                                                                               \
            TWAIT(w_full, mbar_wait(&full[slot], ph, 2));                                  \
            desc_out = ring_desc + (uint64_t)slot * kStageDesc;                            \
            slot_out = slot;                                                               \
            RING_ADVANCE();                                                                \
        } while (0)
#define MMA4_SS(dcol, ad, bd, q0, acc0)                                                                    \
        do {                                                                                               \
            mma_ss<kHalf>(dcol, (ad), (bd), (q0) >= q_neg ? id_neg : id_pos, acc0);                        \
            mma_ss<kHalf>(dcol, (ad) + 2, (bd) + 2, (q0) + 1 >= q_neg ? id_neg : id_pos, 1);               \
            mma_ss<kHalf>(dcol, (ad) + 4, (bd) + 4, (q0) + 2 >= q_neg ? id_neg : id_pos, 1);               \
            mma_ss<kHalf>(dcol, (ad) + 6, (bd) + 6, (q0) + 3 >= q_neg ? id_neg : id_pos, 1);               \
        } while (0)
        // D(i, j0+jj) -> TMEM columns kColD + (jj&1)*128
        auto gemm_d = [&](int jj) {
            const int buf = jj & 1;
            const uint32_t dcol = tmem + kColD + (uint32_t)buf * kTile;
            if (!kGrad) TWAIT(w_drain, mbar_wait(&p_full[buf], ((jj >> 1) & 1) ^ 1, 4));    // epilogue drained this buffer
            for (int kc0 = 0; kc0 < nkc; kc0 += kUPS) {
                uint64_t sd; int ss;
                RING_TAKE(sd, ss);
                fence_after_sync();
                if (elect_one()) {
#pragma unroll
                    for (int u = 0; u < kUPS; ++u) {
                        const int kc = kc0 + u;
                        if (kc < nkc) {
                            const uint64_t ub = sd + (uint64_t)(u * kU) * kBoxDesc;
                            const uint64_t a_hi = kResident ? q_desc + (uint64_t)kc * kBoxDesc : ub;
                            const uint64_t a_lo = kResident ? q_desc + (uint64_t)(nkc + kc) * kBoxDesc : ub + kBoxDesc;
                            const uint64_t b_hi = kResident ? ub : ub + (kSplit ? 2 : 1) * kBoxDesc;
                            const uint64_t b_lo = b_hi + kBoxDesc;
                            MMA4_SS(dcol, a_hi, b_hi, 4 * kc, kc != 0);
                            if (kSplit) {
                                MMA4_SS(dcol, a_hi, b_lo, 4 * kc, 1);
                                MMA4_SS(dcol, a_lo, b_hi, 4 * kc, 1);
                            }
                        }
                    }
                    umma_commit(&empty[ss]);
                    if (kc0 + kUPS >= nkc) umma_commit(&d_full[buf]);
                }
                __syncwarp();
            }
        };
        // O(i, :) += sign(D(i, j)) * Fcat_j     (A = the sign tile the epilogue left in the D columns)
        auto gemm_g = [&](int jj, bool last) {
            const int buf = jj & 1;
            const uint32_t pcol = tmem + kColD + (uint32_t)buf * kTile;
            TWAIT(w_p, mbar_wait(&p_full[buf], (jj >> 1) & 1, 5));
            const int nb = (kTile / kElems) * nbox;
            constexpr int kPCols = kHalf ? 64 : kChunk;      // sign-tile columns per V box: TF32 32; FP16 box jc = packed columns [64 jc, 64 jc + 32)
            for (int i0 = 0; i0 < nb; i0 += kSB) {
                uint64_t bd; int sb;
                RING_TAKE(bd, sb);
                fence_after_sync();
                if (elect_one()) {
                    if (wide) {
                        // boxes (jc, 0) and (jc, 1) are adjacent: one N = 256 instruction per K step
#pragma unroll
                        for (int pr = 0; pr < kSB / 2; ++pr) {
                            const int jc = (i0 >> 1) + pr;
                            const uint32_t acol = pcol + (uint32_t)(jc * kPCols);
                            const uint64_t bp = bd + (uint64_t)(2 * pr) * kBoxDesc;
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) mma_ts<kHalf>(tmem, acol + ks * 8, bp + 2 * ks, id_wide, (jj | jc | ks) != 0);
                        }
                    } else {
#pragma unroll
                        for (int h = 0; h < kSB; ++h) {
                            if (i0 + h < nb) {
                                const int jc = (i0 + h) / nbox, bx = (i0 + h) - jc * nbox;
                                const uint32_t id = bx == nbox - 1 ? id_last : id_pos;
#pragma unroll
                                for (int ks = 0; ks < 4; ++ks)
                                    mma_ts<kHalf>(tmem + (uint32_t)bx * kTile, pcol + (uint32_t)(jc * kPCols + ks * 8),
                                                  bd + (uint64_t)h * kBoxDesc + 2 * ks, id, (jj | jc | ks) != 0);
                            }
                        }
                    }
                    umma_commit(&empty[sb]);
                    if (last && i0 + kSB >= nb) umma_commit(o_full);
                }
                __syncwarp();
            }
        };
        if (kResident) mbar_wait(q_full, 0, 6);
        if (kGrad) {
            gemm_d(0);
            for (int jj = 0; jj < nt; ++jj) {
                if (jj + 1 < nt) gemm_d(jj + 1);   // keeps the tensor pipe busy while the epilogue turns D(j) into signs
                gemm_g(jj, jj == nt - 1);
            }
        } else {
            for (int jj = 0; jj < nt; ++jj) gemm_d(jj);
        }
#undef MMA4_SS
#undef RING_TAKE
#ifdef DSRL_POS_TIMING
        if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0) {
            long long *tm = reinterpret_cast<long long *>(a.sum_out) + 8;      // saved[64..192): role clocks of CTA (0,0,0)
            tm[2] = clock64() - t_begin; tm[3] = w_full; tm[4] = w_p; tm[5] = w_drain;
        }
#else
        (void)t_begin; (void)w_full; (void)w_p; (void)w_drain;
#endif
    } else {
        // ===================================== epilogue warps =====================================
        EpiCtx c;
        c.d_full = d_full; c.p_full = p_full; c.o_full = o_full; c.red = red; c.flag = flag; c.proj = projbuf; c.tmem = tmem;
        c.itile = itile; c.js = js; c.grp = grp; c.b = b; c.j0 = j0; c.nt = nt; c.gN = gN; c.gbeg = gbeg; c.T = T;
        // every channel group computes the same D: group 0 reports the loss partial and lists the near ties
        c.part_index = grp == 0 ? (js * (int)gridDim.z + b) * T + itile : -1;
        c.nparts = (int)(gridDim.x * gridDim.z);
        c.sub = 2 * js; c.nsub = g.fnsub;
        epilogue_role<kGrad, false, kHalf>(g, a, c);
    }
#undef RING_ADVANCE

    fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, kTmemCols);
}

// ---------------------------------------------------------------------------------------------------------------
// CTA-pair form of the gradient variant (cluster of 2, tcgen05 cta_group::2, M = 256)
// ---------------------------------------------------------------------------------------------------------------
// Two CTAs own two consecutive 128-row tiles.  The leader (even CTA) issues every MMA for both: A rows / D rows come from
// each CTA's own shared / tensor memory, and each CTA supplies only HALF of every B tile (64 of the 128 K_j rows, half of
// the V_j channels).  Per column tile a CTA therefore receives half the K_j / V_j boxes and the tensor cores read 6 KB
// instead of 8 KB of shared memory per D step and 4 KB instead of 8 KB per gradient step -- the single-CTA form is
// shared-memory-bandwidth bound.  The full barriers live in the leader (both CTAs' TMA loads count bytes
// there); empty / d_full / o_full are signalled in both CTAs by multicast commits; both epilogues arrive on the leader's
// p_full.
constexpr int kPairKBox = kBoxBytes / 2;           // 64 rows of K_j per CTA

// kHalf: FP16 operands (kind::f16, K = 16): a 128-byte operand row holds 64 channels / positions, so a box covers twice the
// K extent, every MMA does twice the work per shared-memory byte, and the CTA's own rows are always resident.
template <bool kSplit, bool kResident, bool kHalf = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
fa_pos_tiles_pair(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                  const __grid_constant__ CUtensorMap tm_v, const PosGeom g, const PosArgs a) {
    extern __shared__ unsigned char smraw[];
    const uint32_t raw = smem_u32(smraw);
    unsigned char *sm = smraw + (((raw + 1023u) & ~1023u) - raw);
    // per-CTA operand boxes of one 32-channel chunk: [Q hi 16 KB, Q lo 16 KB,] K hi 8 KB [, K lo 8 KB]; a stage holds
    // 32 KB (operand rows resident) or 48 KB (streamed) of them, or two V boxes
    static_assert(!kHalf || (kResident && !kSplit), "FP16 form: resident rows, single pass");
    constexpr int kParts = kSplit ? 2 : 1;
    constexpr int kElems = kHalf ? 64 : kChunk;                     // operand elements per 128-byte row
    constexpr int kVB = kTile / kElems;                             // V boxes per column tile: 4 (TF32) or 2 (FP16)
    constexpr int kUnitBytes = (kResident ? 0 : kParts * kBoxBytes) + kParts * kPairKBox;
    constexpr int kStageBytes = kResident ? 4 * kPairKBox : 2 * (kBoxBytes + kPairKBox);
    constexpr int kUPS = kStageBytes / kUnitBytes;                  // chunks per stage: 4, 2 (split) | 2, 1 (split)
    const int S = kHalf ? g.half_stages : g.pair_stages, nkc = kHalf ? g.nkh : g.nkc, nq = kResident ? nkc * kParts : 0;
    unsigned char *qreg = sm;
    unsigned char *ring = sm + (size_t)nq * kBoxBytes;
    uint64_t *full = reinterpret_cast<uint64_t *>(ring + (size_t)S * kStageBytes);
    uint64_t *empty = full + S;
    uint64_t *q_full = empty + S;
    uint64_t *d_full = q_full + 1;
    uint64_t *p_full = d_full + 2;
    uint64_t *o_full = p_full + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(o_full + 1);
    double *red = reinterpret_cast<double *>(o_full + 2);
    int *flag = reinterpret_cast<int *>(red + 2 * kEpiWarps);
    float *projbuf = reinterpret_cast<float *>(flag + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int T = g.tiles, js = blockIdx.x / T;
    const int itile = blockIdx.x - js * T, grp = blockIdx.y, b = blockIdx.z;          // T is even: the pair shares js
    const int nt = T / g.jsplit, j0 = js * nt;
    const int gN = g.gcnt[grp], gbeg = g.gbeg[grp], vrows = gN / 2;                   // this CTA's share of the V_j channels
    const uint32_t vbytes = (uint32_t)vrows * kChunk * 4;
    const int row_q = b * g.Npad + itile * kTile, lo_rows = g.B * g.Npad;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tm_q);
        prefetch_tmap(&tm_k);
        prefetch_tmap(&tm_v);
        for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(q_full, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&d_full[i], 1); mbar_init(&p_full[i], 2 * kEpiWarps); }       // one arrival per epilogue warp of both CTAs
        mbar_init(o_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc2(tmem_slot, kTmemCols);
    fence_before_sync();
    cluster_sync();                         // peer barriers are initialised before anything signals them
    fence_after_sync();
    const uint32_t tmem = *tmem_slot;

    int slot = 0;
    uint32_t ph = 0;
#define RING_ADVANCE() do { if (++slot == S) { slot = 0; ph ^= 1; } } while (0)

    if (warp == 0) {
        // ===================================== TMA producer (both CTAs, each for its own shared memory) =====================================
        long long w_empty = 0;
        const long long t_begin = clock64();
        if (kResident) {
            if (elect_one()) {
                if (leader) mbar_arrive_expect_tx(q_full, 2u * (uint32_t)nq * kBoxBytes);
                for (int kc = 0; kc < nq; ++kc)
                    tma_load_2d_pair(qreg + (size_t)kc * kBoxBytes, &tm_q, q_full, (kc % nkc) * kElems, row_q + (kc / nkc) * lo_rows);
            }
            __syncwarp();
        }
#define STAGE_FILL(bytes_per_cta, ...)                                                                   \
        do {                                                                                             \
            TWAIT(w_empty, mbar_wait(&empty[slot], ph ^ 1, 11));                                         \
            if (elect_one()) {                                                                           \
                unsigned char *dst = ring + (size_t)slot * kStageBytes;                                  \
                uint64_t *bar = &full[slot];                                                             \
                if (leader) mbar_arrive_expect_tx(bar, 2u * (uint32_t)(bytes_per_cta));                  \
                __VA_ARGS__                                                                              \
            }                                                                                            \
            __syncwarp();                                                                                \
            RING_ADVANCE();                                                                              \
        } while (0)
        auto load_k = [&](int j) {
            const int row_k = b * g.Npad + j * kTile + (int)rank * (kTile / 2);           // this CTA's 64 rows of K_j
            for (int kc0 = 0; kc0 < nkc; kc0 += kUPS) {
                const int nu = min(kUPS, nkc - kc0);
                STAGE_FILL(nu * kUnitBytes, {
                    for (int u = 0; u < nu; ++u) {
                        const int c0 = (kc0 + u) * kElems;
                        unsigned char *d = dst + (size_t)u * kUnitBytes;
                        if (!kResident) {
                            tma_load_2d_pair(d, &tm_q, bar, c0, row_q); d += kBoxBytes;
                            if (kSplit) { tma_load_2d_pair(d, &tm_q, bar, c0, row_q + lo_rows); d += kBoxBytes; }
                        }
                        tma_load_2d_pair(d, &tm_k, bar, c0, row_k); d += kPairKBox;
                        if (kSplit) tma_load_2d_pair(d, &tm_k, bar, c0, row_k + lo_rows);
                    }
                });
            }
        };
        auto load_v = [&](int j) {
            const int row_v = b * g.Kc + gbeg + (int)rank * vrows;                        // this CTA's half of the channels
            for (int jc0 = 0; jc0 < kVB; jc0 += 2) {
                STAGE_FILL(2 * vbytes, {
                    tma_load_2d_pair(dst, &tm_v, bar, j * kTile + jc0 * kElems, row_v);
                    tma_load_2d_pair(dst + kBoxBytes, &tm_v, bar, j * kTile + (jc0 + 1) * kElems, row_v);
                });
            }
        };
        load_k(j0);
        for (int jj = 0; jj < nt; ++jj) {
            if (jj + 1 < nt) load_k(j0 + jj + 1);
            load_v(j0 + jj);
        }
#undef STAGE_FILL
#ifdef DSRL_POS_TIMING
        if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0) {
            long long *tm = reinterpret_cast<long long *>(a.sum_out) + 8;      // saved[64..192): role clocks of CTA (0,0,0)
            tm[0] = clock64() - t_begin; tm[1] = w_empty;
        }
#else
        (void)t_begin; (void)w_empty;
#endif
    } else if (warp == 1) {
        // ===================================== MMA issuer (leader CTA only) =====================================
        if (leader) {
            long long w_full = 0, w_p = 0;
            const long long t_begin = clock64();
            constexpr uint64_t kBoxDesc = kBoxBytes >> 4, kKDesc = kPairKBox >> 4, kStageDesc = kStageBytes >> 4, kUnitDesc = kUnitBytes >> 4;
            const uint64_t ring_desc = smem_desc_sw128(smem_u32(ring)), q_desc = smem_desc_sw128(smem_u32(qreg));
            const uint32_t id_pos = kHalf ? idesc_f16(2 * kTile, kTile, false) : idesc_tf32(2 * kTile, kTile, false);
            const uint32_t id_neg = kHalf ? idesc_f16(2 * kTile, kTile, true) : idesc_tf32(2 * kTile, kTile, true);
            const uint32_t id_g = kHalf ? idesc_f16(2 * kTile, gN, false) : idesc_tf32(2 * kTile, gN, false);
            const int q_neg = g.C1p / (kElems / 4);                 // first K step (32 bytes of a row) of branch 2 (subtracted)
            // the four K steps of one box; q0 = index of the first among all K steps (the branch boundary is a multiple of
            // 32 channels, so it can fall inside an FP16 box but never inside a K step)
#define MMA4_SS(dcol, ad, bd, q0, acc0)                                                                        \
            do {                                                                                               \
                mma_ss_pair<kHalf>(dcol, (ad), (bd), (q0) >= q_neg ? id_neg : id_pos, acc0);                   \
                mma_ss_pair<kHalf>(dcol, (ad) + 2, (bd) + 2, (q0) + 1 >= q_neg ? id_neg : id_pos, 1);          \
                mma_ss_pair<kHalf>(dcol, (ad) + 4, (bd) + 4, (q0) + 2 >= q_neg ? id_neg : id_pos, 1);          \
                mma_ss_pair<kHalf>(dcol, (ad) + 6, (bd) + 6, (q0) + 3 >= q_neg ? id_neg : id_pos, 1);          \
            } while (0)
            auto gemm_d = [&](int jj) {
                const int buf = jj & 1;
                const uint32_t dcol = tmem + kColD + (uint32_t)buf * kTile;
                for (int kc0 = 0; kc0 < nkc; kc0 += kUPS) {
                    TWAIT(w_full, mbar_wait(&full[slot], ph, 12));
                    const uint64_t sd = ring_desc + (uint64_t)slot * kStageDesc;
                    const int ss = slot;
                    RING_ADVANCE();
                    fence_after_sync();
                    if (elect_one()) {
#pragma unroll
                        for (int u = 0; u < kUPS; ++u) {
                            const int kc = kc0 + u;
                            if (kc < nkc) {
                                const uint64_t ub = sd + (uint64_t)u * kUnitDesc;
                                const uint64_t a_hi = kResident ? q_desc + (uint64_t)kc * kBoxDesc : ub;
                                const uint64_t a_lo = kResident ? q_desc + (uint64_t)(nkc + kc) * kBoxDesc : ub + kBoxDesc;
                                const uint64_t b_hi = kResident ? ub : ub + kParts * kBoxDesc;
                                const uint64_t b_lo = b_hi + kKDesc;
                                MMA4_SS(dcol, a_hi, b_hi, 4 * kc, kc != 0);
                                if (kSplit) {
                                    MMA4_SS(dcol, a_hi, b_lo, 4 * kc, 1);
                                    MMA4_SS(dcol, a_lo, b_hi, 4 * kc, 1);
                                }
                            }
                        }
                        umma_commit_pair(&empty[ss]);
                        if (kc0 + kUPS >= nkc) umma_commit_pair(&d_full[buf]);
                    }
                    __syncwarp();
                }
            };
            auto gemm_g = [&](int jj, bool last) {
                const int buf = jj & 1;
                const uint32_t pcol = tmem + kColD + (uint32_t)buf * kTile;
                TWAIT(w_p, mbar_wait(&p_full[buf], (jj >> 1) & 1, 15));
                for (int jc0 = 0; jc0 < kVB; jc0 += 2) {
                    TWAIT(w_full, mbar_wait(&full[slot], ph, 13));
                    const uint64_t sd = ring_desc + (uint64_t)slot * kStageDesc;
                    const int ss = slot;
                    RING_ADVANCE();
                    fence_after_sync();
                    if (elect_one()) {
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            // TF32: 32 sign columns per box; FP16: box h = positions [64h, 64h+64) = 32 packed columns at 64h
                            const uint32_t acol = pcol + (uint32_t)((jc0 + h) * (kHalf ? 64 : kChunk));
                            const uint64_t bd = sd + (uint64_t)h * kBoxDesc;
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) mma_ts_pair<kHalf>(tmem, acol + ks * 8, bd + 2 * ks, id_g, (jj | (jc0 + h) | ks) != 0);
                        }
                        umma_commit_pair(&empty[ss]);
                        if (last && jc0 + 2 >= kVB) umma_commit_pair(o_full);
                    }
                    __syncwarp();
                }
            };
            if (kResident) mbar_wait(q_full, 0, 16);
            gemm_d(0);
            for (int jj = 0; jj < nt; ++jj) {
                if (jj + 1 < nt) gemm_d(jj + 1);
                gemm_g(jj, jj == nt - 1);
            }
#undef MMA4_SS
#ifdef DSRL_POS_TIMING
            if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0) {
                long long *tm = reinterpret_cast<long long *>(a.sum_out) + 8;      // saved[64..192): role clocks of CTA (0,0,0)
                tm[2] = clock64() - t_begin; tm[3] = w_full; tm[4] = w_p; tm[5] = 0;
            }
#else
            (void)t_begin; (void)w_full; (void)w_p;
#endif
        }
    } else {
        EpiCtx c;
        c.d_full = d_full; c.p_full = p_full; c.o_full = o_full; c.red = red; c.flag = flag; c.proj = projbuf; c.tmem = tmem;
        c.itile = itile; c.js = js; c.grp = grp; c.b = b; c.j0 = j0; c.nt = nt; c.gN = gN; c.gbeg = gbeg; c.T = T;
        c.part_index = grp == 0 ? (js * (int)gridDim.z + b) * T + itile : -1;
        c.nparts = (int)(gridDim.x * gridDim.z);
        c.sub = 2 * js; c.nsub = g.fnsub;
        epilogue_role<true, true, kHalf>(g, a, c);
    }
#undef RING_ADVANCE

    fence_before_sync();
    cluster_sync();                         // the leader's MMAs read the peer's shared / tensor memory until the very end
    if (warp == 1) tmem_dealloc2(tmem, kTmemCols);
}

// ---------------------------------------------------------------------------------------------------------------
// sign / is-zero bits -> packed FP16 sign values (pass B of the two-pass form)
// ---------------------------------------------------------------------------------------------------------------
// The bits of a 32-entry strip are laid out for a cheap expansion: bit e = entry 2e, bit 16 + e = entry 2e + 1 (e < 16), so
// that the packed FP16 pair e is ((bits << (15 - e)) & 0x80008000) | 1.0|1.0.
// (Round 2 first removed the D recompute of two channel groups with a cluster of four CTAs -- two CTA pairs splitting the
// column tiles and shipping sign bits through distributed shared memory, st.async + mbarrier -- which reached 11.7 ms at
// BASELINE configs[3]; it could only occupy 132 of the 148 SMs (profiles/r02b) and was superseded by the two-pass form.)
__device__ __forceinline__ void expand_signs(uint32_t neg, uint32_t zero, uint32_t (&pk)[16]) {
    if (zero == 0u) {
#pragma unroll
        for (int e = 0; e < 16; ++e) pk[e] = ((neg << (15 - e)) & 0x80008000u) | 0x3c003c00u;
    } else {
#pragma unroll
        for (int e = 0; e < 16; ++e) {
            const uint32_t one = ~(zero << (15 - e)) & 0x80008000u;           // bit 15 / 31 set where the entry is NOT zero
            pk[e] = ((neg << (15 - e)) & one) | ((one >> 1) - (one >> 5));    // 0x3c00 = 0x4000 - 0x0400 per nonzero half
        }
    }
}


#include "fa_position_ab.cuh"      // the symmetric two-pass form (fa_pos_dsign, fa_pos_grad, sign planes)

// ---------------------------------------------------------------------------------------------------------------
// exact signs: re-decide the listed near ties, correct the raw accumulator rows (or, two-pass form, the sign planes)
// ---------------------------------------------------------------------------------------------------------------
// The tile kernel took sign(D_ij) from the tensor-core value and listed every entry with 0 < |D_ij| < tau.  Here one warp per
// row i re-evaluates those entries from the raw pooled features,
//     D_ij = <P1_i, P1_j> / (n1_i n1_j) - <P2_i, P2_j> / (n2_i n2_j),
// first in FP32 FMA (a row of P_j is 2 KB: the pass is bound by those L2 gathers, ~1e-3 N^2 of them per sample): with unit
// vectors the rounding error of that evaluation is below (16 + 5) 2^-24 per branch whatever the data, so |D| > kSafe32
// settles the sign; the few per cent below it are redone in FP64, where products of two floats are exact.  Where the sign
// differs from the one the tensor cores used, (s_exact - s_used) * Fh_j is added to the accumulator row -- the gradient is
// linear in the sign tile, so this is the row the tensor cores would have produced with the exact signs.  Corrections are
// summed in 2^-30 fixed point (integer adds: the order of a list does not matter), so results are bit-repeatable.  A row
// with more near ties than its lists hold keeps the tensor-core signs for the overflow (counted).
struct ResolveArgs {
    const float *Ppm; const double *inv64; const float *tau; const unsigned *fcnt, *fent; unsigned long long *stats;
    float *opart;            // raw accumulator rows of column share 0: (B, Npad, Kc)
    uint4 *sb;               // two-pass form (kBits): the sign planes; the lists hold upper-triangle entries only
};
constexpr float kSafe32 = 3.0e-6f;

// kNU = ceil(Kc / 128): a lane holds channels 4*lane + 128*u .. +3 for u < kNU (Kc <= 512).  The kernel issues about one
// warp instruction per 10 bytes it gathers, so the instruction count matters as much as the bytes: addresses are one
// 64-bit multiply-add per row, the loads carry immediate offsets, both branches share one warp reduction.
// kBits (two-pass form): nothing is accumulated -- the exact sign of a listed entry (i, j), j > i, is written into the sign
// planes at (i, j) wherever it differs from the tensor-core sign, and inside a diagonal tile at (j, i) always (its lower half
// was converted from its own tensor-core values); off-diagonal blocks below the diagonal do not exist.
// Entries per round (independent row gathers in flight per warp) and CTAs per SM.  The pass is bound by the latency of its L2
// gathers, and warps hide it better than loads per warp do: measured at cfg[3] (batch 8, 5.0 M entries) R/CTAs = 8/1: 2.73 ms,
// 4/2: 1.56, 4/3 (spills): 1.79, 2/4: 1.32, 1/4: 1.24, 1/6 (spills): 1.41, 2/6: 1.87.
#ifndef DSRL_RESOLVE_MINB
#define DSRL_RESOLVE_MINB 4
#endif
#ifndef DSRL_RESOLVE_R
#define DSRL_RESOLVE_R 1
#endif
template <int kNU, bool kBits = false>
__global__ void __launch_bounds__(256, DSRL_RESOLVE_MINB) fa_pos_resolve(PosGeom g, ResolveArgs ex) {
    constexpr int kR = DSRL_RESOLVE_R;            // entries per round: kR * kNU independent 16-byte loads in flight per lane
    constexpr float kFix = 1073741824.f;          // 2^30
    __shared__ long long s_corr[kBits ? 1 : 8][kBits ? 1 : 4 * kNU][32];  // per warp: fixed-point correction of the lane's channels (few entries flip)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long grow_ll = (long long)blockIdx.x * 8 + warp;
    if (grow_ll >= (long long)g.B * g.Npad) return;
    const size_t grow = (size_t)grow_ll;
    const int b = (int)(grow / g.Npad), irow = (int)(grow - (size_t)b * g.Npad);
    const int nsub = g.fnsub;                     // <= 32
    // this row's sub-lists (lane s holds the count of sub-list s) as one sequence of n entries; two-pass form: only the column
    // chunks that exist for the row's pair of tiles were written
    const int nsub_row = kBits ? 4 * dsign_chunks(g.tiles, g.a_chunk, irow / (2 * kTile)) : nsub;
    const unsigned cnt_s = lane < nsub_row ? ex.fcnt[grow * nsub + lane] : 0u;
    const unsigned len_s = min(cnt_s, (unsigned)g.fsub);
    unsigned end_s = len_s;                       // inclusive prefix sum over the sub-lists
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, end_s, o); if (lane >= o) end_s += t; }
    const int n = (int)__shfl_sync(0xffffffffu, end_s, nsub - 1);
    const unsigned listed = (unsigned)warp_sum((int)cnt_s);
    if (listed == 0) return;
    const float tau = ex.tau[b];
    const double *inv1 = ex.inv64 + (size_t)b * 2 * g.Npad, *inv2 = inv1 + g.Npad;
    // the lane's slice of the raw rows of this sample: row j is at lane_base + j * Kc, chunk u 128 floats further; a lane whose
    // last chunk lies past Kc reads chunk 0 again and masks the product (Kc is a multiple of 32, so chunks are whole or absent)
    const float *lane_base = ex.Ppm + (size_t)b * g.Npad * g.Kc + 4 * lane;
    const bool last_ok = 4 * lane + 128 * (kNU - 1) < g.Kc;
    const int last_off = last_ok ? 128 * (kNU - 1) : 0;
    float4 pi[kNU];                               // (kept in registers: a shared-memory copy was slower at every occupancy tried)
    float bsel[kNU];                              // +1: the chunk belongs to branch 1, -1: branch 2, 0: absent
    {
        const float *ri = lane_base + (size_t)irow * g.Kc;
#pragma unroll
        for (int u = 0; u < kNU; ++u) {
            const int off = u == kNU - 1 ? last_off : 128 * u;
            float4 t = __ldg(reinterpret_cast<const float4 *>(ri + off));
            bsel[u] = (u == kNU - 1 && !last_ok) ? 0.f : (4 * lane + 128 * u < g.C1p ? 1.f : -1.f);
            if (bsel[u] == 0.f) t = make_float4(0.f, 0.f, 0.f, 0.f);
            pi[u] = t;
        }
    }
    const double i1 = inv1[irow], i2 = inv2[irow];
    if (!kBits) {
#pragma unroll
        for (int t = 0; t < 4 * kNU; ++t) s_corr[warp][t][lane] = 0;
    }
    uint32_t *sbw = kBits ? reinterpret_cast<uint32_t *>(ex.sb + (size_t)b * sign_blocks(g.tiles) * (kSignBlockBytes / 16)) : nullptr;
    unsigned n_fix = 0;
    float worst = 0.f;
    const unsigned *ent = ex.fent + grow * g.fcap;
    for (int e0 = 0; e0 < n; e0 += 32) {
        // lane l owns entry e0 + l of the sequence: which sub-list holds it, the entry, and the two normalisation weights
        const int e = e0 + lane;
        int sub = 0;
        unsigned start = 0;
        for (int t = 0; t < nsub; ++t) {
            const unsigned end_t = __shfl_sync(0xffffffffu, end_s, t);
            if (end_t <= (unsigned)e) { sub = t + 1; start = end_t; }
        }
        const uint32_t en_l = e < n ? __ldg(ent + (size_t)sub * g.fsub + (e - start)) : 0u;
        const int j_l = (int)(en_l & 0x7fffffffu);
        // weights of the two branches for this (i, j): +1 / (n1_i n1_j) and -1 / (n2_i n2_j), so that D is ONE weighted sum
        const float w1_l = e < n ? (float)(inv1[j_l] * i1) : 0.f, w2_l = e < n ? -(float)(inv2[j_l] * i2) : 0.f;
        const int m = min(32, n - e0);
        for (int h0 = 0; h0 < m; h0 += kR) {
            uint32_t en[kR];
            float4 pj[kR][kNU];
#pragma unroll
            for (int h = 0; h < kR; ++h) {
                en[h] = __shfl_sync(0xffffffffu, en_l, (h0 + h) & 31);
                const float *rj = lane_base + (size_t)(en[h] & 0x7fffffffu) * g.Kc;      // past the list: entry 0 = row 0, result unused
#pragma unroll
                for (int u = 0; u < kNU; ++u) pj[h][u] = __ldg(reinterpret_cast<const float4 *>(rj + (u == kNU - 1 ? last_off : 128 * u)));
            }
#pragma unroll
            for (int h = 0; h < kR; ++h) {
                if (h0 + h >= m) break;                                   // warp-uniform
                const float w1 = __shfl_sync(0xffffffffu, w1_l, (h0 + h) & 31), w2 = __shfl_sync(0xffffffffu, w2_l, (h0 + h) & 31);
                float d = 0.f;
#pragma unroll
                for (int u = 0; u < kNU; ++u) {
                    const float4 a = pi[u];
                    float t = a.x * pj[h][u].x;
                    t = fmaf(a.y, pj[h][u].y, t); t = fmaf(a.z, pj[h][u].z, t); t = fmaf(a.w, pj[h][u].w, t);
                    d = fmaf(t, bsel[u] > 0.f ? w1 : w2, d);              // pi is zero where the chunk is absent
                }
                d = warp_sum(d);
                int s_exact = d > 0.f ? 1 : -1;
                float dabs = fabsf(d);
                if (!(dabs > kSafe32)) {
                    // too close for FP32 (also NaN / overflow): exact products, FP64 sums, FP64 weights
                    const int j = (int)(en[h] & 0x7fffffffu);
                    double d1 = 0.0, d2 = 0.0;
#pragma unroll
                    for (int u = 0; u < kNU; ++u) {
                        const float4 a = pi[u];
                        const double t = (double)a.x * (double)pj[h][u].x + (double)a.y * (double)pj[h][u].y +
                                         (double)a.z * (double)pj[h][u].z + (double)a.w * (double)pj[h][u].w;
                        if (bsel[u] > 0.f) d1 += t; else d2 += t;
                    }
                    d1 = warp_sum(d1);
                    d2 = warp_sum(d2);
                    const double D = d1 * (i1 * inv1[j]) - d2 * (i2 * inv2[j]);
                    s_exact = D > 0.0 ? 1 : (D < 0.0 ? -1 : 0);
                    dabs = (float)fabs(D);
                }
                const int s_used = (en[h] >> 31) ? -1 : 1;
                if (kBits) {
                    const int j = (int)(en[h] & 0x7fffffffu);
                    const bool fix = s_exact != s_used;
                    if (fix) { ++n_fix; worst = fmaxf(worst, dabs / tau); }
                    // only the blocks (tile of i) <= (tile of j) exist: pass B reads them transposed for the lower triangle.  A
                    // diagonal tile holds both (i, j) and (j, i), each from its own tensor-core value: the mirror is set regardless
                    if (lane == 0) {
                        if (fix) sign_plane_set(sbw, g.tiles, irow, j, s_exact);
                        if ((j >> 7) == (irow >> 7)) sign_plane_set(sbw, g.tiles, j, irow, s_exact);
                    }
                    continue;
                }
                if (s_exact != s_used) {
                    ++n_fix;
                    worst = fmaxf(worst, dabs / tau);
                    const int j = (int)(en[h] & 0x7fffffffu);
                    const float f1 = (float)inv1[j] * kFix, f2 = (float)inv2[j] * kFix;      // Fh_j = P_j / n_j, in 2^-30 units
                    const long long dl = (long long)(s_exact - s_used);
#pragma unroll
                    for (int u = 0; u < kNU; ++u) {
                        const float w = bsel[u] > 0.f ? f1 : (bsel[u] < 0.f ? f2 : 0.f);
                        s_corr[warp][4 * u + 0][lane] += dl * (long long)__float2int_rn(pj[h][u].x * w);
                        s_corr[warp][4 * u + 1][lane] += dl * (long long)__float2int_rn(pj[h][u].y * w);
                        s_corr[warp][4 * u + 2][lane] += dl * (long long)__float2int_rn(pj[h][u].z * w);
                        s_corr[warp][4 * u + 3][lane] += dl * (long long)__float2int_rn(pj[h][u].w * w);
                    }
                }
            }
        }
    }
    if (!kBits && n_fix) {
        float *orow = ex.opart + grow * g.Kc + 4 * lane;
#pragma unroll
        for (int u = 0; u < kNU; ++u) {
            if (u == kNU - 1 && !last_ok) break;
            float4 o = *reinterpret_cast<float4 *>(orow + 128 * u);
            o.x += (float)s_corr[warp][4 * u + 0][lane] * (1.f / kFix); o.y += (float)s_corr[warp][4 * u + 1][lane] * (1.f / kFix);
            o.z += (float)s_corr[warp][4 * u + 2][lane] * (1.f / kFix); o.w += (float)s_corr[warp][4 * u + 3][lane] * (1.f / kFix);
            *reinterpret_cast<float4 *>(orow + 128 * u) = o;
        }
    }
    if (lane == 0) {
        atomicAdd(ex.stats + 0, (unsigned long long)listed);
        if (n_fix) atomicAdd(ex.stats + 1, (unsigned long long)n_fix);
        if (listed > (unsigned)n) atomicAdd(ex.stats + 2, (unsigned long long)(listed - (unsigned)n));
        if (n_fix) atomicMax(reinterpret_cast<unsigned *>(ex.stats + 3), __float_as_uint(worst));
    }
}

// ---------------------------------------------------------------------------------------------------------------
// finish: sum the partial accumulators (jsplit shares), apply the normalisation Jacobian, store dP channel-major (or dX
// itself: fused forward + backward without pooling)
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 3) fa_pos_finish(PosGeom g, const float *__restrict__ opart, const float *__restrict__ Fcm,
                                                    const __half *__restrict__ FcmH, const float *__restrict__ nrm, float grad_scale,
                                                    float *__restrict__ dP, float *dx1, float *dx2, const float *go) {
    extern __shared__ float T[];                     // [Kc][33] summed accumulator of a 32-position strip
    __shared__ float s_part[2][8][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.y, p0 = blockIdx.x * 32;
    // position-major rows in: a warp takes four rows, two at a time, 16-byte loads, all eight of a column share in flight at
    // once (a loop of dependent 4-byte loads ran this kernel at 0.6 TB/s); fixed summation order over the shares
#pragma unroll 1
    for (int i0 = 0; i0 < 4; i0 += 2) {
        float4 acc4[2][4];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int u = 0; u < 4; ++u) acc4[i][u] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int js = 0; js < g.jsplit; ++js) {
            float4 v[2][4];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const float *row = opart + (((size_t)js * g.B + b) * g.Npad + p0 + warp + 8 * (i0 + i)) * g.Kc;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int c = 4 * lane + 128 * u;
                    v[i][u] = c < g.Kc ? __ldcs(reinterpret_cast<const float4 *>(row + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    acc4[i][u].x += v[i][u].x; acc4[i][u].y += v[i][u].y; acc4[i][u].z += v[i][u].z; acc4[i][u].w += v[i][u].w;
                }
        }
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int c = 4 * lane + 128 * u, q = warp + 8 * (i0 + i);
                if (c < g.Kc) {
                    T[(c + 0) * 33 + q] = acc4[i][u].x; T[(c + 1) * 33 + q] = acc4[i][u].y;
                    T[(c + 2) * 33 + q] = acc4[i][u].z; T[(c + 3) * 33 + q] = acc4[i][u].w;
                }
            }
    }
    __syncthreads();
    // thread (warp, lane) owns position p0 + lane of the channels warp, warp + 8, ...: projection <Fh_i, O_i> as partial sums per
    // warp, combined in a fixed order; the features are read again for the output (they sit in L1 / L2 by then), which keeps the
    // kernel at three CTAs per SM
    auto feat = [&](int c) {
        const size_t o = ((size_t)b * g.Kc + c) * g.Npad + p0 + lane;
        return Fcm ? Fcm[o] : __half2float(FcmH[o]);      // FP16 form: only the FP16 copy exists
    };
    float p1 = 0.f, p2 = 0.f;
#pragma unroll 8
    for (int c = warp; c < g.Kc; c += 8) {
        const float f = feat(c);
        if (c < g.C1p) p1 = fmaf(f, T[c * 33 + lane], p1); else p2 = fmaf(f, T[c * 33 + lane], p2);
    }
    s_part[0][warp][lane] = p1;
    s_part[1][warp][lane] = p2;
    __syncthreads();
    float proj[2], scale[2];
#pragma unroll
    for (int br = 0; br < 2; ++br) {
        float sp = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) sp += s_part[br][w][lane];
        const float n = nrm[((size_t)b * 2 + br) * g.Npad + p0 + lane];
        proj[br] = n > 1e-12f ? sp : 0.f;             // F/eps branch of the clamp: no projection
        scale[br] = (br ? -grad_scale : grad_scale) / fmaxf(n, 1e-12f);
    }
    const float gmul = go ? __ldg(go) : 1.f;
#pragma unroll 8
    for (int c = warp; c < g.Kc; c += 8) {
        const int br = c >= g.C1p;
        const float val = (T[c * 33 + lane] - feat(c) * proj[br]) * scale[br];
        if (!go) { dP[((size_t)b * g.Kc + c) * g.Npad + p0 + lane] = val; continue; }
        // fused forward + backward without pooling: dX itself, real channels and positions only
        const int cc = br ? c - g.C1p : c, Cr = br ? g.C2 : g.C1;
        if (cc < Cr && p0 + lane < g.N) __stcs((br ? dx2 : dx1) + ((size_t)b * Cr + cc) * g.N + p0 + lane, val * gmul);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// backward proper
// ---------------------------------------------------------------------------------------------------------------
// One thread row (64 threads) per output row: the (b, c, y) decomposition and the source row are computed once per row,
// the inner loop is a broadcast load + one 16-byte store (a per-element 64-bit div/mod version ran at 27 % of the HBM peak).
template <int VEC>
__global__ void __launch_bounds__(256) fa_pos_unpool(PosGeom g, const float *__restrict__ dP, const float *__restrict__ grad_out,
                                                      float *__restrict__ dx1, float *__restrict__ dx2) {
    const float scale = __ldg(grad_out) / (float)(g.k * g.k);
    const long long rows1 = (long long)g.B * g.C1 * g.H, rows2 = (long long)g.B * g.C2 * g.H;
    const long long row = (long long)blockIdx.x * blockDim.y + threadIdx.y;
    if (row >= rows1 + rows2) return;
    const int br = row >= rows1;
    float *dx = br ? dx2 : dx1;
    if (!dx) return;
    const long long rr = br ? row - rows1 : row;
    const int Cr = br ? g.C2 : g.C1;
    const int y = (int)(rr % g.H);
    const long long bcq = rr / g.H;
    const int c = (int)(bcq % Cr), b = (int)(bcq / Cr);
    const int py = y / g.k;
    float *dst = dx + rr * g.W;
    const int wv = g.W / VEC;
    if (py >= g.h) {                                           // rows dropped by the floor of the pooling
        for (int xv = threadIdx.x; xv < wv; xv += blockDim.x) {
            if (VEC == 4) reinterpret_cast<float4 *>(dst)[xv] = make_float4(0.f, 0.f, 0.f, 0.f); else dst[xv] = 0.f;
        }
        return;
    }
    const float *src = dP + ((size_t)b * g.Kc + (br ? g.C1p : 0) + c) * g.Npad + (size_t)py * g.w;
    const int k = g.k, w = g.w;
    for (int xv = threadIdx.x; xv < wv; xv += blockDim.x) {
        if (VEC == 4) {
            const int x0 = xv * 4;
            float4 o;
            if (k == 1 && (w & 3) == 0) {                      // no pooling: a scaled copy, 16 bytes in, 16 bytes out
                const float4 v = __ldg(reinterpret_cast<const float4 *>(src + x0));
                o = x0 < w ? make_float4(v.x * scale, v.y * scale, v.z * scale, v.w * scale) : make_float4(0.f, 0.f, 0.f, 0.f);
            } else if (k % 4 == 0) {                           // the four outputs share one pooled cell
                const int px = x0 / k;
                const float v = px < w ? __ldg(src + px) * scale : 0.f;
                o = make_float4(v, v, v, v);
            } else {
                const int p0 = x0 / k, p1 = (x0 + 1) / k, p2 = (x0 + 2) / k, p3 = (x0 + 3) / k;
                o = make_float4(p0 < w ? __ldg(src + p0) * scale : 0.f, p1 < w ? __ldg(src + p1) * scale : 0.f,
                                p2 < w ? __ldg(src + p2) * scale : 0.f, p3 < w ? __ldg(src + p3) * scale : 0.f);
            }
            __stcs(reinterpret_cast<float4 *>(dst) + xv, o);         // written once, read by nobody here: streaming store
        } else {
            const int px = xv / k;
            dst[xv] = px < w ? __ldg(src + px) * scale : 0.f;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
std::once_flag g_encode_once;
EncodeTiledFn g_encode = nullptr;

EncodeTiledFn encode_fn() {
    std::call_once(g_encode_once, [] {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            g_encode = reinterpret_cast<EncodeTiledFn>(fn);
        (void)cudaGetLastError();
    });
    return g_encode;
}

// rows x cols matrix (fp32, or fp16 with `half`), row-major; box = box_rows rows x 128 bytes, 128-byte swizzle;
// columns past `cols` read as zero
int make_map(CUtensorMap *m, const void *base, uint64_t rows, uint64_t cols, int box_rows = kTile, bool half = false) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) DSRL_FAIL(DSRL_ERR_CUDA, "FA(position): cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {cols * (half ? 2 : 4)};
    const cuuint32_t box[2] = {(cuuint32_t)(half ? 2 * kChunk : kChunk), (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(m, half ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) DSRL_FAIL(DSRL_ERR_CUDA, "FA(position): cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
    return DSRL_OK;
}

template <typename K>
int opt_in_smem(K kern, size_t bytes) {
    if (bytes > 48 * 1024) DSRL_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return DSRL_OK;
}

}  // namespace

int fa_pos_backward(int precision, const float *, const float *, const void *saved_v, size_t saved_bytes, const float *grad_out,
                    float *dx1, float *dx2, int B, int C1, int C2, int H, int W, int k, int, void *, size_t, cudaStream_t st);

// precision = DSRL_PREC_{FP32,TF32,F16}, optionally | DSRL_PREC_EXACT_SIGNS
static bool geom_for(int precision, int B, int C1, int C2, int H, int W, int k, PosGeom &g) {
    const int base = precision & ~DSRL_PREC_EXACT_SIGNS;
    if (base != DSRL_PREC_TF32 && base != DSRL_PREC_FP32 && base != DSRL_PREC_F16) return false;
    return make_geom(B, C1, C2, H, W, k, base == DSRL_PREC_FP32, g, base == DSRL_PREC_F16, (precision & DSRL_PREC_EXACT_SIGNS) != 0);
}

size_t fa_pos_saved_bytes(int precision, int B, int C1, int C2, int H, int W, int k) {
    PosGeom g;
    if (!geom_for(precision, B, C1, C2, H, W, k, g)) return 0;
    return make_saved(g).total;
}

size_t fa_pos_workspace_bytes(int precision, int B, int C1, int C2, int H, int W, int k) {
    PosGeom g;
    if (!geom_for(precision, B, C1, C2, H, W, k, g)) return 0;
    return make_ws(g).total;
}

// operand-rounding model of the tie threshold (fa_pos_tau): relative rms error of one rounding to an 11-bit significand, and
// the variance of the FP32 accumulation noise of a D tile
static constexpr float kRound11 = 2.1e-4f;
static constexpr float kAccNoise = 1.0e-6f;

// go / dx1 / dx2 (all non-null, k == 1): fused forward + backward -- the gradient kernel writes dX directly (scaled by *go) and the
// saved blob's dP is not produced; *fused_out tells the caller that no backward launch is needed.
int fa_pos_forward_impl(int precision, const float *x1, const float *x2, int B, int C1, int C2, int H, int W, int k, int reduction,
                        int need_grad, float *loss_out, void *saved_v, size_t saved_bytes, void *ws_v, size_t ws_bytes, cudaStream_t st,
                        const float *go, float *dx1, float *dx2, int *fused_out) {
    const int base = precision & ~DSRL_PREC_EXACT_SIGNS;
    if (base != DSRL_PREC_TF32 && base != DSRL_PREC_FP32 && base != DSRL_PREC_F16)
        DSRL_FAIL(DSRL_ERR_UNSUPPORTED, "FA(position): precision must be TF32 (one tcgen05 kind::tf32 pass), F16 (kind::f16 operands) or FP32 (3xTF32 split), optionally | EXACT_SIGNS");
    PosGeom g;
    if (!geom_for(need_grad ? precision : base, B, C1, C2, H, W, k, g))        // the loss alone has no sign decisions
        DSRL_FAIL(DSRL_ERR_BAD_SHAPE, "FA(position): unsupported geometry B=%d C=(%d,%d) H=%d W=%d k=%d (channels per branch <= 256)", B, C1, C2, H, W, k);
    precision = base;
    const PosWs wo = make_ws(g);
    const PosSaved so = make_saved(g);
    if (saved_bytes < so.total) DSRL_FAIL(DSRL_ERR_BAD_ARG, "FA(position): saved blob too small (%zu < %zu)", saved_bytes, so.total);
    if (ws_bytes < wo.total) DSRL_FAIL(DSRL_ERR_BAD_ARG, "FA(position): workspace too small (%zu < %zu)", ws_bytes, wo.total);
    if ((reinterpret_cast<uintptr_t>(ws_v) & 15) || (reinterpret_cast<uintptr_t>(saved_v) & 15))
        DSRL_FAIL(DSRL_ERR_BAD_ARG, "FA(position): workspace / saved must be 16-byte aligned");
    unsigned char *ws = static_cast<unsigned char *>(ws_v), *saved = static_cast<unsigned char *>(saved_v);
    float *Fpm = reinterpret_cast<float *>(ws + wo.Fpm), *Fcm = reinterpret_cast<float *>(ws + wo.Fcm);
    float *nrm = reinterpret_cast<float *>(ws + wo.nrm);

    const size_t finish_smem = (size_t)g.Kc * 33 * 4, pack_smem = (size_t)(g.C1p > g.C2p ? g.C1p : g.C2p) * 33 * 4;
    int rc = opt_in_smem(fa_pos_pack<false>, pack_smem);
    if (rc) return rc;
    if ((rc = opt_in_smem(fa_pos_pack<true>, pack_smem))) return rc;
    __half *FpmH = g.half ? reinterpret_cast<__half *>(ws + wo.FpmH) : nullptr;
    __half *FcmH = g.half ? reinterpret_cast<__half *>(ws + wo.FcmH) : nullptr;
    PackExact pex;
    pex.Ppm = reinterpret_cast<float *>(ws + wo.Ppm);
    pex.inv64 = reinterpret_cast<double *>(ws + wo.inv64);
    pex.stats = reinterpret_cast<unsigned long long *>(saved + 8);
    float ksigma = 3.5f, tau_r = kRound11;
    ResolveArgs rex;
    rex.Ppm = pex.Ppm; rex.inv64 = pex.inv64; rex.stats = pex.stats;
    rex.fcnt = reinterpret_cast<unsigned *>(ws + wo.fcnt);
    rex.tau = reinterpret_cast<float *>(ws + wo.tau);
    rex.fent = reinterpret_cast<unsigned *>(ws + wo.fent);
    rex.opart = reinterpret_cast<float *>(ws + wo.opart);
    rex.sb = reinterpret_cast<uint4 *>(ws + wo.sb);
    // FP16 form: the FP16 copies are the only ones written (and read back by the normalisation Jacobian)
    if (g.exact) {
        fa_pos_pack<true><<<dim3(g.Npad / 32, B, 2), 256, pack_smem, st>>>(x1, x2, g, g.half ? nullptr : Fpm, g.half ? nullptr : Fcm, nrm, FpmH, FcmH, pex);
        if (const char *e = getenv("DSRL_POS_KSIGMA")) { const float v = (float)atof(e); if (v >= 0.f && v < 1e6f) ksigma = v; }   // tuning / test hook
        tau_r = g.split ? kRound11 * kRound11 : kRound11;      // 3xTF32: the products missing from the split are second order
        if (!(need_grad && g.ab)) {        // the two-pass form computes the threshold in pass A's prologue (one launch and ~20 us less)
            DSRL_LAUNCH_CHECK();
            fa_pos_tau<<<B, 1024, 0, st>>>(g, pex.Ppm, pex.inv64, tau_r * tau_r, kAccNoise * kAccNoise, ksigma, reinterpret_cast<float *>(ws + wo.tau));
        }
    } else {
        fa_pos_pack<false><<<dim3(g.Npad / 32, B, 2), 256, pack_smem, st>>>(x1, x2, g, g.half ? nullptr : Fpm, g.half ? nullptr : Fcm, nrm, FpmH, FcmH, pex);
    }
    DSRL_LAUNCH_CHECK();
    // the tile kernel's raw accumulator rows -> dP / dX (all variants share it)
    auto resolve = [&]() -> int {
        const long long rows = (long long)B * g.Npad;
        const unsigned blocks = (unsigned)((rows + 7) / 8);
        if (g.ab) {
            switch ((g.Kc + 127) / 128) {
                case 1: fa_pos_resolve<1, true><<<blocks, 256, 0, st>>>(g, rex); break;
                case 2: fa_pos_resolve<2, true><<<blocks, 256, 0, st>>>(g, rex); break;
                case 3: fa_pos_resolve<3, true><<<blocks, 256, 0, st>>>(g, rex); break;
                default: fa_pos_resolve<4, true><<<blocks, 256, 0, st>>>(g, rex); break;
            }
        } else {
            switch ((g.Kc + 127) / 128) {
                case 1: fa_pos_resolve<1><<<blocks, 256, 0, st>>>(g, rex); break;
                case 2: fa_pos_resolve<2><<<blocks, 256, 0, st>>>(g, rex); break;
                case 3: fa_pos_resolve<3><<<blocks, 256, 0, st>>>(g, rex); break;
                default: fa_pos_resolve<4><<<blocks, 256, 0, st>>>(g, rex); break;
            }
        }
        DSRL_LAUNCH_CHECK();
        return DSRL_OK;
    };
    auto finish = [&](const PosArgs &a) -> int {
        int rc2;
        if (g.exact && !g.ab && (rc2 = resolve())) return rc2;
        if ((rc2 = opt_in_smem(fa_pos_finish, finish_smem))) return rc2;
        fa_pos_finish<<<dim3(g.Npad / 32, B), 256, finish_smem, st>>>(g, a.opart, g.half ? nullptr : Fcm, FcmH, nrm, a.grad_scale, a.dP, a.dx[0], a.dx[1], a.direct ? a.go : nullptr);
        DSRL_LAUNCH_CHECK();
        return DSRL_OK;
    };

    CUtensorMap tm_pm, tm_cm;
    if ((rc = make_map(&tm_pm, Fpm, (uint64_t)(1 + g.split) * B * g.Npad, (uint64_t)g.Kc))) return rc;
    if ((rc = make_map(&tm_cm, Fcm, (uint64_t)B * g.Kc + kTile, (uint64_t)g.Npad))) return rc;

    unsigned *ticket = next_ticket_slot(st);
    if (!ticket) return DSRL_ERR_CUDA;
    PosArgs a;
    a.Fpm = Fpm; a.FpmH = FpmH; a.nrm = nrm;
    a.dP = reinterpret_cast<float *>(saved + so.dP);
    a.opart = reinterpret_cast<float *>(ws + wo.opart);
    a.partials = reinterpret_cast<double *>(ws + wo.partials);
    a.ticket = ticket;
    a.sum_out = reinterpret_cast<double *>(saved);
    a.loss_out = loss_out;
    const double Z = reduction == DSRL_REDUCE_MEAN ? (double)B * (double)g.N * (double)g.N : 1.0;
    a.loss_div = Z;
    a.grad_scale = (float)(2.0 / Z);
    a.direct = need_grad && go && dx1 && dx2 && k == 1;
    a.dx[0] = dx1; a.dx[1] = dx2; a.go = go;
    a.tau = rex.tau; a.fcnt = const_cast<unsigned *>(rex.fcnt); a.fent = const_cast<unsigned *>(rex.fent);
    if (fused_out) *fused_out = a.direct;
    const dim3 grid(need_grad ? g.tiles * g.jsplit : g.tiles, need_grad ? g.G : 1, B);
    a.sb = rex.sb;
    a.Ppm = pex.Ppm; a.inv64 = pex.inv64; a.tau_out = reinterpret_cast<float *>(ws + wo.tau);
    a.tau_r2 = tau_r * tau_r; a.tau_floor2 = kAccNoise * kAccNoise; a.tau_ksigma = ksigma;
    if (need_grad && g.ab) {
        // symmetric two-pass form (fa_position_ab.cuh): D tiles j >= i -> loss, near-tie lists, sign planes; [exact signs: the
        // resolve pass sets the bits of the near ties]; sign planes -> gradient contraction -> dX / dP
        CUtensorMap tm_k, tm_v;
        if ((rc = make_map(&tm_pm, FpmH, (uint64_t)B * g.Npad, (uint64_t)g.Kc, kTile, true))) return rc;
        if ((rc = make_map(&tm_k, FpmH, (uint64_t)B * g.Npad, (uint64_t)g.Kc, kTile / 2, true))) return rc;
        if ((rc = make_map(&tm_v, FcmH, (uint64_t)B * g.Kc + kTile, (uint64_t)g.Npad, g.gcnt[0] / 2, true))) return rc;
        if ((rc = opt_in_smem(fa_pos_dsign, g.a_smem_bytes))) return rc;
        fa_pos_dsign<<<dim3(2 * g.a_units, 1, B), kDsignThreads, g.a_smem_bytes, st>>>(tm_pm, tm_k, g, a);
        DSRL_LAUNCH_CHECK();
        if (g.exact && (rc = resolve())) return rc;
        if ((rc = opt_in_smem(fa_pos_grad, g.b_smem_bytes))) return rc;
        fa_pos_grad<<<dim3(g.tiles * g.jsplit * g.G, 1, B), kGradThreads, g.b_smem_bytes, st>>>(tm_v, tm_pm, g, a);
        DSRL_LAUNCH_CHECK();
        if (g.raw_o && (rc = finish(a))) return rc;
        return DSRL_OK;
    }
    if (need_grad && g.pair && (!g.half || g.half_pair)) {
        // both channel groups have the same width when there are two (C1p == C2p) or the V box height would differ per group
        const bool same = g.G == 1 || g.gcnt[0] == g.gcnt[1];
        if (same) {
            CUtensorMap tm_k, tm_v;
            if (g.half) {
                if ((rc = make_map(&tm_pm, FpmH, (uint64_t)B * g.Npad, (uint64_t)g.Kc, kTile, true))) return rc;
                if ((rc = make_map(&tm_k, FpmH, (uint64_t)B * g.Npad, (uint64_t)g.Kc, kTile / 2, true))) return rc;
                if ((rc = make_map(&tm_v, FcmH, (uint64_t)B * g.Kc + kTile, (uint64_t)g.Npad, g.gcnt[0] / 2, true))) return rc;
                if ((rc = opt_in_smem(fa_pos_tiles_pair<false, true, true>, g.half_smem_bytes))) return rc;
                fa_pos_tiles_pair<false, true, true><<<grid, kThreads, g.half_smem_bytes, st>>>(tm_pm, tm_k, tm_v, g, a);
            } else {
            if ((rc = make_map(&tm_k, Fpm, (uint64_t)(1 + g.split) * B * g.Npad, (uint64_t)g.Kc, kTile / 2))) return rc;
            if ((rc = make_map(&tm_v, Fcm, (uint64_t)B * g.Kc + kTile, (uint64_t)g.Npad, g.gcnt[0] / 2))) return rc;
#define LAUNCH_PAIR(SP, RS)                                                                               \
            do {                                                                                          \
                if ((rc = opt_in_smem(fa_pos_tiles_pair<SP, RS>, g.pair_smem_bytes))) return rc;          \
                fa_pos_tiles_pair<SP, RS><<<grid, kThreads, g.pair_smem_bytes, st>>>(tm_pm, tm_k, tm_v, g, a); \
            } while (0)
            if (g.split) { if (g.q_resident) LAUNCH_PAIR(true, true); else LAUNCH_PAIR(true, false); }
            else         { if (g.q_resident) LAUNCH_PAIR(false, true); else LAUNCH_PAIR(false, false); }
#undef LAUNCH_PAIR
            }
            DSRL_LAUNCH_CHECK();
            if (g.raw_o && (rc = finish(a))) return rc;
            return DSRL_OK;
        }
    }
#define LAUNCH_TILES(GR, SP, RS)                                                                          \
    do {                                                                                                  \
        if ((rc = opt_in_smem(fa_pos_tiles<GR, SP, RS>, g.smem_bytes))) return rc;                        \
        fa_pos_tiles<GR, SP, RS><<<grid, kThreads, g.smem_bytes, st>>>(tm_pm, tm_cm, g, a);               \
    } while (0)
    if (g.half) {                  // FP16 operands on the single-CTA kernel: forward only, odd tile counts, unequal channel groups
        if ((rc = make_map(&tm_pm, FpmH, (uint64_t)B * g.Npad, (uint64_t)g.Kc, kTile, true))) return rc;
        if ((rc = make_map(&tm_cm, FcmH, (uint64_t)B * g.Kc + kTile, (uint64_t)g.Npad, kTile, true))) return rc;
        if (need_grad) {
            if ((rc = opt_in_smem(fa_pos_tiles<true, false, true, true>, g.half1_smem_bytes))) return rc;
            fa_pos_tiles<true, false, true, true><<<grid, kThreads, g.half1_smem_bytes, st>>>(tm_pm, tm_cm, g, a);
        } else {
            if ((rc = opt_in_smem(fa_pos_tiles<false, false, true, true>, g.half1_smem_bytes))) return rc;
            fa_pos_tiles<false, false, true, true><<<grid, kThreads, g.half1_smem_bytes, st>>>(tm_pm, tm_cm, g, a);
        }
        DSRL_LAUNCH_CHECK();
        if (need_grad && g.raw_o && (rc = finish(a))) return rc;
        return DSRL_OK;
    }
    const int variant = (need_grad ? 4 : 0) | (g.split ? 2 : 0) | (g.q_resident ? 1 : 0);
    switch (variant) {
        case 0: LAUNCH_TILES(false, false, false); break;
        case 1: LAUNCH_TILES(false, false, true); break;
        case 2: LAUNCH_TILES(false, true, false); break;
        case 3: LAUNCH_TILES(false, true, true); break;
        case 4: LAUNCH_TILES(true, false, false); break;
        case 5: LAUNCH_TILES(true, false, true); break;
        case 6: LAUNCH_TILES(true, true, false); break;
        default: LAUNCH_TILES(true, true, true); break;
    }
#undef LAUNCH_TILES
    DSRL_LAUNCH_CHECK();
    if (need_grad && g.raw_o && (rc = finish(a))) return rc;
    return DSRL_OK;
}

int fa_pos_forward(int precision, const float *x1, const float *x2, int B, int C1, int C2, int H, int W, int k, int reduction,
                   int need_grad, float *loss_out, void *saved_v, size_t saved_bytes, void *ws_v, size_t ws_bytes, cudaStream_t st) {
    return fa_pos_forward_impl(precision, x1, x2, B, C1, C2, H, W, k, reduction, need_grad, loss_out, saved_v, saved_bytes, ws_v, ws_bytes, st,
                               nullptr, nullptr, nullptr, nullptr);
}

int fa_pos_forward_backward(int precision, const float *x1, const float *x2, int B, int C1, int C2, int H, int W, int k, int reduction,
                            const float *grad_out, float *loss_out, float *dx1, float *dx2, void *saved_v, size_t saved_bytes, void *ws_v,
                            size_t ws_bytes, cudaStream_t st) {
    int fused = 0;
    int rc = fa_pos_forward_impl(precision, x1, x2, B, C1, C2, H, W, k, reduction, 1, loss_out, saved_v, saved_bytes, ws_v, ws_bytes, st,
                                 grad_out, dx1, dx2, &fused);
    if (rc || fused) return rc;
    return fa_pos_backward(precision, x1, x2, saved_v, saved_bytes, grad_out, dx1, dx2, B, C1, C2, H, W, k, reduction, ws_v, ws_bytes, st);
}

int fa_pos_backward(int precision, const float *, const float *, const void *saved_v, size_t saved_bytes, const float *grad_out,
                    float *dx1, float *dx2, int B, int C1, int C2, int H, int W, int k, int, void *, size_t, cudaStream_t st) {
    PosGeom g;
    if (!geom_for(precision, B, C1, C2, H, W, k, g)) DSRL_FAIL(DSRL_ERR_BAD_SHAPE, "FA(position): unsupported geometry");
    const PosSaved so = make_saved(g);
    if (saved_bytes < so.total) DSRL_FAIL(DSRL_ERR_BAD_ARG, "FA(position): saved blob too small");
    const float *dP = reinterpret_cast<const float *>(static_cast<const unsigned char *>(saved_v) + so.dP);
    const bool v4 = (W % 4 == 0) && (!dx1 || (reinterpret_cast<uintptr_t>(dx1) & 15) == 0) &&
                    (!dx2 || (reinterpret_cast<uintptr_t>(dx2) & 15) == 0);
    const long long rows = (long long)B * (C1 + C2) * H;
    const dim3 block(64, 4);
    const long long blocks = (rows + block.y - 1) / block.y;
    if (blocks > 0x7fffffffLL) DSRL_FAIL(DSRL_ERR_UNSUPPORTED, "FA(position): gradient too large for one launch");
    if (v4) fa_pos_unpool<4><<<(unsigned)blocks, block, 0, st>>>(g, dP, grad_out, dx1, dx2);
    else fa_pos_unpool<1><<<(unsigned)blocks, block, 0, st>>>(g, dP, grad_out, dx1, dx2);
    DSRL_LAUNCH_CHECK();
    return DSRL_OK;
}

}  // namespace dsrl
