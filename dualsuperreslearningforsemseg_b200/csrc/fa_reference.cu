// FA loss, REFERENCE semantics (what models/losses/FALoss.py:8-34 of the reference computes), sm_100a.
//
//   P  = avgpool_k(X)                                   FALoss.py:23-24
//   S  = (P/sigma)^T (P/sigma),  sigma = ||P||_2        FALoss.py:8-11   (per (b,c): w x w, contraction over h)
//   L  = reduce_{i,j} | vec(S1)_i - vec(S2)_j |         FALoss.py:27-34  (ALL pairs, n = w*w values per side)
//   backward = closed form of the autograd graph
//
// This problem is latency / CUDA-core bound (K = h <= 32 at every shape the reference model produces), so it
// runs in FP32 FMA -- TF32 would already flip enough sign() terms to miss the gradient tolerance -- and the
// n^2 all-pairs tensor the reference materialises (FALoss.py:27-30) never exists.  Only
//   sum |a_i - b_j|   (loss)            and            c_i = sum_j sign(a_i - b_j)   (exact int32, gradient)
// are kept, computed either exactly in O(n log n) (sort one side, fp64 prefix sums, rank the other side: the fused
// training-shape kernel and maps with n >= 8192) or as an n^2 stream of one side through shared memory against
// register-resident values of the other (mid-size maps, where many small CTAs have the lower latency).
//
// Kernels (fa_ref_fused_small: the reference model's training shapes in ONE launch, see below; general path otherwise):
//   fa_ref_prepare   grid (B*C, 2 branches): pool -> one-sided Jacobi for sigma,u1,v1 -> S; zeroes the counters
//   fa_ref_pairs_sorted  n >= 8192: exact O(n log n) all-pairs (sort + fp64 prefix sums + rank searches), one CTA per
//                    (b, c, side, chunk of 16384 sorted values)
//   fa_ref_pairs     smaller n: grid (tiles, 2 passes, B*C): brute-force all-pairs, atomics only on int32 counters
//   fa_ref_grad      grid (B*C, 2 branches): G^ = A^(G+G^T), spectral-norm Jacobian, pooled gradient; loss finish
//   fa_ref_unpool    backward proper: dX = grad_out * dP / k^2 spread over the k x k windows (float4 stores)
//   fa_ref_none_fwd / fa_ref_none_pairs: reduction='none' (the (B,C,n^2) tensor is the API's output there)
#include <math.h>

#include "common.cuh"

namespace dsrl {
namespace {

struct RefGeom {
    int B, C, H, W, k, h, w, n, BC;
    int lda;       // padded row stride of the pooled matrix in shared memory (odd -> conflict-free both ways)
    int m, L, m_pad;  // Jacobi works on the m shorter-side vectors of length L
    int transposed;   // 1: vectors are columns of P (w < h)
};

struct RefSaved {  // byte offsets into the opaque `saved` blob
    size_t P, S, sigma, u, v, dA, cnt, gpart, gtick, total;
};

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// c += sign(x - y) with sign(0) = 0 and sign(NaN) = 0, in three SASS instructions (two FSET producing 0 / -1 and
// one IADD3): this is the inner loop of the all-pairs kernels, which are issue-bound.
__device__ __forceinline__ void sign_acc(int &c, float x, float y) {
    int gt, lt;
    asm("set.gt.s32.f32 %0, %1, %2;" : "=r"(gt) : "f"(x), "f"(y));
    asm("set.lt.s32.f32 %0, %1, %2;" : "=r"(lt) : "f"(x), "f"(y));
    c += lt - gt;   // gt, lt are 0 or -1
}

inline bool make_geom(int B, int C, int H, int W, int k, RefGeom &g) {
    if (B < 1 || C < 1 || H < 1 || W < 1 || k < 1) return false;
    g.B = B; g.C = C; g.H = H; g.W = W; g.k = k;
    g.h = H / k; g.w = W / k;
    if (g.h < 1 || g.w < 1) return false;
    if ((long long)g.w * g.w > (1LL << 24)) return false;
    g.n = g.w * g.w;
    g.BC = B * C;
    g.lda = g.w | 1;
    g.transposed = g.w < g.h;
    g.m = g.transposed ? g.w : g.h;
    g.L = g.transposed ? g.h : g.w;
    g.m_pad = (g.m + 1) & ~1;
    return true;
}

inline RefSaved make_saved(const RefGeom &g) {
    RefSaved s;
    size_t off = 16;  // [0,8) double local |.| sum, [8,16) reserved
    const size_t bc2 = 2 * (size_t)g.BC;
    s.P = off;      off = align_up(off + bc2 * g.h * g.w * 4, 16);
    s.S = off;      off = align_up(off + bc2 * g.n * 4, 16);
    s.sigma = off;  off = align_up(off + bc2 * 4, 16);
    s.u = off;      off = align_up(off + bc2 * g.h * 4, 16);
    s.v = off;      off = align_up(off + bc2 * g.w * 4, 16);
    s.dA = off;     off = align_up(off + bc2 * g.h * g.w * 4, 16);
    s.cnt = off;    off = align_up(off + bc2 * g.n * 4, 16);
    s.gpart = off;  off = align_up(off + bc2 * (size_t)((g.h + 7) / 8) * 8, 16);   // fa_ref_grad_rows: <G^, P> per 8-row chunk
    s.gtick = off;  off = align_up(off + bc2 * 4, 16);                            // ... and its last-CTA ticket per (b, c, branch)
    s.total = off;
    return s;
}

// all-pairs tiling: each CTA owns kPairsBlock*R values of one side and streams `tj` values of the other
constexpr int kPairsBlock = 256;
struct PairsPlan { int R, ti, tj, owner_tiles, ysplits; };
inline PairsPlan make_pairs_plan(const RefGeom &g) {
    PairsPlan p;
    p.R = g.n >= 16384 ? 4 : 1;
    p.ti = kPairsBlock * p.R;
    p.owner_tiles = (g.n + p.ti - 1) / p.ti;
    // aim for >= ~2 waves of CTAs without making the streamed range shorter than 256 values
    long long ctas = (long long)p.owner_tiles * 2 * g.BC;
    int ys = 1;
    while (ctas * ys < 2 * 148 && g.n / (ys * 2) >= 256) ys *= 2;
    if (g.n >= 16384) { while (g.n / ys > 4096) ys *= 2; }
    p.ysplits = ys;
    p.tj = (g.n + ys - 1) / ys;
    return p;
}

// ---------------------------------------------------------------------------------------------------------------
// prepare: pool, sigma/u1/v1, S
// ---------------------------------------------------------------------------------------------------------------
constexpr int kPowerMaxM = 32;  // short side up to which the squaring + power-iteration solver is used
constexpr int kMidMaxM = 128;   // ... and up to which it is TRIED first (one M buffer aliases the pooled matrix), Jacobi being the fallback
struct PrepSmem { size_t A, Wm, vec, scratch, total; int sepW; int mid; size_t M0; };
inline PrepSmem make_prep_smem(const RefGeom &g, size_t limit) {
    PrepSmem s;
    s.mid = 0; s.M0 = 0;
    const size_t a_bytes = align_up((size_t)g.h * g.lda * 4, 16);
    if (g.m > kPowerMaxM && g.m <= kMidMaxM) {
        // Large maps (e.g. 128 x 256 pooled cells): the one-sided Jacobi sweeps of a single CTA took 9.6 of the 10 ms of a
        // forward + backward at w = 256.  The squaring solver is tried first: region X holds the pooled matrix, then the second
        // matrix buffer of the squarings, then the pooled matrix again (reloaded from what this CTA wrote) for the power steps --
        // and, should they not converge (tiny spectral gap), the Jacobi work area; region Y holds the first matrix buffer.
        const size_t mm = align_up((size_t)g.m * g.m * 4, 16), jw = align_up((size_t)g.m_pad * g.L * 4, 16);
        const size_t X = a_bytes > jw ? (a_bytes > mm ? a_bytes : mm) : (jw > mm ? jw : mm);
        const size_t tail = align_up((size_t)(g.h + g.w + g.m_pad + 32) * 8, 16) + 40 * 8 + 16;
        if (X + mm + tail <= limit) {
            s.mid = 1; s.sepW = 0;
            s.A = 0; s.Wm = 0; s.M0 = X;
            s.vec = X + mm;
            s.scratch = s.vec + align_up((size_t)(g.h + g.w + g.m_pad + 32) * 8, 16);
            s.total = s.scratch + 40 * 8 + 16;
            return s;
        }
    }
    const size_t w_bytes = g.m <= kPowerMaxM ? align_up((size_t)2 * g.m * g.m * 4, 16) : align_up((size_t)g.m_pad * g.L * 4, 16);
    const size_t tail = align_up((size_t)(g.h + g.w + g.m_pad + 32) * 8, 16) + 40 * 8 + 16;
    s.sepW = g.m <= kPowerMaxM || (a_bytes + w_bytes + tail) <= limit;   // the power solver never aliases
    s.A = 0;
    s.Wm = s.sepW ? a_bytes : 0;
    s.vec = s.sepW ? a_bytes + w_bytes : (a_bytes > w_bytes ? a_bytes : w_bytes);
    s.scratch = s.vec + align_up((size_t)(g.h + g.w + g.m_pad + 32) * 8, 16);
    s.total = s.scratch + 40 * 8 + 16;
    return s;
}

__device__ __forceinline__ float pool_cell(const float *__restrict__ x, int W, int k, int py, int px, bool vec4) {
    const float *base = x + (size_t)py * k * W + (size_t)px * k;
    float s = 0.f;
    if (vec4) {
        for (int dy = 0; dy < k; ++dy) {
            const float4 *r = reinterpret_cast<const float4 *>(base + (size_t)dy * W);
            for (int q = 0; q < k / 4; ++q) { float4 v = __ldg(r + q); s += v.x; s += v.y; s += v.z; s += v.w; }
        }
    } else {
        for (int dy = 0; dy < k; ++dy)
            for (int dx = 0; dx < k; ++dx) s += __ldg(base + (size_t)dy * W + dx);
    }
    return s / (float)(k * k);
}

__device__ __noinline__ float pool_cell_rolled(const float *__restrict__ base, int W, int k, bool vec4, bool act, float a, float b);
// Large maps: the pooling of a (b, c, branch) read 8 MB through ONE SM when fa_ref_prepare did it (0.5 of its 0.76 ms at
// 1024 x 2048, k = 8); this grid-wide pass writes the pooled maps into the saved blob first and fa_ref_prepare loads them.
__global__ void __launch_bounds__(256) fa_ref_pool(const float *__restrict__ x1, const float *__restrict__ x2, RefGeom g, RefSaved so,
                                                   unsigned char *__restrict__ saved) {
    const long long hw = (long long)g.h * g.w, total = 2LL * g.BC * hw;
    float *P = reinterpret_cast<float *>(saved + so.P);
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long slot = idx / hw;
        const int cell = (int)(idx - slot * hw), py = cell / g.w, px = cell - py * g.w;
        const int br = slot >= g.BC;
        const float *x = (br ? x2 : x1) + (size_t)(slot - (long long)br * g.BC) * g.H * g.W;
        const bool vec4 = (g.k % 4 == 0) && (g.W % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
        P[idx] = pool_cell_rolled(x + (size_t)py * g.k * g.W + (size_t)px * g.k, g.W, g.k, vec4, false, 1.f, 0.f);
    }
}

// One-sided (Hestenes) Jacobi on the rows of Wm (m rows of length L, m_pad = even round-up; the extra row is a
// dummy that never pairs).  Round-robin ordering: m_pad/2 disjoint pairs per round, one warp per pair.
__device__ void jacobi_rows(float *Wm, int m, int m_pad, int L, int *flag) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int np = m_pad >> 1, mm = m_pad - 1;
    const float tol = 1e-6f;
    if (m < 2) return;
    for (int sweep = 0; sweep < 40; ++sweep) {
        __syncthreads();
        if (threadIdx.x == 0) *flag = 0;
        __syncthreads();
        for (int r = 0; r < mm; ++r) {
            for (int i = wid; i < np; i += nw) {
                const int p = (i == 0) ? mm : (r + i) % mm;
                const int q = (i == 0) ? r : (r + mm - i) % mm;
                if (p >= m || q >= m) continue;
                float *rp = Wm + (size_t)p * L, *rq = Wm + (size_t)q * L;
                float a = 0.f, b = 0.f, g = 0.f;
                for (int c = lane; c < L; c += 32) { const float x = rp[c], y = rq[c]; a = fmaf(x, x, a); b = fmaf(y, y, b); g = fmaf(x, y, g); }
                a = warp_sum(a); b = warp_sum(b); g = warp_sum(g);
                if (fabsf(g) > tol * sqrtf(a * b) && a > 0.f && b > 0.f) {
                    const float zeta = (b - a) / (2.f * g);
                    const float t = copysignf(1.f, zeta) / (fabsf(zeta) + sqrtf(1.f + zeta * zeta));
                    const float cs = rsqrtf(1.f + t * t), sn = cs * t;
                    for (int c = lane; c < L; c += 32) {
                        const float x = rp[c], y = rq[c];
                        rp[c] = cs * x - sn * y;
                        rq[c] = sn * x + cs * y;
                    }
                    if (lane == 0) *flag = 1;
                }
            }
            __syncthreads();
        }
        if (*flag == 0) break;
    }
    __syncthreads();
}


// A group of whole warps inside a CTA that synchronises on its own named barrier (bar 0 + all threads == __syncthreads).
struct Grp {
    int tid, nt, bar;
    __device__ __forceinline__ void sync() const { asm volatile("bar.sync %0, %1;" ::"r"(bar), "r"(nt) : "memory"); }
};
template <typename T>
__device__ __forceinline__ T group_sum(T v, T *scratch /* >= 33 */, const Grp &g) {
    const int lane = g.tid & 31, wid = g.tid >> 5, nw = g.nt >> 5;
    v = warp_sum(v);
    g.sync();
    if (lane == 0) scratch[wid] = v;
    g.sync();
    if (wid == 0) {
        T t = lane < nw ? scratch[lane] : T(0);
        t = warp_sum(t);
        if (lane == 0) scratch[32] = t;
    }
    g.sync();
    return scratch[32];
}

// C[i*m + j] = scale * sum_{k < K} X[i*xr + k*xc] * Y[k*yr + j*yc]  (i, j < m) by a whole group: a warp owns 4 rows x 128 columns,
// lane l the columns j0 + l + 32 q (consecutive lanes read consecutive words of Y: no bank conflicts for row-major Y and none
// for a transposed Y with an odd row stride; the four X values are broadcasts).  One fmaf chain over k per entry -- the same
// rounding as the plain loop it replaces, 16 multiply-adds per 8 shared loads instead of 1 per 2.
__device__ void group_mm_tiled(const float *X, int xr, int xc, const float *Y, int yr, int yc, int m, int K, float scale,
                               float *C, const Grp &grp) {
    const int lane = grp.tid & 31, wid = grp.tid >> 5, nw = grp.nt >> 5;
    const int it = (m + 3) / 4, jb = (m + 127) / 128;
    for (int task = wid; task < it * jb; task += nw) {
        const int i0 = (task % it) * 4, j0 = (task / it) * 128 + lane;
        int xo[4], yo[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) { xo[q] = min(i0 + q, m - 1) * xr; yo[q] = min(j0 + 32 * q, m - 1) * yc; }
        float acc[4][4] = {};
#pragma unroll 2
        for (int k = 0; k < K; ++k) {
            float a[4], b[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) { a[q] = X[xo[q] + k * xc]; b[q] = Y[k * yr + yo[q]]; }
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[p][q] = fmaf(a[p], b[q], acc[p][q]);
        }
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (i0 + p < m && j0 + 32 * q < m) C[(i0 + p) * m + j0 + 32 * q] = acc[p][q] * scale;
    }
}

// Top singular triple of the m x L view V[i][c] = sA[i*rs + c*cs] (m <= kPowerMaxM is the short side).
//  1) M = V V^T (fp32);  2) p squarings with trace normalisation -- M^(2^p) is numerically rank one unless the
//  spectral gap is tiny;  3) fp64 power steps on V itself until sigma stalls (its error is second order in the
//  vector error).  On exit xs (len m) and xl (len L) are unit vectors with V^T xs = sigma xl.
// reload != nullptr: M1 aliases the pooled matrix -- it is rebuilt from `reload` (h x w, row-major, written by this CTA) into
// sAw (row stride lda) once the squarings are done.  *stalled (if given) tells whether the power steps ended because sigma
// stopped moving (converged) rather than at the iteration limit.
__device__ double top_singular_power(const float *sA, int rs, int cs, int m, int L, float *M0, float *M1, double *xs,
                                     double *xs2, double *xl, double *scratch, const Grp &grp,
                                     const float *reload = nullptr, float *sAw = nullptr, int h = 0, int w = 0, int lda = 0,
                                     int *stalled = nullptr) {
    const int tid = grp.tid, nt = grp.nt, lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
    const bool tiled = m > kPowerMaxM;                  // short sides 33..128: register tiles (the plain loops are bound by shared-memory loads)
    if (tiled) {
        group_mm_tiled(sA, rs, cs, sA, cs, rs, m, L, 1.f, M0, grp);
    } else {
        for (int o = tid; o < m * m; o += nt) {
            const int i = o / m, j = o - i * m;
            float s = 0.f;
            for (int c = 0; c < L; ++c) s = fmaf(sA[i * rs + c * cs], sA[j * rs + c * cs], s);
            M0[o] = s;
        }
    }
    float *cur = M0, *nxt = M1;
    const int p = m <= 16 ? 8 : 6;
    for (int sq = 0; sq < p; ++sq) {
        grp.sync();
        float tr = 0.f;
        for (int i = 0; i < m; ++i) tr += cur[i * m + i];
        if (!(tr > 0.f)) break;                         // zero (or NaN) matrix: uniform exit
        const float inv = 1.f / tr;
        if (tiled) {
            group_mm_tiled(cur, m, 1, cur, m, 1, m, m, inv * inv, nxt, grp);
        } else {
            for (int o = tid; o < m * m; o += nt) {
                const int i = o / m, j = o - i * m;
                float s = 0.f;
                for (int k = 0; k < m; ++k) s = fmaf(cur[i * m + k], cur[k * m + j], s);
                nxt[o] = s * inv * inv;
            }
        }
        float *t = cur; cur = nxt; nxt = t;
    }
    grp.sync();
    int best = 0;
    for (int i = 1; i < m; ++i) if (cur[i * m + i] > cur[best * m + best]) best = i;
    if (tid < m) xs[tid] = (double)cur[tid * m + best];
    if (reload != nullptr) {
        grp.sync();
        for (int cell = tid; cell < h * w; cell += nt) { const int py = cell / w; sAw[py * lda + (cell - py * w)] = reload[cell]; }
    }
    bool stall = false;
    double sig = 0.0, sig_prev = -1.0;
    for (int it = 0; it < 60; ++it) {
        grp.sync();
        double ns = 0.0;
        for (int i = 0; i < m; ++i) ns += xs[i] * xs[i];
        if (!(ns > 0.0)) { sig = 0.0; break; }
        ns = 1.0 / sqrt(ns);
        double part = 0.0;
        for (int c = tid; c < L; c += nt) {
            double s = 0.0;
            for (int i = 0; i < m; ++i) s += (double)sA[i * rs + c * cs] * xs[i];
            s *= ns;
            xl[c] = s;
            part += s * s;
        }
        const double nl2 = group_sum(part, scratch, grp);
        if (!(nl2 > 0.0)) { sig = 0.0; break; }
        const double nl = 1.0 / sqrt(nl2);
        for (int i = wid; i < m; i += nw) {
            double s = 0.0;
            for (int c = lane; c < L; c += 32) s += (double)sA[i * rs + c * cs] * xl[c];
            s = warp_sum(s);
            if (lane == 0) xs2[i] = s * nl;
        }
        grp.sync();
        double s2 = 0.0;
        for (int i = 0; i < m; ++i) s2 += xs2[i] * xs2[i];
        sig = sqrt(s2);
        double *t = xs; xs = xs2; xs2 = t;
        if (it >= 1 && fabs(sig - sig_prev) <= 1e-11 * sig) { stall = true; break; }
        sig_prev = sig;
    }
    if (stalled != nullptr) *stalled = stall || !(sig > 0.0);
    grp.sync();
    // final consistent pair: us = xs/|xs|, xl = V^T us, sigma = |xl|
    double ns = 0.0;
    for (int i = 0; i < m; ++i) ns += xs[i] * xs[i];
    ns = ns > 0.0 ? 1.0 / sqrt(ns) : 0.0;
    double part = 0.0;
    for (int c = tid; c < L; c += nt) {
        double s = 0.0;
        for (int i = 0; i < m; ++i) s += (double)sA[i * rs + c * cs] * xs[i];
        s *= ns;
        xl[c] = s;
        part += s * s;
    }
    const double sigma2 = group_sum(part, scratch, grp);
    const double sigma = sqrt(sigma2), inv = sigma > 0.0 ? 1.0 / sigma : 0.0;
    for (int c = tid; c < L; c += nt) xl[c] *= inv;
    grp.sync();
    if (tid < m) xs2[tid] = xs[tid] * ns;               // unit short vector lands in xs2 ...
    grp.sync();
    if (tid < m) xs[tid] = xs2[tid];                    // ... and in xs (whichever buffer the caller reads)
    grp.sync();
    return sigma;
}

// Warp-level variant for the fused small-shape kernel (m <= 16, L <= 32): the same algorithm run by ONE warp with
// the vectors in registers / a few shared doubles and only __syncwarp between steps.  M0 must already hold V V^T.
// On exit xs[0..m) / xl[0..L) (shared) are the unit vectors; returns sigma.  Call from one full warp.
// Code size matters here (the fused kernel runs once per launch on a cold instruction cache): loops are kept
// rolled and the fp64 rsqrt lives in one non-inlined helper.
__device__ __noinline__ double wsum_d(double v) {      // one copy of the 10-shuffle butterfly instead of one per call site
#pragma unroll 1
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __noinline__ double group_sum_d(double v, double *scratch, int gtid, int gnt, int bar) {
    const Grp g{gtid, gnt, bar};
    const int lane = gtid & 31, wid = gtid >> 5, nw = gnt >> 5;
    v = wsum_d(v);
    g.sync();
    if (lane == 0) scratch[wid] = v;
    g.sync();
    if (wid == 0) {
        double t = lane < nw ? scratch[lane] : 0.0;
        t = wsum_d(t);
        if (lane == 0) scratch[32] = t;
    }
    g.sync();
    return scratch[32];
}

#ifdef DSRL_FUSED_TIMING
__device__ long long g_fused_dbg[8];
#endif
__device__ __noinline__ float top_singular_warp(const float *sA, int rs, int cs, int m, int L, float *M0, float *M1,
                                                float *xs, float *xl, int squarings) {
#ifdef DSRL_FUSED_TIMING
    const long long dbg_t0 = clock64();
#endif
    // fp32 throughout: sigma only needs ~1e-6 relative (the loss tolerance is 1e-4, the reference itself is fp32) and
    // this single-warp dependent chain is the critical path of the whole forward -- fp64 shuffles/rsqrt tripled it.
    const int lane = threadIdx.x & 31;
    const int qi = 32 / m, qj = 32 - qi * m, i0 = lane / m, j0 = lane - i0 * m;   // (i, j) of entry o advance by (qi, qj) per 32
    float *cur = M0, *nxt = M1;
#pragma unroll 1
    for (int sq = 0; sq < squarings; ++sq) {
        float tr = lane < m ? cur[lane * m + lane] : 0.f;
        tr = warp_sum(tr);
        if (!(tr > 0.f)) break;
        const float inv = __frcp_rn(tr);
        int i = i0, j = j0;
#pragma unroll 1
        for (int o = lane; o < m * m; o += 32) {
            float s0 = 0.f, s1 = 0.f;
#pragma unroll 4
            for (int k = 0; k + 1 < m; k += 2) {
                s0 = fmaf(cur[i * m + k], cur[k * m + j], s0);
                s1 = fmaf(cur[i * m + k + 1], cur[(k + 1) * m + j], s1);
            }
            if (m & 1) s0 = fmaf(cur[i * m + m - 1], cur[(m - 1) * m + j], s0);
            nxt[o] = (s0 + s1) * inv * inv;
            i += qi; j += qj;
            if (j >= m) { j -= m; ++i; }
        }
        __syncwarp();
        float *t = cur; cur = nxt; nxt = t;
    }
    // start: column of the largest diagonal entry
    float dg = lane < m ? cur[lane * m + lane] : -1.f;
    int best = lane;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float od = __shfl_xor_sync(0xffffffffu, dg, o);
        const int ob = __shfl_xor_sync(0xffffffffu, best, o);
        if (od > dg || (od == dg && ob < best)) { dg = od; best = ob; }
    }
    float x = lane < m ? cur[lane * m + best] : 0.f;              // short vector: lane i holds xs[i]
    // Power steps on V itself.  The stopping rule is on the VECTOR (max component change <= 1e-6): the error of
    // sigma is second order in the vector error, so a stalled sigma says nothing about u1/v1, and the rank-one term
    // of the gradient needs them to ~1e-5 (inputs without a spectral gap, e.g. randn, otherwise miss the tolerance).
#pragma unroll 1
    for (int it = 0; it < 96; ++it) {
        float nx = warp_sum(x * x);
        x *= nx > 0.f ? rsqrtf(nx) : 0.f;
        if (lane < m) xs[lane] = x;
        __syncwarp();
        float y0 = 0.f, y1 = 0.f;                                   // long vector: lane c holds xl[c]
        if (lane < L) {
#pragma unroll 4
            for (int i = 0; i + 1 < m; i += 2) { y0 = fmaf(sA[i * rs + lane * cs], xs[i], y0); y1 = fmaf(sA[(i + 1) * rs + lane * cs], xs[i + 1], y1); }
            if (m & 1) y0 = fmaf(sA[(m - 1) * rs + lane * cs], xs[m - 1], y0);
        }
        float y = y0 + y1;
        const float ny = warp_sum(y * y);
        y *= ny > 0.f ? rsqrtf(ny) : 0.f;
        if (lane < L) xl[lane] = y;
        __syncwarp();
        float z0 = 0.f, z1 = 0.f;
        if (lane < m) {
#pragma unroll 4
            for (int c = 0; c + 1 < L; c += 2) { z0 = fmaf(sA[lane * rs + c * cs], xl[c], z0); z1 = fmaf(sA[lane * rs + (c + 1) * cs], xl[c + 1], z1); }
            if (L & 1) z0 = fmaf(sA[lane * rs + (L - 1) * cs], xl[L - 1], z0);
        }
        const float z = z0 + z1;
        const float s2 = warp_sum(z * z);                           // sigma^2 estimate
        float d = fabsf(z * (s2 > 0.f ? rsqrtf(s2) : 0.f) - x);     // change of the unit short vector
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) d = fmaxf(d, __shfl_xor_sync(0xffffffffu, d, o));
        x = z;
        __syncwarp();
#ifdef DSRL_FUSED_TIMING
        if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) { g_fused_dbg[1] = it + 1; g_fused_dbg[2 + (it < 4 ? it : 4)] = __float_as_int(d); }
#endif
        if (!(d > 1e-6f)) break;                                    // also leaves on NaN
    }
#ifdef DSRL_FUSED_TIMING
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) g_fused_dbg[0] = clock64() - dbg_t0;
#endif
    // consistent final pair: us = x/|x|, xl = V^T us, sigma = |xl|
    {
        const float nx = warp_sum(x * x);
        x *= nx > 0.f ? rsqrtf(nx) : 0.f;
    }
    if (lane < m) xs[lane] = x;
    __syncwarp();
    float y0 = 0.f, y1 = 0.f;
    if (lane < L) {
#pragma unroll 4
        for (int i = 0; i + 1 < m; i += 2) { y0 = fmaf(sA[i * rs + lane * cs], xs[i], y0); y1 = fmaf(sA[(i + 1) * rs + lane * cs], xs[i + 1], y1); }
        if (m & 1) y0 = fmaf(sA[(m - 1) * rs + lane * cs], xs[m - 1], y0);
    }
    const float y = y0 + y1, s2 = warp_sum(y * y);
    const float sigma = sqrtf(s2);
    if (lane < L) xl[lane] = s2 > 0.f ? y / sigma : 0.f;
    __syncwarp();
    return sigma;
}

__global__ void fa_ref_prepare(const float *__restrict__ x1, const float *__restrict__ x2, RefGeom g, RefSaved so,
                               unsigned char *__restrict__ saved, PrepSmem ps, int pooled, int gram_later) {
    extern __shared__ __align__(16) unsigned char smraw[];
    float *sA = reinterpret_cast<float *>(smraw + ps.A);
    float *sW = reinterpret_cast<float *>(smraw + ps.Wm);
    double *du = reinterpret_cast<double *>(smraw + ps.vec);  // [h]
    double *dv = du + g.h;                                      // [w]
    double *nrm = dv + g.w;                                     // [m_pad]
    double *scratch = reinterpret_cast<double *>(smraw + ps.scratch);  // [34]
    int *flag = reinterpret_cast<int *>(scratch + 36);

    const int bc = blockIdx.x, br = blockIdx.y, tid = threadIdx.x, nt = blockDim.x;
    const int h = g.h, w = g.w, lda = g.lda;
    const float *x = (br == 0 ? x1 : x2) + (size_t)bc * g.H * g.W;
    const size_t slot = (size_t)br * g.BC + bc;
    float *gP = reinterpret_cast<float *>(saved + so.P) + slot * h * w;
    float *gS = reinterpret_cast<float *>(saved + so.S) + slot * g.n;
    int *gcnt = reinterpret_cast<int *>(saved + so.cnt) + slot * g.n;

    // 1. pool (FALoss.py:23-24)
    const bool small = g.m <= kPowerMaxM, mid = ps.mid != 0;
    const bool vec4 = (g.k % 4 == 0) && (g.W % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    for (int cell = tid; cell < h * w; cell += nt) {
        const int py = cell / w, px = cell - py * w;
        float v;
        if (pooled) v = gP[cell];                              // fa_ref_pool ran first
        else { v = pool_cell_rolled(x + (size_t)py * g.k * g.W + (size_t)px * g.k, g.W, g.k, vec4, false, 1.f, 0.f); gP[cell] = v; }   // all loads of a window in flight
        if (ps.sepW || mid) sA[py * lda + px] = v;
        if (!small && !mid) { if (g.transposed) sW[(size_t)px * g.L + py] = v; else sW[(size_t)py * g.L + px] = v; }
    }
    for (int i = tid; i < g.n; i += nt) gcnt[i] = 0;
    if (tid == 0) reinterpret_cast<unsigned *>(saved + so.gtick)[slot] = 0u;
    __syncthreads();

    // 2. top singular triple (sigma, u1, v1)
    double sigma;
    if (small) {
        float *M0 = sW, *M1 = sW + g.m * g.m;
        double *xs2a = nrm + g.m_pad;                               // [32] spare short vector
        // the solver ping-pongs two short buffers; both end up holding the unit short vector
        const Grp all{tid, nt, 0};
        if (!g.transposed) sigma = top_singular_power(sA, lda, 1, g.m, g.L, M0, M1, du, xs2a, dv, scratch, all);
        else               sigma = top_singular_power(sA, 1, lda, g.m, g.L, M0, M1, dv, xs2a, du, scratch, all);
    } else {
      int conv = 0;
      if (mid) {
        // squarings + power steps first (see make_prep_smem); the vectors end up in du / dv like in the small case
        float *M0 = reinterpret_cast<float *>(smraw + ps.M0), *M1 = sA;
        const Grp all{tid, nt, 0};
        if (!g.transposed) sigma = top_singular_power(sA, lda, 1, g.m, g.L, M0, M1, du, nrm, dv, scratch, all, gP, sA, h, w, lda, &conv);
        else               sigma = top_singular_power(sA, 1, lda, g.m, g.L, M0, M1, dv, nrm, du, scratch, all, gP, sA, h, w, lda, &conv);
        __syncthreads();
        if (!conv) {                                         // uniform: every thread computed the same sigma sequence
            for (int cell = tid; cell < h * w; cell += nt) {
                const int py = cell / w, px = cell - py * w;
                const float v = gP[cell];
                if (g.transposed) sW[(size_t)px * g.L + py] = v; else sW[(size_t)py * g.L + px] = v;
            }
            __syncthreads();
        }
      }
      if (!conv) {
        // one-sided Jacobi in fp32, then one power step in fp64 (error in sigma is second order)
        jacobi_rows(sW, g.m, g.m_pad, g.L, flag);
        {
            const int lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
            for (int r = wid; r < g.m; r += nw) {
                double s = 0.0;
                for (int c = lane; c < g.L; c += 32) { const double t = sW[(size_t)r * g.L + c]; s += t * t; }
                s = warp_sum(s);
                if (lane == 0) nrm[r] = s;
            }
            __syncthreads();
            if (tid == 0) {
                int best = 0;
                for (int r = 1; r < g.m; ++r) if (nrm[r] > nrm[best]) best = r;
                *flag = best;
            }
            __syncthreads();
        }
        const int top = *flag;
        const double top_inv = nrm[top] > 0.0 ? 1.0 / sqrt(nrm[top]) : 0.0;
        // start vector: v0 (length w) if rows were rows of P, u0 (length h) if they were columns
        double *start = g.transposed ? du : dv;
        for (int c = tid; c < g.L; c += nt) start[c] = (double)sW[(size_t)top * g.L + c] * top_inv;
        __syncthreads();
        if (!ps.sepW) {  // the Jacobi work area aliased the pooled matrix: reload it from what this CTA wrote
            for (int cell = tid; cell < h * w; cell += nt) { const int py = cell / w; sA[py * lda + (cell - py * w)] = gP[cell]; }
            __syncthreads();
        }
        auto mul_A = [&]() {   // du = A dv
            for (int y = tid; y < h; y += nt) { double s = 0.0; for (int c = 0; c < w; ++c) s += (double)sA[y * lda + c] * dv[c]; du[y] = s; }
            __syncthreads();
        };
        auto mul_At = [&]() {  // dv = A^T du
            for (int c = tid; c < w; c += nt) { double s = 0.0; for (int y = 0; y < h; ++y) s += (double)sA[y * lda + c] * du[y]; dv[c] = s; }
            __syncthreads();
        };
        auto normalise = [&](double *vec, int len) -> double {
            double s = 0.0;
            for (int i = tid; i < len; i += nt) s += vec[i] * vec[i];
            s = block_sum(s, scratch);
            const double nr = sqrt(s), inv = nr > 0.0 ? 1.0 / nr : 0.0;
            for (int i = tid; i < len; i += nt) vec[i] *= inv;
            __syncthreads();
            return nr;
        };
        if (!g.transposed) { mul_A(); normalise(du, h); mul_At(); sigma = normalise(dv, w); }
        else               { mul_At(); normalise(dv, w); mul_A(); sigma = normalise(du, h); }
      }
    }
    const float sigf = (float)sigma;
    float *gu = reinterpret_cast<float *>(saved + so.u) + slot * h;
    float *gv = reinterpret_cast<float *>(saved + so.v) + slot * w;
    for (int i = tid; i < h; i += nt) gu[i] = (float)du[i];
    for (int i = tid; i < w; i += nt) gv[i] = (float)dv[i];
    if (tid == 0) reinterpret_cast<float *>(saved + so.sigma)[slot] = sigf;

    // 3. S = A^^T A^ (FALoss.py:10-11); sigma == 0 gives 0/0 = NaN exactly like the reference
    if (gram_later) return;                                    // wide maps: fa_ref_gram does it on the whole grid
    for (int cell = tid; cell < h * w; cell += nt) { const int py = cell / w; float *p = &sA[py * lda + (cell - py * w)]; *p = *p / sigf; }
    __syncthreads();
    if (w < 64) {
        for (int o = tid; o < g.n; o += nt) {
            const int i = o / w, j = o - i * w;
            float s = 0.f;
            for (int y = 0; y < h; ++y) s = fmaf(sA[y * lda + i], sA[y * lda + j], s);
            gS[o] = s;
        }
    } else {  // 4 x 4 register tiles
        const int tw = (w + 3) / 4;
        for (int o = tid; o < tw * tw; o += nt) {
            const int ti = (o / tw) * 4, tj = (o % tw) * 4;
            float acc[4][4] = {};
            for (int y = 0; y < h; ++y) {
                float a[4], b[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) { a[q] = (ti + q < w) ? sA[y * lda + ti + q] : 0.f; b[q] = (tj + q < w) ? sA[y * lda + tj + q] : 0.f; }
#pragma unroll
                for (int p = 0; p < 4; ++p)
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[p][q] = fmaf(a[p], b[q], acc[p][q]);
            }
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (ti + p < w && tj + q < w) gS[(size_t)(ti + p) * w + tj + q] = acc[p][q];
        }
    }
}

// S = A^^T A^ for wide maps as its own grid-wide pass (inside fa_ref_prepare one SM per (b, c, branch) spent 0.17 ms on it at
// w = 256): a CTA owns a 64 x 64 tile of S, a thread 4 rows x 4 columns (columns tx + 16 q: conflict-free), the contraction
// over the pooled rows runs in ascending order with one fmaf chain per entry like the in-kernel form.
constexpr int kGramTile = 64, kGramRows = 32;
__global__ void __launch_bounds__(256) fa_ref_gram(RefGeom g, RefSaved so, unsigned char *__restrict__ saved) {
    __shared__ __align__(16) float sI[kGramRows][kGramTile], sJ[kGramRows][kGramTile];
    const int slot = blockIdx.z, i0 = blockIdx.y * kGramTile, j0 = blockIdx.x * kGramTile, tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4, h = g.h, w = g.w;
    const float *gP = reinterpret_cast<const float *>(saved + so.P) + (size_t)slot * h * w;
    float *gS = reinterpret_cast<float *>(saved + so.S) + (size_t)slot * g.n;
    const float sigf = reinterpret_cast<const float *>(saved + so.sigma)[slot];
    float acc[4][4] = {};
    for (int y0 = 0; y0 < h; y0 += kGramRows) {
        __syncthreads();
        for (int o = tid; o < kGramRows * kGramTile; o += 256) {
            const int r = o >> 6, c = o & 63, y = y0 + r;
            sI[r][c] = (y < h && i0 + c < w) ? gP[(size_t)y * w + i0 + c] / sigf : 0.f;
            sJ[r][c] = (y < h && j0 + c < w) ? gP[(size_t)y * w + j0 + c] / sigf : 0.f;
        }
        __syncthreads();
        const int lim = min(kGramRows, h - y0);
        for (int r = 0; r < lim; ++r) {
            const float4 a4 = *reinterpret_cast<const float4 *>(&sI[r][ty * 4]);
            const float a[4] = {a4.x, a4.y, a4.z, a4.w};
            float b[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) b[q] = sJ[r][tx + 16 * q];
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[p][q] = fmaf(a[p], b[q], acc[p][q]);
        }
    }
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int i = i0 + ty * 4 + p, j = j0 + tx + 16 * q;
            if (i < w && j < w) gS[(size_t)i * w + j] = acc[p][q];
        }
}

// ---------------------------------------------------------------------------------------------------------------
// all pairs: c_x = sum_y sign(x - y)  (both directions), sum |x - y| (pass 0 only)
// ---------------------------------------------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(kPairsBlock) fa_ref_pairs(RefGeom g, RefSaved so, unsigned char *__restrict__ saved,
                                                            double *__restrict__ partials, PairsPlan plan) {
    __shared__ __align__(16) float sy[1024];
    __shared__ double scratch[34];
    const int bc = blockIdx.z, pass = blockIdx.y;
    const int ot = blockIdx.x / plan.ysplits, ys = blockIdx.x % plan.ysplits;
    const float *X = reinterpret_cast<const float *>(saved + so.S) + ((size_t)pass * g.BC + bc) * g.n;
    const float *Y = reinterpret_cast<const float *>(saved + so.S) + ((size_t)(1 - pass) * g.BC + bc) * g.n;
    int *cnt = reinterpret_cast<int *>(saved + so.cnt) + ((size_t)pass * g.BC + bc) * g.n;
    const int n = g.n, tid = threadIdx.x;

    float xv[R];
    int c[R];
    float acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int i = ot * plan.ti + r * kPairsBlock + tid;
        xv[r] = i < n ? X[i] : 0.f;
        c[r] = 0;
        acc[r] = 0.f;
    }
    const int y0 = ys * plan.tj, y1 = min(n, y0 + plan.tj);
    for (int base = y0; base < y1; base += 1024) {
        const int len = min(1024, y1 - base);
        __syncthreads();
        for (int i = tid; i < 1024; i += kPairsBlock) sy[i] = i < len ? Y[base + i] : 0.f;
        __syncthreads();
        const int len4 = len & ~3;
        for (int j = 0; j < len4; j += 4) {
            const float4 y = *reinterpret_cast<const float4 *>(&sy[j]);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const float x = xv[r];
                sign_acc(c[r], x, y.x);
                sign_acc(c[r], x, y.y);
                sign_acc(c[r], x, y.z);
                sign_acc(c[r], x, y.w);
                if (pass == 0) { acc[r] += fabsf(x - y.x); acc[r] += fabsf(x - y.y); acc[r] += fabsf(x - y.z); acc[r] += fabsf(x - y.w); }
            }
        }
        for (int j = len4; j < len; ++j) {
            const float y = sy[j];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                sign_acc(c[r], xv[r], y);
                if (pass == 0) acc[r] += fabsf(xv[r] - y);
            }
        }
    }
    double local = 0.0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int i = ot * plan.ti + r * kPairsBlock + tid;
        if (i < n) {
            if (plan.ysplits == 1) cnt[i] = c[r]; else if (c[r]) atomicAdd(&cnt[i], c[r]);
            local += (double)acc[r];
        }
    }
    if (pass == 0) {
        const double tot = block_sum(local, scratch);
        if (tid == 0) partials[(size_t)bc * gridDim.x + blockIdx.x] = tot;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// all pairs, exact in O(n log n): sort one side, prefix-sum it in fp64, rank the other side.  The sorted side is cut into
// chunks of <= 16384 values (one CTA's shared memory); ranks and sums add over chunks.  For every x, per chunk:
//     lt = #{y < x}, le = #{y <= x}, gt = n_chunk - le
//     sum_j sign(x - y_j) = lt - gt (exact integer)        sum_j |x - y_j| = x (lt - gt) - pre[lt] + (pre[n] - pre[le])
// A NaN anywhere in either operand (dead channel) makes the loss NaN and every sign 0, as torch's sign() does.
// (sigma = 0 turns the WHOLE map of a (b, c, branch) into NaN, so a chunk-local check sees it.)
// grid (direction, B*C, chunk): direction 0 ranks branch 1 against sorted branch 2 (and owns the loss), 1 the reverse.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kSortThreads = 1024, kSortMinN = 8192;
inline int next_pow2(int n) { int p = 1; while (p < n) p <<= 1; return p; }
// Sorted values per CTA.  Sorting gets cheaper with short chunks (n log^2 n per chunk, more CTAs), ranking dearer (every value of
// the other side is searched in every chunk): about eight chunks balance the two -- measured, batch 8: n = 16384: 184 / 112 /
// 77 us with chunks of 16384 / 8192 / 4096; n = 65536: 389 / 291 / 436 us.
inline int sort_chunk_cap(int n) { const int c = next_pow2((n + 7) / 8); return c < 4096 ? 4096 : (c > 16384 ? 16384 : c); }
inline int sort_chunks(int n) { const int cap = sort_chunk_cap(n); return (n + cap - 1) / cap; }
inline int sort_chunk_len(int n) { const int c = sort_chunks(n); return (n + c - 1) / c; }
inline size_t sorted_smem_bytes(int m) { return align_up((size_t)next_pow2(m) * 4, 16) + ((size_t)m + 1) * 8 + 40 * 8; }

__global__ void __launch_bounds__(kSortThreads) fa_ref_pairs_sorted(RefGeom g, RefSaved so, unsigned char *__restrict__ saved,
                                                                   double *__restrict__ partials, int chunk_len, int P) {
    extern __shared__ __align__(16) unsigned char smraw[];
    float *sy = reinterpret_cast<float *>(smraw);
    double *pre = reinterpret_cast<double *>(smraw + align_up((size_t)P * 4, 16));
    double *scratch = pre + chunk_len + 1;                        // [34] + totals
    const int dir = blockIdx.x, bc = blockIdx.y, ch = blockIdx.z, nx = g.n, tid = threadIdx.x;
    const int y0 = ch * chunk_len, n = min(chunk_len, nx - y0);  // this CTA's chunk of the sorted side
    const float *X = reinterpret_cast<const float *>(saved + so.S) + ((size_t)dir * g.BC + bc) * nx;
    const float *Y = reinterpret_cast<const float *>(saved + so.S) + ((size_t)(1 - dir) * g.BC + bc) * nx + y0;
    int *cnt = reinterpret_cast<int *>(saved + so.cnt) + ((size_t)dir * g.BC + bc) * nx;
    const bool single = gridDim.z == 1;

    int nan = 0;
    for (int i = tid; i < P; i += kSortThreads) {
        const float y = i < n ? Y[i] : INFINITY;
        sy[i] = y;
        nan |= (y != y);
    }
    for (int i = tid; i < nx; i += kSortThreads) { const float x = X[i]; nan |= (x != x); }
    const bool has_nan = __syncthreads_or(nan) != 0;              // also orders the stores above

    // bitonic sort of P values: pair t of stride j is (i, i | j), i = t with a zero inserted at bit log2(j)
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < (P >> 1); t += kSortThreads) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), q = i | j;
                const float a = sy[i], b = sy[q];
                if ((a > b) == ((i & k) == 0)) { sy[i] = b; sy[q] = a; }
            }
            __syncthreads();
        }
    }

    // exclusive fp64 prefix sums: each thread owns a contiguous run, runs are chained by a block scan of their totals
    const int run = (n + kSortThreads - 1) / kSortThreads, r0 = min(n, tid * run), r1 = min(n, r0 + run);
    double mine = 0.0;
    for (int i = r0; i < r1; ++i) mine += (double)sy[i];
    double incl = mine;                                           // inclusive scan over threads: warp, then across warps
    const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) scratch[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        double w = scratch[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double t = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += t;
        }
        scratch[lane] = w;                                        // inclusive totals of warps 0..lane
    }
    __syncthreads();
    double acc = (wid ? scratch[wid - 1] : 0.0) + incl - mine;    // sum of everything before this thread's run
    for (int i = r0; i < r1; ++i) { pre[i] = acc; acc += (double)sy[i]; }
    if (r1 == n && r0 < n) pre[n] = acc;
    if (n == 0 && tid == 0) pre[0] = 0.0;
    __syncthreads();

    double local = 0.0;
    for (int i = tid; i < nx; i += kSortThreads) {
        const float x = X[i];
        int lo = 0, hi = n;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (sy[mid] < x) lo = mid + 1; else hi = mid; }
        const int lt = lo;
        hi = n;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (sy[mid] <= x) lo = mid + 1; else hi = mid; }
        const int le = lo, gt = n - le;
        const int c = has_nan ? 0 : lt - gt;
        if (single) cnt[i] = c; else if (c) atomicAdd(&cnt[i], c);          // counters were zeroed by fa_ref_prepare
        if (dir == 0) local += (double)x * (double)(lt - gt) - pre[lt] + (pre[n] - pre[le]);
    }
    if (dir == 0) {
        __syncthreads();
        const double tot = block_sum(local, scratch);
        if (tid == 0) partials[(size_t)bc * gridDim.z + ch] = has_nan ? (double)NAN : tot;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// grad: pooled-resolution gradient for unit upstream gradient, and the loss finish
// ---------------------------------------------------------------------------------------------------------------
struct GradSmem { size_t A, Gs, uv, scratch, total; int gs_in_smem; };
inline GradSmem make_grad_smem(const RefGeom &g) {
    GradSmem s;
    s.A = 0;
    size_t off = align_up((size_t)g.h * g.lda * 4, 16);
    s.gs_in_smem = g.w <= 64;
    s.Gs = off;
    if (s.gs_in_smem) off += align_up((size_t)g.w * (g.w + 1) * 4, 16);
    s.uv = off;      off += align_up((size_t)(g.h + g.w) * 4, 16);
    s.scratch = off; off += 40 * 8;
    s.total = off;
    return s;
}

// g_scale: 1/Z for mean, 1 for sum (counts are integers); for reduction='none' gfl holds float g values already
// weighted by the upstream gradient and g_scale = 1.
__global__ void fa_ref_grad(RefGeom g, RefSaved so, unsigned char *__restrict__ saved, const float *__restrict__ gfl,
                            float g_scale, GradSmem gs, const double *__restrict__ partials, int num_partials,
                            double loss_div, float *__restrict__ loss_out, int do_grad) {
    extern __shared__ __align__(16) unsigned char smraw[];
    double *scratch = reinterpret_cast<double *>(smraw + gs.scratch);
    const int tid = threadIdx.x, nt = blockDim.x;

    if (blockIdx.x == 0 && blockIdx.y == 0 && loss_out != nullptr) {  // deterministic loss finish
        double s = 0.0;
        for (int i = tid; i < num_partials; i += nt) s += partials[i];
        s = block_sum(s, scratch);
        if (tid == 0) {
            *reinterpret_cast<double *>(saved) = s;
            *loss_out = (float)(s / loss_div);
        }
    }
    if (!do_grad) return;

    float *sA = reinterpret_cast<float *>(smraw + gs.A);
    float *sG = reinterpret_cast<float *>(smraw + gs.Gs);
    float *su = reinterpret_cast<float *>(smraw + gs.uv);
    float *sv = su + g.h;
    const int bc = blockIdx.x, br = blockIdx.y, h = g.h, w = g.w, lda = g.lda, ldg_ = w + 1;
    const size_t slot = (size_t)br * g.BC + bc;
    const float *gP = reinterpret_cast<const float *>(saved + so.P) + slot * h * w;
    const int *gcnt = reinterpret_cast<const int *>(saved + so.cnt) + slot * g.n;
    const float *gg = gfl ? gfl + slot * g.n : nullptr;
    float *gdA = reinterpret_cast<float *>(saved + so.dA) + slot * h * w;
    const float sigma = reinterpret_cast<const float *>(saved + so.sigma)[slot];

    for (int cell = tid; cell < h * w; cell += nt) { const int py = cell / w; sA[py * lda + (cell - py * w)] = gP[cell] / sigma; }
    for (int i = tid; i < h; i += nt) su[i] = reinterpret_cast<const float *>(saved + so.u)[slot * h + i];
    for (int i = tid; i < w; i += nt) sv[i] = reinterpret_cast<const float *>(saved + so.v)[slot * w + i];
    auto gval = [&](int idx) -> float { return gg ? gg[idx] : (float)gcnt[idx] * g_scale; };
    if (gs.gs_in_smem) {
        for (int o = tid; o < g.n; o += nt) { const int a = o / w, b = o - a * w; sG[a * ldg_ + b] = gval(a * w + b) + gval(b * w + a); }
    }
    __syncthreads();

    // G^ = A^ (G + G^T);  inner = <G^, P>
    double inner = 0.0;
    for (int cell = tid; cell < h * w; cell += nt) {
        const int y = cell / w, xo = cell - y * w;
        float s = 0.f;
        if (gs.gs_in_smem) {
            for (int xp = 0; xp < w; ++xp) s = fmaf(sA[y * lda + xp], sG[xp * ldg_ + xo], s);
        } else {
            for (int xp = 0; xp < w; ++xp) s = fmaf(sA[y * lda + xp], gval(xp * w + xo) + gval(xo * w + xp), s);
        }
        gdA[cell] = s;
        inner += (double)s * (double)gP[cell];
    }
    inner = block_sum(inner, scratch);
    const float coef = (float)(inner / ((double)sigma * (double)sigma));
    for (int cell = tid; cell < h * w; cell += nt) {   // each thread re-reads exactly what it wrote
        const int y = cell / w, xo = cell - y * w;
        gdA[cell] = gdA[cell] / sigma - coef * su[y] * sv[xo];
    }
}

// Wide maps (w > 64, where G + G^T no longer fits next to A^ in shared memory): the product above took 4.7 of the 6.3 ms of a
// forward + backward at 128 x 256 pooled cells because ONE CTA per (b, c, branch) did all h * w * w multiply-adds with both G
// operands read from global memory, one of them with stride w.  Here a CTA owns kGradRows rows of G^ (grid.z chunks), stages
// 32 x 256 tiles of G + G^T in shared memory (both parts read along rows of G) and keeps one accumulator per owned row and
// column; the summation order per cell is the one of fa_ref_grad (ascending x', one fmaf chain), so the two kernels agree bit
// for bit.  <G^, P> goes through per-chunk partials and a self-resetting ticket: the last CTA of a (b, c, branch) adds them in
// chunk order (deterministic) and applies the spectral-norm term to the whole map.
constexpr int kGradRows = 8, kGradCols = 256, kGradTile = 32;
inline size_t grad_rows_smem(const RefGeom &g) {
    return align_up((size_t)kGradRows * g.w * 4, 16) + align_up((size_t)kGradTile * (kGradCols + 1) * 4, 16) + 40 * 8 + 16;
}
__global__ void __launch_bounds__(kGradCols) fa_ref_grad_rows(RefGeom g, RefSaved so, unsigned char *__restrict__ saved,
                                                              const float *__restrict__ gfl, float g_scale,
                                                              const double *__restrict__ partials, int num_partials,
                                                              double loss_div, float *__restrict__ loss_out) {
    extern __shared__ __align__(16) unsigned char smraw[];
    const int tid = threadIdx.x, nt = kGradCols, h = g.h, w = g.w, ldt = kGradCols + 1;
    float *sA = reinterpret_cast<float *>(smraw);                                                  // [kGradRows][w]
    float *sT = reinterpret_cast<float *>(smraw + align_up((size_t)kGradRows * w * 4, 16));       // [kGradTile][ldt]
    double *scratch = reinterpret_cast<double *>(smraw + align_up((size_t)kGradRows * w * 4, 16) + align_up((size_t)kGradTile * ldt * 4, 16));
    int *s_last = reinterpret_cast<int *>(scratch + 40);

    if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && loss_out != nullptr) {  // deterministic loss finish
        double s = 0.0;
        for (int i = tid; i < num_partials; i += nt) s += partials[i];
        s = block_sum(s, scratch);
        if (tid == 0) {
            *reinterpret_cast<double *>(saved) = s;
            *loss_out = (float)(s / loss_div);
        }
        __syncthreads();
    }

    const int bc = blockIdx.x, br = blockIdx.y, y0 = blockIdx.z * kGradRows, nchunks = gridDim.z;
    const int rows = min(kGradRows, h - y0);
    const size_t slot = (size_t)br * g.BC + bc;
    const float *gP = reinterpret_cast<const float *>(saved + so.P) + slot * h * w;
    const int *gcnt = reinterpret_cast<const int *>(saved + so.cnt) + slot * g.n;
    const float *gg = gfl ? gfl + slot * g.n : nullptr;
    float *gdA = reinterpret_cast<float *>(saved + so.dA) + slot * h * w;
    const float sigma = reinterpret_cast<const float *>(saved + so.sigma)[slot];
    auto gval = [&](int idx) -> float { return gg ? gg[idx] : (float)gcnt[idx] * g_scale; };

    for (int o = tid; o < kGradRows * w; o += nt) {
        const int r = o / w;
        sA[o] = r < rows ? gP[(size_t)(y0 + r) * w + (o - r * w)] / sigma : 0.f;
    }

    double inner = 0.0;
    for (int cb = 0; cb < w; cb += kGradCols) {
        float acc[kGradRows];
#pragma unroll
        for (int r = 0; r < kGradRows; ++r) acc[r] = 0.f;
        const int xo = cb + tid;
        for (int xp0 = 0; xp0 < w; xp0 += kGradTile) {
            __syncthreads();
            // G part of the tile: rows x' of G, this thread's column
#pragma unroll 4
            for (int xl = 0; xl < kGradTile; ++xl) {
                const int xp = xp0 + xl;
                sT[xl * ldt + tid] = (xp < w && xo < w) ? gval(xp * w + xo) : 0.f;
            }
            __syncthreads();
            // G^T part: rows (cb + c) of G, 32 consecutive x' per warp
            {
                const int xl = tid & 31, xp = xp0 + xl;
                for (int c = tid >> 5; c < kGradCols; c += kGradCols / 32) {
                    if (xp < w && cb + c < w) sT[xl * ldt + c] += gval((cb + c) * w + xp);
                }
            }
            __syncthreads();
            const int lim = min(kGradTile, w - xp0);
            for (int xl = 0; xl < lim; ++xl) {
                const float t = sT[xl * ldt + tid];
#pragma unroll
                for (int r = 0; r < kGradRows; ++r) acc[r] = fmaf(sA[r * w + xp0 + xl], t, acc[r]);
            }
        }
        if (xo < w) {
#pragma unroll
            for (int r = 0; r < kGradRows; ++r) {
                if (r < rows) {
                    const size_t cell = (size_t)(y0 + r) * w + xo;
                    gdA[cell] = acc[r];
                    inner += (double)acc[r] * (double)gP[cell];
                }
            }
        }
    }
    inner = block_sum(inner, scratch);
    double *gpart = reinterpret_cast<double *>(saved + so.gpart) + slot * nchunks;
    if (tid == 0) {
        gpart[blockIdx.z] = inner;
        __threadfence();
        const unsigned old = atomicInc(reinterpret_cast<unsigned *>(saved + so.gtick) + slot, (unsigned)nchunks - 1u);
        *s_last = old == (unsigned)nchunks - 1u;
    }
    __syncthreads();
    if (!*s_last) return;
    __threadfence();
    double tot = 0.0;
    for (int i = 0; i < nchunks; ++i) tot += __ldcg(gpart + i);     // same order in every thread
    const float coef = (float)(tot / ((double)sigma * (double)sigma));
    const float *gu = reinterpret_cast<const float *>(saved + so.u) + slot * h;
    const float *gv = reinterpret_cast<const float *>(saved + so.v) + slot * w;
    for (int cell = tid; cell < h * w; cell += nt) {
        const int y = cell / w, x = cell - y * w;
        gdA[cell] = __ldcg(gdA + cell) / sigma - coef * gu[y] * gv[x];
    }
}

// ---------------------------------------------------------------------------------------------------------------
// backward proper: spread the pooled gradient over the k x k windows
// ---------------------------------------------------------------------------------------------------------------
// One CTA per output row (branch, b*c, y): 32-bit index arithmetic only -- the flat 64-bit decomposition this replaces made
// the kernel instruction bound (134 MB of gradients in 69 us = 1.9 TB/s at 1024 x 2048, batch 8).
template <int VEC>
__global__ void __launch_bounds__(256) fa_ref_unpool(RefGeom g, RefSaved so, const unsigned char *__restrict__ saved,
                                                     const float *__restrict__ grad_out, float *__restrict__ dx1, float *__restrict__ dx2) {
    const float go = grad_out ? __ldg(grad_out) : 1.f;
    const float scale = go / (float)(g.k * g.k);
    const int wv = g.W / VEC, rows_per_branch = g.BC * g.H;
    for (int row = blockIdx.x; row < 2 * rows_per_branch; row += gridDim.x) {
        const int br = row >= rows_per_branch, r = row - br * rows_per_branch;
        float *dx = br ? dx2 : dx1;
        if (!dx) continue;
        const int bc = r / g.H, y = r - bc * g.H, py = y / g.k;
        const float *dA = reinterpret_cast<const float *>(saved + so.dA) + ((size_t)br * g.BC + bc) * g.h * g.w + (size_t)py * g.w;
        float *dst = dx + ((size_t)bc * g.H + y) * g.W;
        const bool live = py < g.h;
        for (int xv = threadIdx.x; xv < wv; xv += 256) {
            float out[VEC];
#pragma unroll
            for (int q = 0; q < VEC; ++q) {
                const int px = (xv * VEC + q) / g.k;
                out[q] = (live && px < g.w) ? dA[px] * scale : 0.f;
            }
            if (VEC == 4) __stcs(reinterpret_cast<float4 *>(dst) + xv, make_float4(out[0], out[1], out[2], out[3]));
            else dst[xv] = out[0];
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// reduction = 'none'
// ---------------------------------------------------------------------------------------------------------------
__global__ void fa_ref_none_fwd(RefGeom g, RefSaved so, const unsigned char *__restrict__ saved, float *__restrict__ out) {
    const int bc = blockIdx.y;
    const float *a = reinterpret_cast<const float *>(saved + so.S) + (size_t)bc * g.n;
    const float *b = reinterpret_cast<const float *>(saved + so.S) + ((size_t)g.BC + bc) * g.n;
    float *o = out + (size_t)bc * g.n * g.n;
    const long long nn = (long long)g.n * g.n;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < nn; idx += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(idx / g.n), j = (int)(idx - (long long)i * g.n);
        o[idx] = fabsf(__ldg(a + i) - __ldg(b + j));      // FALoss.py:27-30 index order: i*n + j
    }
}

// g1[i] = sum_j go[i,j] sign(a_i - b_j)   (role 0: one warp per row)
// g2[j] = -sum_i go[i,j] sign(a_i - b_j)  (role 1: one thread per column)
__global__ void fa_ref_none_pairs(RefGeom g, RefSaved so, const unsigned char *__restrict__ saved,
                                  const float *__restrict__ grad_out, float *__restrict__ gfl) {
    const int bc = blockIdx.z, role = blockIdx.y, n = g.n;
    const float *a = reinterpret_cast<const float *>(saved + so.S) + (size_t)bc * n;
    const float *b = reinterpret_cast<const float *>(saved + so.S) + ((size_t)g.BC + bc) * n;
    const float *go = grad_out + (size_t)bc * n * n;
    if (role == 0) {
        const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
        for (int i = blockIdx.x * wpb + (threadIdx.x >> 5); i < n; i += gridDim.x * wpb) {
            const float ai = a[i];
            float s = 0.f;
            for (int j = lane; j < n; j += 32) { const float bj = b[j]; s += go[(size_t)i * n + j] * (float)((ai > bj) - (ai < bj)); }
            s = warp_sum(s);
            if (lane == 0) gfl[(size_t)bc * n + i] = s;
        }
    } else {
        for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
            const float bj = b[j];
            float s = 0.f;
            for (int i = 0; i < n; ++i) { const float ai = a[i]; s += go[(size_t)i * n + j] * (float)((ai > bj) - (ai < bj)); }
            gfl[((size_t)g.BC + bc) * n + j] = -s;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// fused small-shape forward: everything for one (b, c) in ONE CTA (the reference model's training shapes:
// pooled map h <= 32, w <= 16, n = w*w <= 256).  Two 256-thread groups handle the two branches concurrently on
// their own named barriers (pool -> sigma,u1,v1 -> S), sort their own S, then group g ranks its values against the
// other group's sorted S (all pairs in O(n log n)), then each group produces its branch's pooled gradient.  The loss is finished by the last CTA to arrive (self-resetting
// atomicInc ticket), so forward is a single launch with no memset and a deterministic summation order.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kFusedMaxW = 16, kFusedMaxH = 32, kFusedLda = kFusedMaxW + 1;
struct FusedBranch {
    float P[kFusedMaxH * kFusedLda];   // pooled map
    float A[kFusedMaxH * kFusedLda];   // P / sigma
    float M[2 * kFusedMaxW * kFusedMaxW];
    float S[kFusedMaxW * kFusedMaxW];
    int cnt[kFusedMaxW * kFusedMaxW];
    float Gs[kFusedMaxW * kFusedLda];
    float sorted[kFusedMaxW * kFusedMaxW];     // this branch's S values, ascending
    double pre[kFusedMaxW * kFusedMaxW + 1];   // their exclusive prefix sums
    float xs[32], xl[32];
    double scratch[34];
    float sf[kFusedMaxW * kFusedMaxW];         // the sorted values scaled to S, what the other branch's searches read
    double keep[3];                            // max |P| (the scale of the unscaled Gram entries), sigma, (max|P| / sigma)^2
    unsigned wmax[8];                          // per-warp maxima of |P| (bit patterns)
};
inline bool fused_ok(const RefGeom &g) { return g.w <= kFusedMaxW && g.h <= kFusedMaxH; }

// act: the map is a feature transformer's convolution output and the cell averages relu(a * z + b) -- BatchNorm2d(1) + ReLU
// (models/DSRL.py:86-95) folded into the pooling read
__device__ __noinline__ float pool_cell_rolled(const float *__restrict__ base, int W, int k, bool vec4, bool act, float a, float b) {
    float s = 0.f;
    auto f = [&](float v) { return act ? fmaxf(fmaf(a, v, b), 0.f) : v; };
    if (vec4 && k == 8) {
        // the model's subsample factor: all 16 vector loads of the 8 x 8 window in flight at once (the inputs are cold in
        // HBM; issued one row at a time the pooling phase was eight dependent memory latencies long)
        float4 v[16];
#pragma unroll
        for (int dy = 0; dy < 8; ++dy) {
            const float4 *r = reinterpret_cast<const float4 *>(base + (size_t)dy * W);
            v[2 * dy] = __ldg(r);
            v[2 * dy + 1] = __ldg(r + 1);
        }
#pragma unroll
        for (int q = 0; q < 16; ++q) { s += f(v[q].x); s += f(v[q].y); s += f(v[q].z); s += f(v[q].w); }
    } else if (vec4) {
#pragma unroll 1
        for (int dy = 0; dy < k; ++dy) {
            const float4 *r = reinterpret_cast<const float4 *>(base + (size_t)dy * W);
#pragma unroll 2
            for (int q = 0; q < (k >> 2); ++q) { const float4 v = __ldg(r + q); s += f(v.x); s += f(v.y); s += f(v.z); s += f(v.w); }
        }
    } else {
#pragma unroll 1
        for (int dy = 0; dy < k; ++dy)
#pragma unroll 1
            for (int dx = 0; dx < k; ++dx) s += f(__ldg(base + (size_t)dy * W + dx));
    }
    return s / (float)(k * k);
}

// Thread-block cluster plumbing of the fused kernel's loss reduction (B*C <= 8: the whole grid is one cluster).
__device__ __forceinline__ uint32_t ref_cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void ref_cluster_arrive() { asm volatile("barrier.cluster.arrive.release;" ::: "memory"); }
__device__ __forceinline__ void ref_cluster_wait() { asm volatile("barrier.cluster.wait.acquire;" ::: "memory"); }
__device__ __forceinline__ void st_cluster_f64(double *own_smem, uint32_t rank, double v) {      // store into CTA `rank`'s copy
    uint32_t a = (uint32_t)__cvta_generic_to_shared(own_smem), r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
    asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(r), "d"(v) : "memory");
}

constexpr int kFusedThreads = 576, kFusedBranchThreads = 288;   // per branch: 256 workers + one solver warp
__global__ void __launch_bounds__(kFusedThreads) fa_ref_fused_small(const float *__restrict__ x1, const float *__restrict__ x2, RefGeom g,
                                                          RefSaved so, unsigned char *__restrict__ saved,
                                                          double *__restrict__ partials, unsigned *__restrict__ ticket,
                                                          float g_scale, double loss_div, float *__restrict__ loss_out,
                                                          int need_grad, const float *__restrict__ grad_out,
                                                          float *__restrict__ dx1, float *__restrict__ dx2,
                                                          const float *__restrict__ bn, int use_cluster) {
    // Roles (round 2): S = P^T P / sigma^2, and sorting P^T P orders S, so nothing but the final scale needs sigma.  Per branch
    // the 256 workers pool, form the unscaled Gram entries and sort / prefix-sum them WHILE one solver warp runs M = V V^T, the
    // squarings and the power polish; both meet at one barrier.  The sigma solve (7.3 k cycles) and the Gram + sort (5.5 k) used
    // to run one after the other.
    __shared__ FusedBranch sb[2];
    __shared__ int is_last;
    __shared__ double cl_part[64];                      // cluster form: CTA 0 collects the per-warp loss partials of the (<= 8) CTAs here
    const int tid = threadIdx.x, br = tid / kFusedBranchThreads, ht = tid - br * kFusedBranchThreads;
    const bool solver = ht >= 256;                       // warp 8 of the branch
    const Grp grp{ht, 256, 1 + br};                      // the branch's workers
    const Grp all{ht, kFusedBranchThreads, 3 + br};      // workers + solver warp
    FusedBranch &fb = sb[br];
    const int bc = blockIdx.x, h = g.h, w = g.w, n = g.n, lda = kFusedLda, hw = h * w;
    // every per-thread index decomposition is done once: ht = qw*w + rw (cell / S entry), ht + 256 = second cell
    const int qw = ht / w, rw = ht - qw * w;
    const int qw2 = (ht + 256) / w, rw2 = (ht + 256) - qw2 * w;
    const bool c0 = !solver && ht < hw, c1 = !solver && ht + 256 < hw;
    const int a0 = qw * lda + rw, a1 = qw2 * lda + rw2;           // padded smem offsets of the thread's cells

#ifdef DSRL_FUSED_TIMING
    long long *tm = reinterpret_cast<long long *>(partials + 64);
#define TSTAMP(i) do { if (bc == 0 && tid == 0) tm[i] = clock64(); } while (0)
#else
#define TSTAMP(i)
#endif
    TSTAMP(0);
    // 1. pool (FALoss.py:23-24); the largest magnitude of the map rides along (one redux per warp): the Gram entries are formed
    //    from P / max|P|, so they neither overflow nor underflow whatever the scale of the input
    const float *x = (br ? x2 : x1) + (size_t)bc * g.H * g.W;
    const bool vec4 = (g.k % 4 == 0) && (g.W % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    const bool act = bn != nullptr;                          // {a, b, mean, invstd} per branch from ft_bn_forward_kernel
    const float bn_a = act ? __ldg(bn + 4 * br) : 1.f, bn_b = act ? __ldg(bn + 4 * br + 1) : 0.f;
    unsigned *wmax = fb.wmax;
    if (!solver) {
        float p0v = 0.f, p1v = 0.f;
        if (c0) { p0v = pool_cell_rolled(x + (size_t)qw * g.k * g.W + (size_t)rw * g.k, g.W, g.k, vec4, act, bn_a, bn_b); fb.P[a0] = p0v; }
        if (c1) { p1v = pool_cell_rolled(x + (size_t)qw2 * g.k * g.W + (size_t)rw2 * g.k, g.W, g.k, vec4, act, bn_a, bn_b); fb.P[a1] = p1v; }
        // |x| bit patterns order like the values; a NaN (exponent all ones, mantissa non-zero) wins the maximum and poisons the scale
        const unsigned mx = __reduce_max_sync(0xffffffffu, max(__float_as_uint(fabsf(p0v)), __float_as_uint(fabsf(p1v))));
        if ((ht & 31) == 0) wmax[ht >> 5] = mx;
    }
    all.sync();                                              // barrier 1: P (and the maxima) visible to workers and solver

    TSTAMP(1);
    const int rs = g.transposed ? 1 : lda, cs = g.transposed ? lda : 1, m = g.m, L = g.L;
    float sraw = 0.f;                                        // worker: its entry of (P/r)^T (P/r)
    float rmax = 0.f;
    double local = 0.0;
    if (solver) {
        // 2. sigma, u1, v1 (FALoss.py:10) by the solver warp alone: M = V V^T, six trace-normalised squarings (M^64 is numerically
        //    rank one unless the spectral gap is tiny), power steps on V itself until the vectors stand still
        const int lane = ht - 256;
        float sg;
        if (m <= 8) {
            // the model's shapes (pooled 8 x 16): M lives in registers, lane 8 i + j holds M[i][j] and M[i + 4][j] (rows / columns
            // past m are zero); a squaring is 24 shuffles + 16 multiply-adds instead of a chain of shared-memory round trips
            const int i = lane >> 3, j = lane & 7;
            float a0 = 0.f, a1 = 0.f;
#ifdef DSRL_FUSED_TIMING
#define SSTAMP(q) do { if (bc == 0 && br == 0 && lane == 0) g_fused_dbg[q] = clock64(); } while (0)
#else
#define SSTAMP(q)
#endif
            SSTAMP(0);
            {
                const bool v0 = i < m && j < m, v1 = i + 4 < m && j < m;
                const float *ri = fb.P + (v0 ? i : 0) * rs, *ri4 = fb.P + (v1 ? i + 4 : 0) * rs, *rj = fb.P + (j < m ? j : 0) * rs;
#pragma unroll 8
                for (int c = 0; c < L; ++c) {
                    const float vj = rj[c * cs];
                    a0 = fmaf(ri[c * cs], vj, a0);
                    a1 = fmaf(ri4[c * cs], vj, a1);
                }
                if (!v0) a0 = 0.f;
                if (!v1) a1 = 0.f;
            }
            SSTAMP(1);
            const float o0 = a0, o1 = a1;                       // M itself, for the power steps below
#pragma unroll 1
            for (int sq = 0; sq < 6; ++sq) {
                const float tr = warp_sum((i == j ? a0 : 0.f) + (i + 4 == j ? a1 : 0.f));
                if (!(tr > 0.f)) break;                          // zero / NaN matrix: same decision in every lane
                // the matrix squared last time had trace 1, so tr = sum lambda^2 / (sum lambda)^2, which is 1 exactly when it has
                // rank one: stop squaring as soon as that holds to rounding (three squarings for post-ReLU maps); the power
                // steps below check the vector anyway
                if (sq > 0 && tr > 1.f - 2e-6f) break;
                const float inv = __frcp_rn(tr);
                float n0 = 0.f, n1 = 0.f;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float mik0 = __shfl_sync(0xffffffffu, a0, (i << 3) | k), mik1 = __shfl_sync(0xffffffffu, a1, (i << 3) | k);
                    const float mkj = __shfl_sync(0xffffffffu, k < 4 ? a0 : a1, ((k & 3) << 3) | j);
                    n0 = fmaf(mik0, mkj, n0);
                    n1 = fmaf(mik1, mkj, n1);
                }
                a0 = n0 * inv * inv;
                a1 = n1 * inv * inv;
            }
            // Power steps with everything in registers: every lane holds the whole vector x and row (lane & 7) of M, so a step
            // is 8 multiply-adds, an all-gather of 8 shuffles and LOCAL norms / convergence test -- no warp reductions (the
            // shared-memory form with V-steps spent 2.5 k cycles per step on them).  Start: column of the largest diagonal
            // entry of M^64.  Same stopping rule as top_singular_warp: on the vector, because the gradient's rank-one term
            // needs u1 v1^T to ~1e-5.
            SSTAMP(2);
            float xv[8], row[8];
            {
                float dg[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) dg[k] = __shfl_sync(0xffffffffu, k < 4 ? a0 : a1, ((k & 3) << 3) | k);
                int best = 0;
#pragma unroll
                for (int k = 1; k < 8; ++k) if (k < m && dg[k] > dg[best]) best = k;
#pragma unroll
                for (int k = 0; k < 8; ++k) xv[k] = __shfl_sync(0xffffffffu, k < 4 ? a0 : a1, ((k & 3) << 3) | best);
                const int r = lane & 7;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float lo = __shfl_sync(0xffffffffu, o0, ((r & 3) << 3) | k), hi4 = __shfl_sync(0xffffffffu, o1, ((r & 3) << 3) | k);
                    row[k] = r < 4 ? lo : hi4;
                }
            }
            SSTAMP(3);
#pragma unroll 1
            for (int it = 0; it < 96; ++it) {
                float nx = 0.f;
#pragma unroll
                for (int k = 0; k < 8; ++k) nx = fmaf(xv[k], xv[k], nx);
                const float rn = nx > 0.f ? rsqrtf(nx) : 0.f;
                float z = 0.f;
#pragma unroll
                for (int k = 0; k < 8; ++k) { xv[k] *= rn; z = fmaf(row[k], xv[k], z); }
                float zk[8], s2 = 0.f;
#pragma unroll
                for (int k = 0; k < 8; ++k) { zk[k] = __shfl_sync(0xffffffffu, z, k); s2 = fmaf(zk[k], zk[k], s2); }
                const float rz = s2 > 0.f ? rsqrtf(s2) : 0.f;
                float d = 0.f;
#pragma unroll
                for (int k = 0; k < 8; ++k) { d = fmaxf(d, fabsf(zk[k] * rz - xv[k])); xv[k] = zk[k]; }
                if (!(d > 1e-6f)) break;                         // also leaves on NaN (fmaxf drops it: checked below)
                if (s2 != s2) break;
            }
            SSTAMP(4);
            {   // consistent final pair: us = x/|x|, xl = V^T us / sigma, sigma = |V^T us|
                float nx = 0.f;
#pragma unroll
                for (int k = 0; k < 8; ++k) nx = fmaf(xv[k], xv[k], nx);
                const float rn = nx > 0.f ? rsqrtf(nx) : (nx == 0.f ? 0.f : NAN);
#pragma unroll
                for (int k = 0; k < 8; ++k) if (lane == k && k < m) fb.xs[k] = xv[k] * rn;
                float y = 0.f;
                if (lane < L) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) if (k < m) y = fmaf(fb.P[k * rs + lane * cs], xv[k] * rn, y);
                }
                const float s2 = warp_sum(y * y);
                sg = sqrtf(s2);
                if (lane < L) fb.xl[lane] = s2 > 0.f ? y / sg : 0.f;
                __syncwarp();
            }
            SSTAMP(5);
        } else {
            for (int o = lane; o < m * m; o += 32) {
                const int i = o / m, j = o - i * m;
                float s = 0.f;
#pragma unroll 4
                for (int c = 0; c < L; ++c) s = fmaf(fb.P[i * rs + c * cs], fb.P[j * rs + c * cs], s);
                fb.M[o] = s;
            }
            __syncwarp();
            sg = top_singular_warp(fb.P, rs, cs, m, L, fb.M, fb.M + m * m, fb.xs, fb.xl, 6);
        }
        if (lane == 0) fb.keep[1] = (double)sg;
    } else {
        // 3. unscaled Gram entries (FALoss.py:10-11 without the 1/sigma^2), then
        // 4. all pairs (FALoss.py:27-34), exact in O(n log n) instead of n^2: the workers sort their own branch's values (bitonic:
        //    shuffles inside a warp, 6 shared-memory exchanges across warps) and prefix-sum them in fp64; after the barrier
        //    every value x of the OTHER branch is ranked with two binary searches:
        //        lt = #{y < x}, le = #{y <= x}:   sum_j sign(x - y_j) = lt - (n - le)        (exact integer)
        //                                         sum_j |x - y_j|     = x (lt - gt) - pre[lt] + (pre[n] - pre[le])
        //    A NaN anywhere (dead channel, sigma = 0) makes the loss NaN and every sign 0, as torch's sign() does.
#pragma unroll
        for (int i = 0; i < 8; ++i) rmax = fmaxf(rmax, __uint_as_float(wmax[i]));     // NaN patterns: fmaxf drops them ...
        {
            unsigned any = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) any = max(any, wmax[i]);
            if (any > 0x7f800000u) rmax = NAN;                                          // ... so they are put back here
        }
        const float rinv = rmax > 0.f ? 1.f / rmax : 0.f;                              // NaN / inf scale: rinv = NaN / 0 -> NaN entries below
        if (ht < n) {
            float sacc = 0.f;
#pragma unroll 4
            for (int y = 0; y < h; ++y) sacc = fmaf(fb.P[y * lda + qw] * rinv, fb.P[y * lda + rw] * rinv, sacc);
            sraw = rmax > 0.f ? sacc : (rmax == 0.f ? 0.f : NAN);
        }
        const int lane = ht & 31;
        float v = ht < n ? sraw : INFINITY;
        if (v != v) v = INFINITY;                                // NaNs do not take part in the ordering (the result is NaN anyway)
        int xbuf = 0;
#pragma unroll
        for (int k = 2; k <= 256; k <<= 1) {
#pragma unroll
            for (int j = k >> 1; j > 0; j >>= 1) {
                float o;
                if (j >= 32) {                                   // partner in another warp: through shared memory, two buffers
                    float *xb = xbuf ? fb.A : fb.sorted;         // in turn (A is written after barrier 2): the barrier of the
                    xbuf ^= 1;                                   // next exchange orders this one's reads before the buffer's reuse
                    xb[ht] = v;
                    grp.sync();
                    o = xb[ht ^ j];
                } else {
                    o = __shfl_xor_sync(0xffffffffu, v, j);
                }
                const bool keep_min = ((ht & k) == 0) == ((ht & j) == 0);
                v = keep_min ? fminf(v, o) : fmaxf(v, o);
            }
        }
        fb.sorted[ht] = v;                                       // (the last exchange went through A, which is rebuilt after barrier 2)
        // exclusive prefix sums of the sorted values (the +inf padding sits at the end and is never read)
        double p = ht < n ? (double)v : 0.0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double t = __shfl_up_sync(0xffffffffu, p, o);
            if (lane >= o) p += t;
        }
        if (lane == 31) fb.scratch[ht >> 5] = p;
        grp.sync();
        double base = 0.0;
        for (int wq = 0; wq < (ht >> 5); ++wq) base += fb.scratch[wq];
        fb.pre[ht + 1] = base + p;
        if (ht == 0) { fb.pre[0] = 0.0; fb.keep[0] = (double)rmax; }
    }
    all.sync();                                              // barrier 2: sigma, u1, v1 <-> sorted values, prefix sums, scale
    TSTAMP(2);
    TSTAMP(3);
    const double sigma = fb.keep[1];
    const float *du = g.transposed ? fb.xl : fb.xs, *dv = g.transposed ? fb.xs : fb.xl;
    const float sigf = (float)sigma;
    if (c0) fb.A[a0] = fb.P[a0] / sigf;                      // A^ = P / sigma for the gradient (sigma == 0 -> NaN like the reference)
    if (c1) fb.A[a1] = fb.P[a1] / sigf;
    // both branches sorted and solved; S = (r / sigma)^2 * (P/r)^T (P/r) is NaN throughout for a dead (or non-finite) map
    // values compared in double: float Gram entry * (r/sigma)^2, the same expression on both sides, so both directions of the
    // ranking see the same numbers; every worker scales one sorted value of its own branch for the other branch's searches
    const double ro = fb.keep[0] / sigma, sc = ro * ro;
    if (!solver && ht < n) fb.sf[ht] = (float)((double)fb.sorted[ht] * sc);
    if (ht == 0) fb.keep[2] = sc;
    const bool bad = !(sigma > 0.0 && sigma < (double)INFINITY) || (!solver && ht < n && !(fabsf(sraw) < INFINITY));
    const bool has_nan = __syncthreads_or(bad) != 0;
    TSTAMP(4);
    double gs_term = 0.0;                                    // cnt_i * S_i: its sum is <G, S> = <G^, P> / (2 sigma g_scale)
    if (!solver && ht < n && (br == 0 || need_grad)) {
        const FusedBranch &ob = sb[1 - br];
        const double osc = ob.keep[2];
        const double xd = (double)sraw * sc;
        const float xf = (float)xd;                          // the very float the other branch finds for this entry in fb.sf
#ifdef DSRL_FUSED_TIMING
        if (bc == 0 && tid == 0) g_fused_dbg[6] = clock64();
#endif
        // lt = #{y < x}, le = #{y <= x}: two branch-free searches over the n <= 256 sorted values, interleaved (fixed 9 probes
        // each).  Ordering is decided on the float values of S, like the reference's fp32 tensors (FP64 compares took 270
        // cycles per probe here); the sums below use the doubles, so a pair within one float ulp counts as a tie with an error
        // below that ulp.
        int lt = 0, le = 0;
#pragma unroll
        for (int step = 256; step > 0; step >>= 1) {           // counts up to 256 = binary digits 256 .. 1
            const int p1 = lt + step, p2 = le + step;
            const float y1 = ob.sf[min(p1, n) - 1], y2 = ob.sf[min(p2, n) - 1];
            if (p1 <= n && y1 < xf) lt = p1;
            if (p2 <= n && y2 <= xf) le = p2;
        }
        const int gt = n - le;
#ifdef DSRL_FUSED_TIMING
        if (bc == 0 && tid == 0) g_fused_dbg[7] = clock64() + (lt & 0);
#endif
        fb.cnt[ht] = has_nan ? 0 : lt - gt;
        gs_term = (double)(has_nan ? 0 : lt - gt) * xd;
        if (br == 0)
            local = has_nan ? (double)NAN : xd * (double)(lt - gt) - ob.pre[lt] * osc + (ob.pre[n] - ob.pre[le]) * osc;
    }
    TSTAMP(5);
    // Loss partial of this (b, c).  Grids of up to 8 CTAs are launched as ONE cluster: the partial goes into CTA 0's shared
    // memory and every thread arrives (once, without waiting) at the cluster barrier; CTA 0 waits on it only at the very end, so
    // the exchange overlaps the gradient phase.  The global ticket it replaces (fence + atomic round trip here, fence + L2 reads
    // in the last CTA) cost 2 k + 2 k cycles on the critical path.  Larger grids keep the ticket.
    if (solver) {
        if (use_cluster) ref_cluster_arrive();
        return;                                              // no block-wide barrier follows
    }
    // One reduction pass for both sums a branch needs: the loss terms (branch 0) and cnt_i * S_i (the gradient's <G^, P>, which
    // equals 2 sigma <G, S> because S is symmetric -- so the gradient phase needs no reduction of its own).  Warp sums only: the
    // eight per-warp loss partials go straight into CTA 0's shared memory (cluster form) and are added there in a fixed order.
    {
        const int lane = ht & 31, wid = ht >> 5;
        const double tg = wsum_d(gs_term);
        if (lane == 0) fb.scratch[wid] = tg;
        if (br == 0) {
            const double tl = wsum_d(local);
            if (lane == 0) {
                if (use_cluster) st_cluster_f64(&cl_part[ref_cluster_rank() * 8 + wid], 0u, tl);
                else fb.scratch[8 + wid] = tl;
            }
        }
    }
    grp.sync();                                              // also orders the cnt writes before the G + G^T build below
    double gs_sum = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) gs_sum += fb.scratch[i];
    if (use_cluster) {
        ref_cluster_arrive();
    } else if (br == 0 && ht == 0) {
        double tot = 0.0;
#pragma unroll
        for (int i = 0; i < 8; ++i) tot += fb.scratch[8 + i];
        partials[bc] = tot;
        __threadfence();
        const unsigned old = atomicInc(ticket, gridDim.x - 1);     // wraps to 0 after the last CTA: self-resetting
        is_last = (old == gridDim.x - 1);
    }

    TSTAMP(6);
    // 5. pooled gradient for unit upstream gradient
    if (need_grad) {
        if (ht < n) fb.Gs[qw * lda + rw] = (float)(fb.cnt[ht] + fb.cnt[rw * w + qw]) * g_scale;     // G + G^T
        grp.sync();
        float gh0 = 0.f, gh1 = 0.f;
        if (c0) {
#pragma unroll 4
            for (int xp = 0; xp < w; ++xp) gh0 = fmaf(fb.A[qw * lda + xp], fb.Gs[xp * lda + rw], gh0);
        }
        if (c1) {
#pragma unroll 4
            for (int xp = 0; xp < w; ++xp) gh1 = fmaf(fb.A[qw2 * lda + xp], fb.Gs[xp * lda + rw2], gh1);
        }
        const float coef = (float)(2.0 * (double)g_scale * gs_sum / sigma);      // <G^, P> / sigma^2
        float *gdA = reinterpret_cast<float *>(saved + so.dA) + ((size_t)br * g.BC + bc) * hw;
        const float ga0 = gh0 / sigf - coef * du[qw] * dv[rw];
        const float ga1 = c1 ? gh1 / sigf - coef * du[qw2] * dv[rw2] : 0.f;
        if (c0) gdA[ht] = ga0;
        if (c1) gdA[ht + 256] = ga1;
        // single-launch forward + backward (dsrl_fa_forward_backward): the upstream gradient is already known, so the
        // pooled gradient is spread over its k x k windows here (host guarantees H % k == 0, W % k == 0, k % 4 == 0 and
        // 16-byte aligned outputs) instead of by fa_ref_unpool
        float *dx = br ? dx2 : dx1;
        if (dx != nullptr) {
            const float sc = __ldg(grad_out) / (float)(g.k * g.k);
            float *base = dx + (size_t)bc * g.H * g.W;
#pragma unroll 1
            for (int cell = 0; cell < 2; ++cell) {
                if (cell ? c1 : c0) {
                    const float v = (cell ? ga1 : ga0) * sc;
                    const float4 v4 = make_float4(v, v, v, v);
                    float *p = base + (size_t)(cell ? qw2 : qw) * g.k * g.W + (size_t)(cell ? rw2 : rw) * g.k;
#pragma unroll 1
                    for (int dy = 0; dy < g.k; ++dy)
                        for (int q4 = 0; q4 < (g.k >> 2); ++q4) reinterpret_cast<float4 *>(p + (size_t)dy * g.W)[q4] = v4;
                }
            }
        }
    }

    TSTAMP(7);
    // 6. loss finish by the last CTA (fixed summation order -> deterministic): the ticket holder's own warp sums the partials -- it
    //    wrote `is_last` itself, so the last CTA, which is the critical path of the launch, needs no block-wide barrier
    if (use_cluster) {
        if (tid < 32 && ref_cluster_rank() == 0) {
            ref_cluster_wait();                              // every thread of every CTA has arrived: all partials are in cl_part
            if (tid == 0) {
                double s = 0.0;
                for (int i = 0; i < 8 * (int)gridDim.x; ++i) s += cl_part[i];
                *reinterpret_cast<double *>(saved) = s;
                *loss_out = (float)(s / loss_div);
            }
        }
    } else if (tid < 32) {
        __syncwarp();
        if (is_last) {
            __threadfence();
            double s = 0.0;
#pragma unroll 1
            for (int i = tid; i < (int)gridDim.x; i += 32) s += __ldcg(partials + i);
            s = wsum_d(s);
            if (tid == 0) {
                *reinterpret_cast<double *>(saved) = s;
                *loss_out = (float)(s / loss_div);
            }
        }
    }
    TSTAMP(8);
#ifdef DSRL_FUSED_TIMING
    if (bc == 0 && tid == 0) { for (int q = 0; q < 8; ++q) tm[9 + q] = g_fused_dbg[q]; }
#endif
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
constexpr size_t kSmemLimit = 220 * 1024;

template <typename K>
int opt_in_smem(K kern, size_t bytes) {
    if (bytes > 48 * 1024) DSRL_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return DSRL_OK;
}

// The fused kernel as one thread-block cluster when the grid has at most 8 CTAs (its loss reduction then runs over distributed
// shared memory, see the kernel); a cluster size the device refuses falls back to the plain launch with the global ticket.
int launch_fused(const float *x1, const float *x2, const RefGeom &g, const RefSaved &so, unsigned char *saved, double *partials,
                 unsigned *ticket, float g_scale, double loss_div, float *loss_out, int need_grad, const float *grad_out, float *dx1,
                 float *dx2, const float *bn, cudaStream_t st) {
    static std::atomic<int> cluster_ok[9];                   // per cluster size: 0 untried, 1 works, -1 refused
    if (g.BC <= 8 && cluster_ok[g.BC].load() >= 0) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(g.BC); cfg.blockDim = dim3(kFusedThreads); cfg.dynamicSmemBytes = 0; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = g.BC; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        const cudaError_t e = cudaLaunchKernelEx(&cfg, fa_ref_fused_small, x1, x2, g, so, saved, partials, ticket, g_scale, loss_div, loss_out,
                                                 need_grad, grad_out, dx1, dx2, bn, 1);
        if (e == cudaSuccess) { cluster_ok[g.BC].store(1); ::dsrl::count_launch(); return DSRL_OK; }
        (void)cudaGetLastError();
        cluster_ok[g.BC].store(-1);
    }
    fa_ref_fused_small<<<g.BC, kFusedThreads, 0, st>>>(x1, x2, g, so, saved, partials, ticket, g_scale, loss_div, loss_out, need_grad, grad_out,
                                                       dx1, dx2, bn, 0);
    DSRL_LAUNCH_CHECK();
    return DSRL_OK;
}

int launch_grad(const RefGeom &g, const RefSaved &so, unsigned char *saved, const float *gfl, float g_scale,
                const double *partials, int num_partials, double loss_div, float *loss_out, int do_grad, cudaStream_t st) {
    if (do_grad && g.w > 32) {                 // (w = 64: 18.4 us in the single-CTA kernel below, 15.5 us here)
        const size_t smem = grad_rows_smem(g);
        if (smem > kSmemLimit) DSRL_FAIL(DSRL_ERR_UNSUPPORTED, "FA(reference): pooled map %dx%d too wide for shared memory", g.h, g.w);
        int rc = opt_in_smem(fa_ref_grad_rows, smem);
        if (rc) return rc;
        fa_ref_grad_rows<<<dim3(g.BC, 2, (g.h + kGradRows - 1) / kGradRows), kGradCols, smem, st>>>(g, so, saved, gfl, g_scale, partials,
                                                                                              num_partials, loss_div, loss_out);
        DSRL_LAUNCH_CHECK();
        return DSRL_OK;
    }
    GradSmem gs = make_grad_smem(g);
    if (gs.total > kSmemLimit) DSRL_FAIL(DSRL_ERR_UNSUPPORTED, "FA(reference): pooled map %dx%d too large for shared memory", g.h, g.w);
    int rc = opt_in_smem(fa_ref_grad, gs.total);
    if (rc) return rc;
    const int threads = g.h * g.w >= 8192 ? 512 : 256;
    dim3 grid(do_grad ? g.BC : 1, do_grad ? 2 : 1);
    fa_ref_grad<<<grid, threads, gs.total, st>>>(g, so, saved, gfl, g_scale, gs, partials, num_partials, loss_div, loss_out, do_grad);
    DSRL_LAUNCH_CHECK();
    return DSRL_OK;
}

int launch_unpool(const RefGeom &g, const RefSaved &so, const unsigned char *saved, const float *grad_out, float *dx1,
                  float *dx2, cudaStream_t st) {
    const bool v4 = (g.W % 4 == 0) && (!dx1 || (reinterpret_cast<uintptr_t>(dx1) & 15) == 0) &&
                    (!dx2 || (reinterpret_cast<uintptr_t>(dx2) & 15) == 0);
    const int threads = 256;
    const int blocks = (int)std::min<long long>(2LL * g.BC * g.H, 148LL * 32);
    if (v4) fa_ref_unpool<4><<<blocks, threads, 0, st>>>(g, so, saved, grad_out, dx1, dx2);
    else fa_ref_unpool<1><<<blocks, threads, 0, st>>>(g, so, saved, grad_out, dx1, dx2);
    DSRL_LAUNCH_CHECK();
    return DSRL_OK;
}

}  // namespace

// entry points used by fa_api.cu --------------------------------------------------------------------------------
size_t fa_ref_saved_bytes(int B, int C, int H, int W, int k) {
    RefGeom g;
    if (!make_geom(B, C, H, W, k, g)) return 0;
    return make_saved(g).total;
}

size_t fa_ref_workspace_bytes(int B, int C, int H, int W, int k) {
    RefGeom g;
    if (!make_geom(B, C, H, W, k, g)) return 0;
    PairsPlan p = make_pairs_plan(g);
    const size_t brute = (size_t)g.BC * p.owner_tiles * p.ysplits, sorted = (size_t)g.BC * sort_chunks(g.n);
    const size_t partials = (brute > sorted ? brute : sorted) * sizeof(double);
    const size_t gfl = 2 * (size_t)g.BC * g.n * sizeof(float);  // reduction='none' backward
    return align_up(partials > gfl ? partials : gfl, 256) + 256;
}

int fa_ref_forward(const float *x1, const float *x2, int B, int C, int H, int W, int k, int reduction, int need_grad,
                   float *loss_out, void *saved_v, size_t saved_bytes, void *ws, size_t ws_bytes, cudaStream_t st) {
    RefGeom g;
    if (!make_geom(B, C, H, W, k, g)) DSRL_FAIL(DSRL_ERR_BAD_SHAPE, "FA(reference): bad geometry B=%d C=%d H=%d W=%d k=%d", B, C, H, W, k);
    RefSaved so = make_saved(g);
    if (saved_bytes < so.total) DSRL_FAIL(DSRL_ERR_BAD_ARG, "FA(reference): saved blob too small (%zu < %zu)", saved_bytes, so.total);
    if (ws_bytes < fa_ref_workspace_bytes(B, C, H, W, k)) DSRL_FAIL(DSRL_ERR_BAD_ARG, "FA(reference): workspace too small");
    unsigned char *saved = static_cast<unsigned char *>(saved_v);

    if (reduction != DSRL_REDUCE_NONE && fused_ok(g)) {
        unsigned *ticket = next_ticket_slot(st);
        if (!ticket) return DSRL_ERR_CUDA;
        const double Z = reduction == DSRL_REDUCE_MEAN ? (double)g.BC * (double)g.n * (double)g.n : 1.0;
        return launch_fused(x1, x2, g, so, saved, static_cast<double *>(ws), ticket, (float)(1.0 / Z), Z, loss_out, need_grad, nullptr,
                            nullptr, nullptr, nullptr, st);
    }

    PrepSmem ps = make_prep_smem(g, kSmemLimit);
    if (ps.total > kSmemLimit) DSRL_FAIL(DSRL_ERR_UNSUPPORTED, "FA(reference): pooled map %dx%d too large for shared memory", g.h, g.w);
    int rc = opt_in_smem(fa_ref_prepare, ps.total);
    if (rc) return rc;
    const int pthreads = g.h * g.w >= 8192 ? 512 : 256;
    const int pooled = (long long)g.h * g.w * g.k * g.k >= (1LL << 18);    // >= 1 MB of input per (b, c, branch): pool grid-wide
    if (pooled) {
        const long long cells = 2LL * g.BC * g.h * g.w;
        fa_ref_pool<<<(int)std::min<long long>((cells + 255) / 256, 148LL * 8), 256, 0, st>>>(x1, x2, g, so, saved);
        DSRL_LAUNCH_CHECK();
    }
    const int gram_later = g.w >= 128 && 2 * g.BC <= 65535;
    fa_ref_prepare<<<dim3(g.BC, 2), pthreads, ps.total, st>>>(x1, x2, g, so, saved, ps, pooled, gram_later);
    DSRL_LAUNCH_CHECK();
    if (gram_later) {
        const int t = (g.w + kGramTile - 1) / kGramTile;
        fa_ref_gram<<<dim3(t, t, 2 * g.BC), 256, 0, st>>>(g, so, saved);
        DSRL_LAUNCH_CHECK();
    }

    if (reduction == DSRL_REDUCE_NONE) {
        const long long nn = (long long)g.n * g.n;
        const int blocks = (int)std::min<long long>((nn + 255) / 256, 148LL * 8);
        fa_ref_none_fwd<<<dim3(blocks, g.BC), 256, 0, st>>>(g, so, saved, loss_out);
        DSRL_LAUNCH_CHECK();
        return DSRL_OK;  // the gradient needs the upstream tensor: all of it happens in backward
    }

    double *partials = static_cast<double *>(ws);
    const double Zs = reduction == DSRL_REDUCE_MEAN ? (double)g.BC * (double)g.n * (double)g.n : 1.0;
    if (g.n >= kSortMinN) {      // exact O(n log n) all-pairs (smaller maps: the n^2 stream over many CTAs has less latency)
        const int nch = sort_chunks(g.n), len = sort_chunk_len(g.n);
        const size_t smem = sorted_smem_bytes(len);
        if ((rc = opt_in_smem(fa_ref_pairs_sorted, smem))) return rc;
        fa_ref_pairs_sorted<<<dim3(need_grad ? 2 : 1, g.BC, nch), kSortThreads, smem, st>>>(g, so, saved, partials, len, next_pow2(len));
        DSRL_LAUNCH_CHECK();
        return launch_grad(g, so, saved, nullptr, (float)(1.0 / Zs), partials, g.BC * nch, Zs, loss_out, need_grad, st);
    }
    PairsPlan plan = make_pairs_plan(g);
    dim3 grid(plan.owner_tiles * plan.ysplits, need_grad ? 2 : 1, g.BC);
    if (plan.R == 4) fa_ref_pairs<4><<<grid, kPairsBlock, 0, st>>>(g, so, saved, partials, plan);
    else fa_ref_pairs<1><<<grid, kPairsBlock, 0, st>>>(g, so, saved, partials, plan);
    DSRL_LAUNCH_CHECK();

    const double Z = reduction == DSRL_REDUCE_MEAN ? (double)g.BC * (double)g.n * (double)g.n : 1.0;
    return launch_grad(g, so, saved, nullptr, (float)(1.0 / Z), partials, g.BC * plan.owner_tiles * plan.ysplits, Z,
                       loss_out, need_grad, st);
}

// Forward + backward in ONE launch when the geometry allows it (training shapes, windows tile the map exactly);
// returns DSRL_ERR_UNSUPPORTED otherwise and the caller falls back to forward + backward.
// bn (optional): x1, x2 are the convolution outputs of the two feature transformers and the loss is taken on
// relu(a * x + b) with {a, b} = bn[4 * branch + {0, 1}] (device memory, written by dsrl_ft_bn_forward); dx then is the
// gradient w.r.t. the transformer OUTPUT (the BatchNorm / ReLU backward is dsrl_ft_bn_backward)
int fa_ref_forward_backward(const float *x1, const float *x2, int B, int C, int H, int W, int k, int reduction,
                            const float *grad_out, float *loss_out, float *dx1, float *dx2, void *saved_v, size_t saved_bytes,
                            void *ws, size_t ws_bytes, cudaStream_t st, const float *bn) {
    RefGeom g;
    if (!make_geom(B, C, H, W, k, g)) DSRL_FAIL(DSRL_ERR_BAD_SHAPE, "FA(reference): bad geometry B=%d C=%d H=%d W=%d k=%d", B, C, H, W, k);
    const bool aligned = (!dx1 || (reinterpret_cast<uintptr_t>(dx1) & 15) == 0) && (!dx2 || (reinterpret_cast<uintptr_t>(dx2) & 15) == 0);
    if (reduction == DSRL_REDUCE_NONE || !fused_ok(g) || H % k || W % k || k % 4 || !aligned) return DSRL_ERR_UNSUPPORTED;
    RefSaved so = make_saved(g);
    if (saved_bytes < so.total) DSRL_FAIL(DSRL_ERR_BAD_ARG, "FA(reference): saved blob too small (%zu < %zu)", saved_bytes, so.total);
    if (ws_bytes < fa_ref_workspace_bytes(B, C, H, W, k)) DSRL_FAIL(DSRL_ERR_BAD_ARG, "FA(reference): workspace too small");
    unsigned *ticket = next_ticket_slot(st);
    if (!ticket) return DSRL_ERR_CUDA;
    const double Z = reduction == DSRL_REDUCE_MEAN ? (double)g.BC * (double)g.n * (double)g.n : 1.0;
    return launch_fused(x1, x2, g, so, static_cast<unsigned char *>(saved_v), static_cast<double *>(ws), ticket, (float)(1.0 / Z), Z, loss_out,
                        1, grad_out, dx1, dx2, bn, st);
}

int fa_ref_backward(const void *saved_v, size_t saved_bytes, const float *grad_out, float *dx1, float *dx2, int B, int C,
                    int H, int W, int k, int reduction, void *ws, size_t ws_bytes, cudaStream_t st) {
    RefGeom g;
    if (!make_geom(B, C, H, W, k, g)) DSRL_FAIL(DSRL_ERR_BAD_SHAPE, "FA(reference): bad geometry");
    RefSaved so = make_saved(g);
    if (saved_bytes < so.total) DSRL_FAIL(DSRL_ERR_BAD_ARG, "FA(reference): saved blob too small");
    unsigned char *saved = const_cast<unsigned char *>(static_cast<const unsigned char *>(saved_v));
    if (reduction == DSRL_REDUCE_NONE) {
        if (ws_bytes < fa_ref_workspace_bytes(B, C, H, W, k)) DSRL_FAIL(DSRL_ERR_BAD_ARG, "FA(reference): workspace too small");
        float *gfl = static_cast<float *>(ws);
        const int bx = std::max(1, std::min(64, (g.n + 7) / 8));
        fa_ref_none_pairs<<<dim3(bx, 2, g.BC), 256, 0, st>>>(g, so, saved, grad_out, gfl);
        DSRL_LAUNCH_CHECK();
        int rc = launch_grad(g, so, saved, gfl, 1.f, nullptr, 0, 1.0, nullptr, 1, st);
        if (rc) return rc;
        return launch_unpool(g, so, saved, nullptr, dx1, dx2, st);
    }
    return launch_unpool(g, so, saved, grad_out, dx1, dx2, st);
}

}  // namespace dsrl
