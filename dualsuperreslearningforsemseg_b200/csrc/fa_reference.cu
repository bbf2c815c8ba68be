// FA loss, REFERENCE semantics (what models/losses/FALoss.py:8-34 of the reference computes), sm_100a.
//
//   P  = avgpool_k(X)                                   FALoss.py:23-24
//   S  = (P/sigma)^T (P/sigma),  sigma = ||P||_2        FALoss.py:8-11   (per (b,c): w x w, contraction over h)
//   L  = reduce_{i,j} | vec(S1)_i - vec(S2)_j |         FALoss.py:27-34  (ALL pairs, n = w*w values per side)
//   backward = closed form of the autograd graph
//
// This problem is latency / CUDA-core bound (K = h <= 32 at every shape the reference model produces), so it
// runs in FP32 FMA -- TF32 would already flip enough sign() terms to miss the gradient tolerance -- and the
// n^2 all-pairs tensor the reference materialises (FALoss.py:27-30) never exists: each CTA streams one side
// through shared memory against register-resident values of the other side and keeps only
//   sum |a_i - b_j|   (loss)            and            c_i = sum_j sign(a_i - b_j)   (exact int32, gradient).
//
// Kernels (general path, any n that fits):
//   fa_ref_prepare   grid (B*C, 2 branches): pool -> one-sided Jacobi for sigma,u1,v1 -> S; zeroes the counters
//   fa_ref_pairs     grid (tiles, 2 passes, B*C): brute-force all-pairs, atomics only on int32 counters
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
    size_t P, S, sigma, u, v, dA, cnt, total;
};

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

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
struct PrepSmem { size_t A, Wm, vec, scratch, total; int sepW; };
inline PrepSmem make_prep_smem(const RefGeom &g, size_t limit) {
    PrepSmem s;
    const size_t a_bytes = align_up((size_t)g.h * g.lda * 4, 16), w_bytes = align_up((size_t)g.m_pad * g.L * 4, 16);
    const size_t tail = align_up((size_t)(g.h + g.w + g.m_pad) * 8, 16) + 40 * 8 + 16;
    s.sepW = (a_bytes + w_bytes + tail) <= limit;
    s.A = 0;
    s.Wm = s.sepW ? a_bytes : 0;
    s.vec = s.sepW ? a_bytes + w_bytes : (a_bytes > w_bytes ? a_bytes : w_bytes);
    s.scratch = s.vec + align_up((size_t)(g.h + g.w + g.m_pad) * 8, 16);
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

__global__ void fa_ref_prepare(const float *__restrict__ x1, const float *__restrict__ x2, RefGeom g, RefSaved so,
                               unsigned char *__restrict__ saved, PrepSmem ps) {
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
    const bool vec4 = (g.k % 4 == 0) && (g.W % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    for (int cell = tid; cell < h * w; cell += nt) {
        const int py = cell / w, px = cell - py * w;
        const float v = pool_cell(x, g.W, g.k, py, px, vec4);
        gP[cell] = v;
        if (ps.sepW) sA[py * lda + px] = v;
        if (g.transposed) sW[(size_t)px * g.L + py] = v; else sW[(size_t)py * g.L + px] = v;
    }
    for (int i = tid; i < g.n; i += nt) gcnt[i] = 0;
    __syncthreads();

    // 2. top singular triple: Jacobi in fp32, then one power step in fp64 (error in sigma is second order)
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
    double sigma;
    if (!g.transposed) { mul_A(); normalise(du, h); mul_At(); sigma = normalise(dv, w); }
    else               { mul_At(); normalise(dv, w); mul_A(); sigma = normalise(du, h); }
    const float sigf = (float)sigma;
    float *gu = reinterpret_cast<float *>(saved + so.u) + slot * h;
    float *gv = reinterpret_cast<float *>(saved + so.v) + slot * w;
    for (int i = tid; i < h; i += nt) gu[i] = (float)du[i];
    for (int i = tid; i < w; i += nt) gv[i] = (float)dv[i];
    if (tid == 0) reinterpret_cast<float *>(saved + so.sigma)[slot] = sigf;

    // 3. S = A^^T A^ (FALoss.py:10-11); sigma == 0 gives 0/0 = NaN exactly like the reference
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
                c[r] += (x > y.x) - (x < y.x);
                c[r] += (x > y.y) - (x < y.y);
                c[r] += (x > y.z) - (x < y.z);
                c[r] += (x > y.w) - (x < y.w);
                if (pass == 0) { acc[r] += fabsf(x - y.x); acc[r] += fabsf(x - y.y); acc[r] += fabsf(x - y.z); acc[r] += fabsf(x - y.w); }
            }
        }
        for (int j = len4; j < len; ++j) {
            const float y = sy[j];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                c[r] += (xv[r] > y) - (xv[r] < y);
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

// ---------------------------------------------------------------------------------------------------------------
// backward proper: spread the pooled gradient over the k x k windows
// ---------------------------------------------------------------------------------------------------------------
template <int VEC>
__global__ void fa_ref_unpool(RefGeom g, RefSaved so, const unsigned char *__restrict__ saved,
                              const float *__restrict__ grad_out, float *__restrict__ dx1, float *__restrict__ dx2) {
    const float go = grad_out ? __ldg(grad_out) : 1.f;
    const float scale = go / (float)(g.k * g.k);
    const long long per_branch = (long long)g.BC * g.H * (g.W / VEC);
    const int wv = g.W / VEC;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < 2 * per_branch; idx += (long long)gridDim.x * blockDim.x) {
        const int br = idx >= per_branch;
        float *dx = br ? dx2 : dx1;
        if (!dx) continue;
        long long r = idx - (long long)br * per_branch;
        const int xv = (int)(r % wv); r /= wv;
        const int y = (int)(r % g.H);
        const int bc = (int)(r / g.H);
        const float *dA = reinterpret_cast<const float *>(saved + so.dA) + ((size_t)br * g.BC + bc) * g.h * g.w;
        const int py = y / g.k;
        float out[VEC];
#pragma unroll
        for (int q = 0; q < VEC; ++q) {
            const int xx = xv * VEC + q, px = xx / g.k;
            out[q] = (py < g.h && px < g.w) ? dA[py * g.w + px] * scale : 0.f;
        }
        float *dst = dx + ((size_t)bc * g.H + y) * g.W + (size_t)xv * VEC;
        if (VEC == 4) *reinterpret_cast<float4 *>(dst) = make_float4(out[0], out[1], out[2], out[3]);
        else dst[0] = out[0];
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
// host side
// ---------------------------------------------------------------------------------------------------------------
constexpr size_t kSmemLimit = 220 * 1024;

template <typename K>
int opt_in_smem(K kern, size_t bytes) {
    if (bytes > 48 * 1024) DSRL_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return DSRL_OK;
}

int launch_grad(const RefGeom &g, const RefSaved &so, unsigned char *saved, const float *gfl, float g_scale,
                const double *partials, int num_partials, double loss_div, float *loss_out, int do_grad, cudaStream_t st) {
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
    const long long total = 2LL * g.BC * g.H * (g.W / (v4 ? 4 : 1));
    const int threads = 256;
    const int blocks = (int)std::min<long long>((total + threads - 1) / threads, 148LL * 16);
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
    const size_t partials = (size_t)g.BC * p.owner_tiles * p.ysplits * sizeof(double);
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

    PrepSmem ps = make_prep_smem(g, kSmemLimit);
    if (ps.total > kSmemLimit) DSRL_FAIL(DSRL_ERR_UNSUPPORTED, "FA(reference): pooled map %dx%d too large for shared memory", g.h, g.w);
    int rc = opt_in_smem(fa_ref_prepare, ps.total);
    if (rc) return rc;
    const int pthreads = g.h * g.w >= 8192 ? 512 : 256;
    fa_ref_prepare<<<dim3(g.BC, 2), pthreads, ps.total, st>>>(x1, x2, g, so, saved, ps);
    DSRL_LAUNCH_CHECK();

    if (reduction == DSRL_REDUCE_NONE) {
        const long long nn = (long long)g.n * g.n;
        const int blocks = (int)std::min<long long>((nn + 255) / 256, 148LL * 8);
        fa_ref_none_fwd<<<dim3(blocks, g.BC), 256, 0, st>>>(g, so, saved, loss_out);
        DSRL_LAUNCH_CHECK();
        return DSRL_OK;  // the gradient needs the upstream tensor: all of it happens in backward
    }

    PairsPlan plan = make_pairs_plan(g);
    double *partials = static_cast<double *>(ws);
    dim3 grid(plan.owner_tiles * plan.ysplits, need_grad ? 2 : 1, g.BC);
    if (plan.R == 4) fa_ref_pairs<4><<<grid, kPairsBlock, 0, st>>>(g, so, saved, partials, plan);
    else fa_ref_pairs<1><<<grid, kPairsBlock, 0, st>>>(g, so, saved, partials, plan);
    DSRL_LAUNCH_CHECK();

    const double Z = reduction == DSRL_REDUCE_MEAN ? (double)g.BC * (double)g.n * (double)g.n : 1.0;
    return launch_grad(g, so, saved, nullptr, (float)(1.0 / Z), partials, g.BC * plan.owner_tiles * plan.ysplits, Z,
                       loss_out, need_grad, st);
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
