// Cross-entropy over NCHW logits with an ignore label -- the caller-side loss next to FA in the reference's training step
// (`t.nn.CrossEntropyLoss(ignore_index=IGNORE_CLASS_LABEL)` at command_handlers/train_or_resume.py:116, applied to the
// (B,19,512,1024) SSSR output at :435).  HBM bound.  PyTorch runs it as log_softmax -> nll_loss and two
// backward kernels, about seven passes over the logits-sized tensor; here:
//
//   ce_forward   ONE pass over the logits: online log-sum-exp per pixel, picks x[target], sums the loss and the number of
//                valid pixels (deterministic two-level reduction, last CTA by ticket), keeps (max, log2 sum) per pixel (8 B/px)
//   ce_backward  ONE pass: dlogits = (2^((x - max) log2 e - log2 sum) - [c == target]) * grad_out / valid   (0 at ignored pixels)
//
// Algorithmic bytes per pixel: forward 4C + sizeof(target) + 8, backward 4C + 4C + 8 + sizeof(target).
#include "common.cuh"

namespace dsrl {
namespace {

constexpr int kCeThreads = 256;

struct CeHeader {            // first bytes of the saved blob
    double sum;              // sum of the per-pixel losses over valid pixels
    long long valid;         // number of valid pixels
    long long bad;           // targets that are neither ignore_index nor a class index: the loss is NaN when there is one
};
struct CePartial { double sum; long long valid; };

constexpr size_t kCePartialsOff = 64;

inline size_t ce_blocks(int B, long long HW) { return (size_t)B * (size_t)((HW + kCeThreads - 1) / kCeThreads); }   // VEC = 1 worst case
inline size_t ce_lse_off(int B, long long HW) { return (kCePartialsOff + ce_blocks(B, HW) * sizeof(CePartial) + 255) / 256 * 256; }

constexpr float kLog2e = 1.4426950408889634f, kLn2 = 0.6931471805599453f;
// 2^x on the SFU, no range fix-up code around it (inputs are <= 0 here; -inf -> 0)
__device__ __forceinline__ float ex2_fast(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float lg2_fast(float x) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

template <typename T> __device__ __forceinline__ long long ce_load_target(const T *t, long long i) { return (long long)t[i]; }

// One thread = VEC consecutive pixels of one image; channel c of those pixels is one VEC*4-byte load, coalesced across the warp.
template <typename TT, int VEC, int kCeChunk = 5>
__global__ void __launch_bounds__(kCeThreads) ce_forward_kernel(const float *__restrict__ logits, const TT *__restrict__ target,
                                                                int nimg, int C, long long HW, long long ignore_index, int mean,
                                                                unsigned char *__restrict__ saved, size_t lse_off, unsigned *ticket,
                                                                float *__restrict__ loss_out) {
    __shared__ double s_sum[33];
    __shared__ long long s_cnt[33];
    __shared__ int s_last;
    // persistent CTAs: tile = (image, strip of kCeThreads * VEC pixels); one block-level reduction per CTA at the end
    const long long strips = (HW + (long long)kCeThreads * VEC - 1) / ((long long)kCeThreads * VEC), tiles = strips * nimg;
    float lsum = 0.f;
    int lcnt = 0, lbad = 0;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int b = (int)(tile / strips);
        const long long p0 = ((tile - (long long)b * strips) * kCeThreads + threadIdx.x) * VEC;
        if (p0 >= HW) continue;
        const float *x = logits + (size_t)b * C * HW + p0;
        // running maximum m and s = sum of 2^((x - m) log2 e): a subtract, a multiply, one SFU op and an add per logit.  x - m is
        // exact near the maximum (where the term matters), and x == m gives exactly 1, so a one-class problem has loss 0.
        float m[VEC], s[VEC], xt[VEC];
        int t[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            m[v] = -3.402823466e38f; s[v] = 0.f; xt[v] = 0.f;
            const long long tl = (p0 + v < HW) ? ce_load_target(target, (long long)b * HW + p0 + v) : ignore_index;
            t[v] = (tl != ignore_index && tl >= 0 && tl < C) ? (int)tl : -1;          // -1: ignored
            lbad += (tl != ignore_index && (tl < 0 || tl >= C)) ? 1 : 0;              // a label-mapping bug: torch asserts here
        }
        // kCeChunk channels per round: all their loads are issued before any arithmetic, then one branch-free online-softmax
        // update per pixel (chunk sizes 4..20 measured within 10 % of each other at the training shape; 5 was the fastest)
        const float *xp = x;                                // walks the channel planes: one 64-bit add per load, no index arithmetic
        for (int c0 = 0; c0 < C; c0 += kCeChunk) {
            float xv[kCeChunk][VEC];
            const bool whole = c0 + kCeChunk <= C;          // all but the last chunk: no per-channel guard
#pragma unroll
            for (int u = 0; u < kCeChunk; ++u) {
                if (whole || c0 + u < C) {
                    if (VEC == 4) {
                        const uint4 q = ldg_stream_u4(xp);
                        xv[u][0] = __uint_as_float(q.x); xv[u][1 % VEC] = __uint_as_float(q.y); xv[u][2 % VEC] = __uint_as_float(q.z); xv[u][3 % VEC] = __uint_as_float(q.w);
                    } else {
                        xv[u][0] = ldg_stream_f32(xp);
                    }
                    xp += HW;
                } else {
#pragma unroll
                    for (int v = 0; v < VEC; ++v) xv[u][v] = -3.402823466e38f;
                }
            }
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                float cm = xv[0][v];
#pragma unroll
                for (int u = 1; u < kCeChunk; ++u) cm = fmaxf(cm, xv[u][v]);
                const float mn = fmaxf(m[v], cm);
                float acc = s[v] * ex2_fast((m[v] - mn) * kLog2e);
#pragma unroll
                for (int u = 0; u < kCeChunk; ++u) {
                    acc += ex2_fast((xv[u][v] - mn) * kLog2e);         // padding channels: 2^-inf = 0
                    if (c0 + u == t[v]) xt[v] = xv[u][v];
                }
                s[v] = acc;
                m[v] = mn;
            }
        }
        // per pixel the pair (m, log2 s) is kept for the backward kernel rather than their sum: probabilities and the loss then
        // carry ~1e-7 relative error at any logit magnitude
        float l2s[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            l2s[v] = lg2_fast(s[v]);
            if (t[v] >= 0) { lsum += l2s[v] * kLn2 + (m[v] - xt[v]); ++lcnt; }
        }
        float2 *lp = reinterpret_cast<float2 *>(saved + lse_off) + (size_t)b * HW + p0;
        if (VEC == 4) {
            *reinterpret_cast<float4 *>(lp) = make_float4(m[0], l2s[0], m[1 % VEC], l2s[1 % VEC]);
            *reinterpret_cast<float4 *>(lp + 2) = make_float4(m[2 % VEC], l2s[2 % VEC], m[3 % VEC], l2s[3 % VEC]);
        } else {
            lp[0] = make_float2(m[0], l2s[0]);
        }
    }
    const double bsum = block_sum<double>((double)lsum, s_sum);
    // valid and bad counts travel in one word (pixel counts stay far below 2^40)
    const long long bcnt = block_sum<long long>((long long)lcnt + ((long long)lbad << 40), s_cnt);
    CePartial *parts = reinterpret_cast<CePartial *>(saved + kCePartialsOff);
    const unsigned nblk = gridDim.x, blk = blockIdx.x;
    if (threadIdx.x == 0) {
        parts[blk].sum = bsum;
        parts[blk].valid = bcnt;
        __threadfence();
        s_last = atomicInc(ticket, nblk - 1) == nblk - 1;          // self-resetting
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    double ts = 0.0;
    long long tc = 0;
    for (unsigned i = threadIdx.x; i < nblk; i += kCeThreads) { ts += __ldcg(&parts[i].sum); tc += __ldcg(&parts[i].valid); }   // fixed order
    ts = block_sum<double>(ts, s_sum);
    tc = block_sum<long long>(tc, s_cnt);
    if (threadIdx.x == 0) {
        CeHeader *h = reinterpret_cast<CeHeader *>(saved);
        const long long bad = tc >> 40;
        tc &= (1ll << 40) - 1;
        h->sum = ts;
        h->valid = tc;
        h->bad = bad;
        // no valid pixel: 0/0 = NaN, like torch.  A target outside [0, C) that is not ignore_index is a device assert in torch;
        // here the loss is NaN (the backward kernel still treats those pixels as ignored), which the reference's own NaN checks
        // (train_or_resume.py:406-411) turn into a stop
        *loss_out = bad ? __int_as_float(0x7fc00000) : (mean ? (float)(ts / (double)tc) : (float)ts);
    }
}

template <typename TT, int VEC>
__global__ void __launch_bounds__(kCeThreads) ce_backward_kernel(const float *__restrict__ logits, const TT *__restrict__ target,
                                                                 int C, long long HW, long long ignore_index, int mean,
                                                                 const unsigned char *__restrict__ saved, size_t lse_off,
                                                                 const float *__restrict__ grad_out, float *__restrict__ dlogits) {
    const int b = blockIdx.y;
    const long long p0 = ((long long)blockIdx.x * kCeThreads + threadIdx.x) * VEC;
    if (p0 >= HW) return;
    const CeHeader *h = reinterpret_cast<const CeHeader *>(saved);
    const float scale = mean ? (float)((double)__ldg(grad_out) / (double)h->valid) : __ldg(grad_out);
    const float *x = logits + (size_t)b * C * HW + p0;
    float *g = dlogits + (size_t)b * C * HW + p0;
    const float2 *lp = reinterpret_cast<const float2 *>(saved + lse_off) + (size_t)b * HW + p0;
    float m2[VEC], l2s[VEC], sc[VEC];                        // (m, log2 s) as the forward kernel left them
    int t[VEC];
    if (VEC == 4) {
        const float4 q0 = __ldg(reinterpret_cast<const float4 *>(lp)), q1 = __ldg(reinterpret_cast<const float4 *>(lp + 2));
        m2[0] = q0.x; l2s[0] = q0.y; m2[1 % VEC] = q0.z; l2s[1 % VEC] = q0.w;
        m2[2 % VEC] = q1.x; l2s[2 % VEC] = q1.y; m2[3 % VEC] = q1.z; l2s[3 % VEC] = q1.w;
    } else {
        const float2 q = __ldg(lp);
        m2[0] = q.x; l2s[0] = q.y;
    }
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        const long long tl = (p0 + v < HW) ? ce_load_target(target, (long long)b * HW + p0 + v) : ignore_index;
        const bool valid = tl != ignore_index && tl >= 0 && tl < C;
        t[v] = valid ? (int)tl : -1;
        sc[v] = valid ? scale : 0.f;
    }
    // 4 channels per round: a read + write stream wants resident threads more than registers (8 per round measured slower)
    constexpr int kCh = 4;
    const float *xp = x;                                    // walk the channel planes: one 64-bit add per access
    float *gp = g;
    for (int c0 = 0; c0 < C; c0 += kCh) {
        float xv[kCh][VEC];
        const bool whole = c0 + kCh <= C;
#pragma unroll
        for (int u = 0; u < kCh; ++u) {
            if (whole || c0 + u < C) {
                if (VEC == 4) {
                    const uint4 q = ldg_stream_u4(xp);
                    xv[u][0] = __uint_as_float(q.x); xv[u][1 % VEC] = __uint_as_float(q.y); xv[u][2 % VEC] = __uint_as_float(q.z); xv[u][3 % VEC] = __uint_as_float(q.w);
                } else {
                    xv[u][0] = ldg_stream_f32(xp);
                }
                xp += HW;
            }
        }
#pragma unroll
        for (int u = 0; u < kCh; ++u) {
            if (whole || c0 + u < C) {
                float o[VEC];
#pragma unroll
                for (int v = 0; v < VEC; ++v)
                    o[v] = t[v] >= 0 ? (ex2_fast((xv[u][v] - m2[v]) * kLog2e - l2s[v]) - (c0 + u == t[v] ? 1.f : 0.f)) * sc[v] : 0.f;
                if (VEC == 4) __stcs(reinterpret_cast<float4 *>(gp), make_float4(o[0], o[1 % VEC], o[2 % VEC], o[3 % VEC]));
                else gp[0] = o[0];
                gp += HW;
            }
        }
    }
}

int ce_check(const void *logits, const void *target, int B, int C, long long HW, int reduction) {
    if (!logits || !target) DSRL_FAIL(DSRL_ERR_BAD_ARG, "cross_entropy: null pointer");
    if (B < 0 || C < 1 || HW < 0 || B > 65535) DSRL_FAIL(DSRL_ERR_BAD_SHAPE, "cross_entropy: bad shape B=%d C=%d HW=%lld", B, C, HW);
    if (reduction != DSRL_REDUCE_MEAN && reduction != DSRL_REDUCE_SUM)
        DSRL_FAIL(DSRL_ERR_UNSUPPORTED, "cross_entropy: reduction must be mean or sum");
    return DSRL_OK;
}

}  // namespace
}  // namespace dsrl

using namespace dsrl;

extern "C" size_t dsrl_ce_saved_bytes(int B, int64_t HW) {
    if (B < 0 || HW < 0) return 0;
    return ce_lse_off(B, HW) + (size_t)B * (size_t)HW * 8 + 16;      // (m, log2 s) per pixel
}

#define CE_DISPATCH(KERNEL, ...)                                                                                     \
    do {                                                                                                             \
        switch (target_dtype) {                                                                                      \
            case DSRL_U8:  if (vec) KERNEL<uint8_t, 4><<<grid, kCeThreads, 0, st>>>(logits, static_cast<const uint8_t *>(target), __VA_ARGS__);      \
                           else KERNEL<uint8_t, 1><<<grid, kCeThreads, 0, st>>>(logits, static_cast<const uint8_t *>(target), __VA_ARGS__); break;   \
            case DSRL_I32: if (vec) KERNEL<int32_t, 4><<<grid, kCeThreads, 0, st>>>(logits, static_cast<const int32_t *>(target), __VA_ARGS__);      \
                           else KERNEL<int32_t, 1><<<grid, kCeThreads, 0, st>>>(logits, static_cast<const int32_t *>(target), __VA_ARGS__); break;   \
            case DSRL_I64: if (vec) KERNEL<long long, 4><<<grid, kCeThreads, 0, st>>>(logits, static_cast<const long long *>(target), __VA_ARGS__);  \
                           else KERNEL<long long, 1><<<grid, kCeThreads, 0, st>>>(logits, static_cast<const long long *>(target), __VA_ARGS__); break; \
            default: DSRL_FAIL(DSRL_ERR_BAD_DTYPE, "cross_entropy: unknown target dtype %d", target_dtype);          \
        }                                                                                                            \
    } while (0)

extern "C" int dsrl_ce_forward(const float *logits, const void *target, int target_dtype, int B, int C, int64_t HW,
                               int64_t ignore_index, int reduction, float *loss_out, void *saved, size_t saved_bytes,
                               dsrl_stream_t stream) {
    int rc = ce_check(logits, target, B, C, HW, reduction);
    if (rc) return rc;
    if (!loss_out || !saved) DSRL_FAIL(DSRL_ERR_BAD_ARG, "cross_entropy: null output");
    if (saved_bytes < dsrl_ce_saved_bytes(B, HW)) DSRL_FAIL(DSRL_ERR_BAD_ARG, "cross_entropy: saved blob too small");
    if (reinterpret_cast<uintptr_t>(saved) & 15) DSRL_FAIL(DSRL_ERR_BAD_ARG, "cross_entropy: saved must be 16-byte aligned");
    if ((rc = require_device())) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (B == 0 || HW == 0) {       // torch: mean over nothing = NaN, sum = 0
        const float v = reduction == DSRL_REDUCE_MEAN ? __builtin_nanf("") : 0.f;
        DSRL_CUDA_TRY(cudaMemsetAsync(saved, 0, sizeof(CeHeader), st));
        DSRL_CUDA_TRY(cudaMemcpyAsync(loss_out, &v, 4, cudaMemcpyHostToDevice, st));
        return DSRL_OK;
    }
    const bool vec = HW % 4 == 0 && (reinterpret_cast<uintptr_t>(logits) & 15) == 0;
    const int per_block = kCeThreads * (vec ? 4 : 1);
    const dim3 grid((unsigned)((HW + per_block - 1) / per_block), (unsigned)B);
    unsigned *ticket = next_ticket_slot(st);
    if (!ticket) return DSRL_ERR_CUDA;
    const size_t lse_off = ce_lse_off(B, HW);
    unsigned char *sv = static_cast<unsigned char *>(saved);
    {
        const long long tiles = (long long)grid.x * B, cap = (long long)device_sm_count() * 8;
        const dim3 grid((unsigned)(tiles < cap ? tiles : cap));          // shadows the (strips, B) grid the backward kernel uses
        CE_DISPATCH(ce_forward_kernel, B, C, (long long)HW, (long long)ignore_index, reduction == DSRL_REDUCE_MEAN, sv, lse_off, ticket, loss_out);
    }
    DSRL_LAUNCH_CHECK();
    return DSRL_OK;
}

extern "C" int dsrl_ce_backward(const float *logits, const void *target, int target_dtype, int B, int C, int64_t HW,
                                int64_t ignore_index, int reduction, const void *saved, size_t saved_bytes,
                                const float *grad_out, float *dlogits, dsrl_stream_t stream) {
    int rc = ce_check(logits, target, B, C, HW, reduction);
    if (rc) return rc;
    if (!saved || !grad_out || !dlogits) DSRL_FAIL(DSRL_ERR_BAD_ARG, "cross_entropy backward: null pointer");
    if (saved_bytes < dsrl_ce_saved_bytes(B, HW)) DSRL_FAIL(DSRL_ERR_BAD_ARG, "cross_entropy backward: saved blob too small");
    if ((rc = require_device())) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (B == 0 || HW == 0) return DSRL_OK;
    const bool vec = HW % 4 == 0 && ((reinterpret_cast<uintptr_t>(logits) | reinterpret_cast<uintptr_t>(dlogits)) & 15) == 0;
    const int per_block = kCeThreads * (vec ? 4 : 1);
    const dim3 grid((unsigned)((HW + per_block - 1) / per_block), (unsigned)B);
    const size_t lse_off = ce_lse_off(B, HW);
    const unsigned char *sv = static_cast<const unsigned char *>(saved);
    CE_DISPATCH(ce_backward_kernel, C, (long long)HW, (long long)ignore_index, reduction == DSRL_REDUCE_MEAN, sv, lse_off, grad_out, dlogits);
    DSRL_LAUNCH_CHECK();
    return DSRL_OK;
}
