// Cross-entropy over NCHW logits with an ignore label -- the caller-side loss next to FA in the reference's training step
// (`t.nn.CrossEntropyLoss(ignore_index=IGNORE_CLASS_LABEL)` at command_handlers/train_or_resume.py:116, applied to the
// (B,19,512,1024) SSSR output at :435).  HBM bound.  PyTorch runs it as log_softmax -> nll_loss and two
// backward kernels, about seven passes over the logits-sized tensor; here:
//
//   ce_forward   ONE pass over the logits: online log-sum-exp per pixel, picks x[target], sums the loss and the number of
//                valid pixels (deterministic two-level reduction, last CTA by ticket), keeps lse per pixel (4 B/px)
//   ce_backward  ONE pass: dlogits = (exp(x - lse) - [c == target]) * grad_out / valid   (0 at ignored pixels)
//
// Algorithmic bytes per pixel: forward 4C + sizeof(target) + 4 (lse), backward 4C + 4C + 4 + sizeof(target).
#include "common.cuh"

namespace dsrl {
namespace {

constexpr int kCeThreads = 256;

struct CeHeader {            // first bytes of the saved blob
    double sum;              // sum of the per-pixel losses over valid pixels
    long long valid;         // number of valid pixels
};
struct CePartial { double sum; long long valid; };

constexpr size_t kCePartialsOff = 64;

inline size_t ce_blocks(int B, long long HW) { return (size_t)B * (size_t)((HW + kCeThreads - 1) / kCeThreads); }   // VEC = 1 worst case
inline size_t ce_lse_off(int B, long long HW) { return (kCePartialsOff + ce_blocks(B, HW) * sizeof(CePartial) + 255) / 256 * 256; }

template <typename T> __device__ __forceinline__ long long ce_load_target(const T *t, long long i) { return (long long)t[i]; }

// One thread = VEC consecutive pixels of one image; channel c of those pixels is one VEC*4-byte load, coalesced across the warp.
template <typename TT, int VEC, int kCeChunk = 5>
__global__ void __launch_bounds__(kCeThreads) ce_forward_kernel(const float *__restrict__ logits, const TT *__restrict__ target,
                                                                int C, long long HW, long long ignore_index, int mean,
                                                                unsigned char *__restrict__ saved, size_t lse_off, unsigned *ticket,
                                                                float *__restrict__ loss_out) {
    __shared__ double s_sum[33];
    __shared__ long long s_cnt[33];
    __shared__ int s_last;
    const int b = blockIdx.y;
    const long long p0 = ((long long)blockIdx.x * kCeThreads + threadIdx.x) * VEC;
    float lsum = 0.f;
    int lcnt = 0;
    if (p0 < HW) {
        const float *x = logits + (size_t)b * C * HW + p0;
        float m[VEC], s[VEC], xt[VEC];
        long long t[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) { m[v] = -3.402823466e38f; s[v] = 0.f; xt[v] = 0.f; t[v] = (p0 + v < HW) ? ce_load_target(target, (long long)b * HW + p0 + v) : ignore_index; }
        // kCeChunk channels per round: all their loads are issued before any arithmetic, then one branch-free online-softmax
        // update per pixel (chunk sizes 4..20 measured within 10 % of each other at the training shape; 5 was the fastest)
        for (int c0 = 0; c0 < C; c0 += kCeChunk) {
            float xv[kCeChunk][VEC];
#pragma unroll
            for (int u = 0; u < kCeChunk; ++u) {
                if (c0 + u < C) {
                    if (VEC == 4) {
                        const uint4 q = ldg_stream_u4(x + (size_t)(c0 + u) * HW);
                        xv[u][0] = __uint_as_float(q.x); xv[u][1 % VEC] = __uint_as_float(q.y); xv[u][2 % VEC] = __uint_as_float(q.z); xv[u][3 % VEC] = __uint_as_float(q.w);
                    } else {
                        xv[u][0] = ldg_stream_f32(x + (size_t)(c0 + u) * HW);
                    }
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
                float acc = s[v] * __expf(m[v] - mn);
#pragma unroll
                for (int u = 0; u < kCeChunk; ++u) {
                    acc += __expf(xv[u][v] - mn);              // padding channels: exp(-FLT_MAX - mn) = 0
                    if ((long long)(c0 + u) == t[v]) xt[v] = xv[u][v];
                }
                s[v] = acc;
                m[v] = mn;
            }
        }
        float lse[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            lse[v] = m[v] + __logf(s[v]);
            const bool valid = t[v] != ignore_index && t[v] >= 0 && t[v] < C;
            if (valid) { lsum += lse[v] - xt[v]; ++lcnt; }
        }
        float *lp = reinterpret_cast<float *>(saved + lse_off) + (size_t)b * HW + p0;
        if (VEC == 4) *reinterpret_cast<float4 *>(lp) = make_float4(lse[0], lse[1 % VEC], lse[2 % VEC], lse[3 % VEC]);
        else lp[0] = lse[0];
    }
    const double bsum = block_sum<double>((double)lsum, s_sum);
    const long long bcnt = block_sum<long long>((long long)lcnt, s_cnt);
    CePartial *parts = reinterpret_cast<CePartial *>(saved + kCePartialsOff);
    const unsigned nblk = gridDim.x * gridDim.y, blk = blockIdx.y * gridDim.x + blockIdx.x;
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
        h->sum = ts;
        h->valid = tc;
        *loss_out = mean ? (float)(ts / (double)tc) : (float)ts;    // no valid pixel: 0/0 = NaN, like torch
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
    const float *lp = reinterpret_cast<const float *>(saved + lse_off) + (size_t)b * HW + p0;
    float lse[VEC], sc[VEC];
    long long t[VEC];
    if (VEC == 4) {
        const float4 q = __ldg(reinterpret_cast<const float4 *>(lp));
        lse[0] = q.x; lse[1 % VEC] = q.y; lse[2 % VEC] = q.z; lse[3 % VEC] = q.w;
    } else {
        lse[0] = __ldg(lp);
    }
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        t[v] = (p0 + v < HW) ? ce_load_target(target, (long long)b * HW + p0 + v) : ignore_index;
        const bool valid = t[v] != ignore_index && t[v] >= 0 && t[v] < C;
        sc[v] = valid ? scale : 0.f;
    }
    // 4 channels per round: a read + write stream wants resident threads more than registers (8 per round measured slower)
    constexpr int kCh = 4;
    for (int c0 = 0; c0 < C; c0 += kCh) {
        float xv[kCh][VEC];
#pragma unroll
        for (int u = 0; u < kCh; ++u) {
            if (c0 + u < C) {
                if (VEC == 4) {
                    const uint4 q = ldg_stream_u4(x + (size_t)(c0 + u) * HW);
                    xv[u][0] = __uint_as_float(q.x); xv[u][1 % VEC] = __uint_as_float(q.y); xv[u][2 % VEC] = __uint_as_float(q.z); xv[u][3 % VEC] = __uint_as_float(q.w);
                } else {
                    xv[u][0] = ldg_stream_f32(x + (size_t)(c0 + u) * HW);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < kCh; ++u) {
            if (c0 + u < C) {
                float o[VEC];
#pragma unroll
                for (int v = 0; v < VEC; ++v)
                    o[v] = sc[v] != 0.f ? (__expf(xv[u][v] - lse[v]) - ((long long)(c0 + u) == t[v] ? 1.f : 0.f)) * sc[v] : 0.f;
                if (VEC == 4) __stcs(reinterpret_cast<float4 *>(g + (size_t)(c0 + u) * HW), make_float4(o[0], o[1 % VEC], o[2 % VEC], o[3 % VEC]));
                else g[(size_t)(c0 + u) * HW] = o[0];
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
    return ce_lse_off(B, HW) + (size_t)B * (size_t)HW * 4 + 16;
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
    unsigned *ticket = next_ticket_slot();
    if (!ticket) return DSRL_ERR_CUDA;
    const size_t lse_off = ce_lse_off(B, HW);
    unsigned char *sv = static_cast<unsigned char *>(saved);
    CE_DISPATCH(ce_forward_kernel, C, (long long)HW, (long long)ignore_index, reduction == DSRL_REDUCE_MEAN, sv, lse_off, ticket, loss_out);
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
