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

// Feature-transformer taps.  The reference's stage-3 model feeds FA with two "feature transformers",
// Conv2d(C -> 1, kernel 1, stride 8, no bias) + BatchNorm2d(1) + ReLU, applied to the very tensors the other two losses read:
// the SSSR logits (cross-entropy) and the SISR image (MSE) -- models/DSRL.py:86-95,181,184, train_or_resume.py:435-437.
// The strided 1x1 convolution touches 1/64 of the pixels the loss pass streams anyway, so the forward passes emit it on the
// way (FtTap), and the backward passes add its gradient dz * w[c] at those pixels and collect dw (FtGrad): the transformer
// never launches a kernel of its own, and its sparse input gradient is never materialised as a tensor of zeros.
struct FtTap {            // forward
    const float *w;       // [C] convolution weight; nullptr: no tap
    float *z;             // (B, Hf, Wf) convolution output
    int stride, W, Hf, Wf;
};
struct FtGrad {           // backward
    const float *w;       // nullptr: none
    const float *dz;      // (B, Hf, Wf) gradient w.r.t. the convolution output (BatchNorm + ReLU backward already applied)
    float *dw_part;       // [blocks][C] per-block partial sums of dz * x, written by EVERY block (zeros without strided pixels)
    int stride, W, Hf, Wf;
};
constexpr int kTapMaxC = 32;

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
                                                                float *__restrict__ loss_out, FtTap tap) {
    __shared__ double s_sum[33];
    __shared__ long long s_cnt[33];
    __shared__ int s_last;
    __shared__ float s_tapw[kTapMaxC];
    if (tap.w != nullptr) {
        if (threadIdx.x < C) s_tapw[threadIdx.x] = tap.w[threadIdx.x];
        __syncthreads();
    }
    // persistent CTAs: tile = (image, strip of kCeThreads * VEC pixels); one block-level reduction per CTA at the end
    const long long strips = (HW + (long long)kCeThreads * VEC - 1) / ((long long)kCeThreads * VEC), tiles = strips * nimg;
    float lsum = 0.f;
    int lcnt = 0, lbad = 0;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int b = (int)(tile / strips);
        const long long p0 = ((tile - (long long)b * strips) * kCeThreads + threadIdx.x) * VEC;
        if (p0 >= HW) continue;
        const float *x = logits + (size_t)b * C * HW + p0;
        // feature-transformer tap: the first pixel of this thread lies on the stride grid (host: W % VEC == 0, stride % VEC == 0)
        bool tapped = false;
        long long tap_o = 0;
        float zacc = 0.f;
        if (tap.w != nullptr) {
            const long long ty = p0 / tap.W;
            const int tx = (int)(p0 - ty * tap.W);
            tapped = (ty % tap.stride == 0) && (tx % tap.stride == 0);
            tap_o = ((long long)b * tap.Hf + ty / tap.stride) * tap.Wf + tx / tap.stride;
        }
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
            if (tapped) {
#pragma unroll
                for (int u = 0; u < kCeChunk; ++u)
                    if (c0 + u < C) zacc = fmaf(s_tapw[c0 + u], xv[u][0], zacc);
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
        if (tapped) tap.z[tap_o] = zacc;
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
                                                                 const float *__restrict__ grad_out, float *__restrict__ dlogits, FtGrad ft) {
    __shared__ float s_tapw[kTapMaxC];
    __shared__ float s_dw[kCeThreads / 32][kTapMaxC];
    const int b = blockIdx.y;
    const long long p0 = ((long long)blockIdx.x * kCeThreads + threadIdx.x) * VEC;
    float tdz = 0.f;                                         // dz of this thread's strided pixel (0: none)
    if (ft.w != nullptr) {
        if (threadIdx.x < C) s_tapw[threadIdx.x] = ft.w[threadIdx.x];
        __syncthreads();
        if (p0 < HW) {
            const long long ty = p0 / ft.W;
            const int tx = (int)(p0 - ty * ft.W);
            if ((ty % ft.stride == 0) && (tx % ft.stride == 0))
                tdz = __ldg(ft.dz + ((long long)b * ft.Hf + ty / ft.stride) * ft.Wf + tx / ft.stride);
        }
    }
    const bool live = p0 < HW;
    if (!live && ft.w == nullptr) return;
    const CeHeader *h = reinterpret_cast<const CeHeader *>(saved);
    const float scale = mean ? (float)((double)__ldg(grad_out) / (double)h->valid) : __ldg(grad_out);
    const float *x = logits + (size_t)b * C * HW + p0;
    float *g = dlogits + (size_t)b * C * HW + p0;
    const float2 *lp = reinterpret_cast<const float2 *>(saved + lse_off) + (size_t)b * HW + p0;
    float m2[VEC], l2s[VEC], sc[VEC];                        // (m, log2 s) as the forward kernel left them
    int t[VEC];
    if (!live) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) { m2[v] = 0.f; l2s[v] = 0.f; }
    } else if (VEC == 4) {
        const float4 q0 = __ldg(reinterpret_cast<const float4 *>(lp)), q1 = __ldg(reinterpret_cast<const float4 *>(lp + 2));
        m2[0] = q0.x; l2s[0] = q0.y; m2[1 % VEC] = q0.z; l2s[1 % VEC] = q0.w;
        m2[2 % VEC] = q1.x; l2s[2 % VEC] = q1.y; m2[3 % VEC] = q1.z; l2s[3 % VEC] = q1.w;
    } else {
        const float2 q = __ldg(lp);
        m2[0] = q.x; l2s[0] = q.y;
    }
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        const long long tl = (live && p0 + v < HW) ? ce_load_target(target, (long long)b * HW + p0 + v) : ignore_index;
        const bool valid = tl != ignore_index && tl >= 0 && tl < C;
        t[v] = valid ? (int)tl : -1;
        sc[v] = valid ? scale : 0.f;
    }
    // 4 channels per round: a read + write stream wants resident threads more than registers (8 per round measured slower)
    constexpr int kCh = 4;
    const float *xp = x;                                    // walk the channel planes: one 64-bit add per access
    float *gp = g;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // with a tap every thread walks the loop (the dw sums below are warp-wide); threads past the image load and store nothing
    for (int c0 = 0; c0 < C; c0 += kCh) {
        float xv[kCh][VEC];
        const bool whole = c0 + kCh <= C;
#pragma unroll
        for (int u = 0; u < kCh; ++u) {
#pragma unroll
            for (int v = 0; v < VEC; ++v) xv[u][v] = 0.f;
            if (live && (whole || c0 + u < C)) {
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
            if (live && (whole || c0 + u < C)) {
                float o[VEC];
#pragma unroll
                for (int v = 0; v < VEC; ++v)
                    o[v] = t[v] >= 0 ? (ex2_fast((xv[u][v] - m2[v]) * kLog2e - l2s[v]) - (c0 + u == t[v] ? 1.f : 0.f)) * sc[v] : 0.f;
                if (ft.w != nullptr) o[0] = fmaf(tdz, s_tapw[c0 + u], o[0]);      // the transformer path: dz * w[c] at the strided pixel
                if (VEC == 4) __stcs(reinterpret_cast<float4 *>(gp), make_float4(o[0], o[1 % VEC], o[2 % VEC], o[3 % VEC]));
                else gp[0] = o[0];
                gp += HW;
            }
        }
        if (ft.w != nullptr) {
            // dw[c] += dz * x[c]: fixed-order sums (lanes by shuffle, then warps below): deterministic
#pragma unroll
            for (int u = 0; u < kCh; ++u) {
                if (c0 + u < C) {
                    const float part = warp_sum(tdz * xv[u][0]);
                    if (lane == 0) s_dw[warp][c0 + u] = part;
                }
            }
        }
    }
    if (ft.w != nullptr) {
        __syncthreads();
        if (threadIdx.x < C) {
            float tot = 0.f;
#pragma unroll
            for (int wq = 0; wq < kCeThreads / 32; ++wq) tot += s_dw[wq][threadIdx.x];
            ft.dw_part[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * C + threadIdx.x] = tot;
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

// `tap_w == nullptr`: plain cross-entropy.  With a tap: W is the image width (HW = H * W), the tap needs 16-byte aligned
// logits, W % 4 == 0 and stride % 4 == 0 (the vector path; the first of a thread's four pixels is the only one that can lie on
// the stride grid) or falls back to the scalar path.
static int tap_geom(int64_t HW, int W, int stride, int C, int *Hf, int *Wf) {
    if (W < 1 || stride < 1 || HW % W != 0) DSRL_FAIL(DSRL_ERR_BAD_SHAPE, "feature-transformer tap: bad geometry HW=%lld W=%d stride=%d", (long long)HW, W, stride);
    if (C > kTapMaxC) DSRL_FAIL(DSRL_ERR_UNSUPPORTED, "feature-transformer tap: at most %d input channels", kTapMaxC);
    const int H = (int)(HW / W);
    *Hf = (H - 1) / stride + 1;
    *Wf = (W - 1) / stride + 1;
    return DSRL_OK;
}

static int ce_forward_impl(const float *logits, const void *target, int target_dtype, int B, int C, int64_t HW,
                           int64_t ignore_index, int reduction, float *loss_out, void *saved, size_t saved_bytes,
                           const float *tap_w, float *tap_z, int W, int tap_stride, dsrl_stream_t stream) {
    int rc = ce_check(logits, target, B, C, HW, reduction);
    if (rc) return rc;
    if (!loss_out || !saved) DSRL_FAIL(DSRL_ERR_BAD_ARG, "cross_entropy: null output");
    if (saved_bytes < dsrl_ce_saved_bytes(B, HW)) DSRL_FAIL(DSRL_ERR_BAD_ARG, "cross_entropy: saved blob too small");
    if (reinterpret_cast<uintptr_t>(saved) & 15) DSRL_FAIL(DSRL_ERR_BAD_ARG, "cross_entropy: saved must be 16-byte aligned");
    FtTap tap = {nullptr, nullptr, 1, 1, 1, 1};
    if (tap_w) {
        if (!tap_z) DSRL_FAIL(DSRL_ERR_BAD_ARG, "cross_entropy: tap output missing");
        tap.w = tap_w; tap.z = tap_z; tap.stride = tap_stride; tap.W = W;
        if ((rc = tap_geom(HW, W, tap_stride, C, &tap.Hf, &tap.Wf))) return rc;
    }
    if ((rc = require_device())) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (B == 0 || HW == 0) {       // torch: mean over nothing = NaN, sum = 0
        const float v = reduction == DSRL_REDUCE_MEAN ? __builtin_nanf("") : 0.f;
        DSRL_CUDA_TRY(cudaMemsetAsync(saved, 0, sizeof(CeHeader), st));
        DSRL_CUDA_TRY(cudaMemcpyAsync(loss_out, &v, 4, cudaMemcpyHostToDevice, st));
        return DSRL_OK;
    }
    const bool vec = HW % 4 == 0 && (reinterpret_cast<uintptr_t>(logits) & 15) == 0 && (!tap_w || (W % 4 == 0 && tap_stride % 4 == 0));
    const int per_block = kCeThreads * (vec ? 4 : 1);
    const dim3 grid((unsigned)((HW + per_block - 1) / per_block), (unsigned)B);
    unsigned *ticket = next_ticket_slot(st);
    if (!ticket) return DSRL_ERR_CUDA;
    const size_t lse_off = ce_lse_off(B, HW);
    unsigned char *sv = static_cast<unsigned char *>(saved);
    {
        const long long tiles = (long long)grid.x * B, cap = (long long)device_sm_count() * 8;
        const dim3 grid((unsigned)(tiles < cap ? tiles : cap));          // shadows the (strips, B) grid the backward kernel uses
        CE_DISPATCH(ce_forward_kernel, B, C, (long long)HW, (long long)ignore_index, reduction == DSRL_REDUCE_MEAN, sv, lse_off, ticket, loss_out, tap);
    }
    DSRL_LAUNCH_CHECK();
    return DSRL_OK;
}

static dim3 ce_backward_grid(int B, int64_t HW, bool vec) {
    const int per_block = kCeThreads * (vec ? 4 : 1);
    return dim3((unsigned)((HW + per_block - 1) / per_block), (unsigned)B);
}

static int ce_backward_impl(const float *logits, const void *target, int target_dtype, int B, int C, int64_t HW,
                            int64_t ignore_index, int reduction, const void *saved, size_t saved_bytes,
                            const float *grad_out, float *dlogits, const float *tap_w, const float *tap_dz, float *tap_dw_part,
                            int W, int tap_stride, dsrl_stream_t stream) {
    int rc = ce_check(logits, target, B, C, HW, reduction);
    if (rc) return rc;
    if (!saved || !grad_out || !dlogits) DSRL_FAIL(DSRL_ERR_BAD_ARG, "cross_entropy backward: null pointer");
    if (saved_bytes < dsrl_ce_saved_bytes(B, HW)) DSRL_FAIL(DSRL_ERR_BAD_ARG, "cross_entropy backward: saved blob too small");
    FtGrad ft = {nullptr, nullptr, nullptr, 1, 1, 1, 1};
    if (tap_w) {
        if (!tap_dz || !tap_dw_part) DSRL_FAIL(DSRL_ERR_BAD_ARG, "cross_entropy backward: tap gradient buffers missing");
        ft.w = tap_w; ft.dz = tap_dz; ft.dw_part = tap_dw_part; ft.stride = tap_stride; ft.W = W;
        if ((rc = tap_geom(HW, W, tap_stride, C, &ft.Hf, &ft.Wf))) return rc;
    }
    if ((rc = require_device())) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (B == 0 || HW == 0) return DSRL_OK;
    const bool vec = HW % 4 == 0 && ((reinterpret_cast<uintptr_t>(logits) | reinterpret_cast<uintptr_t>(dlogits)) & 15) == 0 &&
                     (!tap_w || (W % 4 == 0 && tap_stride % 4 == 0));
    const dim3 grid = ce_backward_grid(B, HW, vec);
    const size_t lse_off = ce_lse_off(B, HW);
    const unsigned char *sv = static_cast<const unsigned char *>(saved);
    CE_DISPATCH(ce_backward_kernel, C, (long long)HW, (long long)ignore_index, reduction == DSRL_REDUCE_MEAN, sv, lse_off, grad_out, dlogits, ft);
    DSRL_LAUNCH_CHECK();
    return DSRL_OK;
}

extern "C" int dsrl_ce_forward(const float *logits, const void *target, int target_dtype, int B, int C, int64_t HW,
                               int64_t ignore_index, int reduction, float *loss_out, void *saved, size_t saved_bytes,
                               dsrl_stream_t stream) {
    return ce_forward_impl(logits, target, target_dtype, B, C, HW, ignore_index, reduction, loss_out, saved, saved_bytes, nullptr, nullptr, 1, 1, stream);
}

extern "C" int dsrl_ce_backward(const float *logits, const void *target, int target_dtype, int B, int C, int64_t HW,
                                int64_t ignore_index, int reduction, const void *saved, size_t saved_bytes,
                                const float *grad_out, float *dlogits, dsrl_stream_t stream) {
    return ce_backward_impl(logits, target, target_dtype, B, C, HW, ignore_index, reduction, saved, saved_bytes, grad_out, dlogits, nullptr,
                            nullptr, nullptr, 1, 1, stream);
}

extern "C" int dsrl_ce_forward_tap(const float *logits, const void *target, int target_dtype, int B, int C, int H, int W,
                                   int64_t ignore_index, int reduction, float *loss_out, void *saved, size_t saved_bytes,
                                   const float *tap_w, float *tap_z, int tap_stride, dsrl_stream_t stream) {
    if (!tap_w) DSRL_FAIL(DSRL_ERR_BAD_ARG, "cross_entropy (tap): convolution weight missing");
    return ce_forward_impl(logits, target, target_dtype, B, C, (int64_t)H * W, ignore_index, reduction, loss_out, saved, saved_bytes, tap_w, tap_z, W,
                           tap_stride, stream);
}

extern "C" int64_t dsrl_tap_dw_blocks(int B, int H, int W, int tap_stride) {
    if (B < 0 || H < 1 || W < 1) return 0;
    const int64_t HW = (int64_t)H * W;
    const bool vec = HW % 4 == 0 && W % 4 == 0 && tap_stride % 4 == 0;       // the pointers are checked again at launch: the larger count
    const dim3 g1 = ce_backward_grid(B, HW, false), g4 = ce_backward_grid(B, HW, vec);
    const int64_t n1 = (int64_t)g1.x * g1.y, n4 = (int64_t)g4.x * g4.y;
    return n1 > n4 ? n1 : n4;
}

extern "C" int dsrl_ce_backward_tap(const float *logits, const void *target, int target_dtype, int B, int C, int H, int W,
                                    int64_t ignore_index, int reduction, const void *saved, size_t saved_bytes,
                                    const float *grad_out, float *dlogits, const float *tap_w, const float *tap_dz,
                                    float *tap_dw_part, int tap_stride, int64_t *dw_blocks_used, dsrl_stream_t stream) {
    if (!tap_w) DSRL_FAIL(DSRL_ERR_BAD_ARG, "cross_entropy backward (tap): convolution weight missing");
    const int64_t HW = (int64_t)H * W;
    if (dw_blocks_used) {
        const bool vec = HW % 4 == 0 && ((reinterpret_cast<uintptr_t>(logits) | reinterpret_cast<uintptr_t>(dlogits)) & 15) == 0 &&
                         W % 4 == 0 && tap_stride % 4 == 0;
        const dim3 g = ce_backward_grid(B, HW, vec);
        *dw_blocks_used = (int64_t)g.x * g.y;
    }
    return ce_backward_impl(logits, target, target_dtype, B, C, HW, ignore_index, reduction, saved, saved_bytes, grad_out, dlogits, tap_w, tap_dz,
                            tap_dw_part, W, tap_stride, stream);
}

// ---------------------------------------------------------------------------------------------------------------
// MSE over NCHW tensors with the same taps (the SISR loss, `t.nn.MSELoss()` at train_or_resume.py:117,436)
// ---------------------------------------------------------------------------------------------------------------
namespace dsrl {
namespace {

// One thread = four consecutive pixels of one image, all C channels (C <= kTapMaxC): sum of squared differences, and the tap.
__global__ void __launch_bounds__(kCeThreads) mse_forward_kernel(const float *__restrict__ x, const float *__restrict__ y, int nimg, int C,
                                                                 long long HW, double *__restrict__ partials, unsigned *ticket,
                                                                 double inv_count, float *__restrict__ loss_out, FtTap tap) {
    __shared__ double s_sum[33];
    __shared__ int s_last;
    __shared__ float s_tapw[kTapMaxC];
    if (tap.w != nullptr) {
        if (threadIdx.x < C) s_tapw[threadIdx.x] = tap.w[threadIdx.x];
        __syncthreads();
    }
    const long long strips = (HW + (long long)kCeThreads * 4 - 1) / ((long long)kCeThreads * 4), tiles = strips * nimg;
    float lsum = 0.f;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int b = (int)(tile / strips);
        const long long p0 = ((tile - (long long)b * strips) * kCeThreads + threadIdx.x) * 4;
        if (p0 >= HW) continue;
        bool tapped = false;
        long long tap_o = 0;
        float zacc = 0.f;
        if (tap.w != nullptr) {
            const long long ty = p0 / tap.W;
            const int tx = (int)(p0 - ty * tap.W);
            tapped = (ty % tap.stride == 0) && (tx % tap.stride == 0);
            tap_o = ((long long)b * tap.Hf + ty / tap.stride) * tap.Wf + tx / tap.stride;
        }
        const float *xp = x + (size_t)b * C * HW + p0, *yp = y + (size_t)b * C * HW + p0;
        for (int c = 0; c < C; ++c) {
            const uint4 qx = ldg_stream_u4(xp), qy = ldg_stream_u4(yp);
            const float d0 = __uint_as_float(qx.x) - __uint_as_float(qy.x), d1 = __uint_as_float(qx.y) - __uint_as_float(qy.y);
            const float d2 = __uint_as_float(qx.z) - __uint_as_float(qy.z), d3 = __uint_as_float(qx.w) - __uint_as_float(qy.w);
            lsum = fmaf(d0, d0, lsum); lsum = fmaf(d1, d1, lsum); lsum = fmaf(d2, d2, lsum); lsum = fmaf(d3, d3, lsum);
            if (tapped) zacc = fmaf(s_tapw[c], __uint_as_float(qx.x), zacc);
            xp += HW; yp += HW;
        }
        if (tapped) tap.z[tap_o] = zacc;
    }
    const double bsum = block_sum<double>((double)lsum, s_sum);
    const unsigned nblk = gridDim.x, blk = blockIdx.x;
    if (threadIdx.x == 0) {
        partials[blk] = bsum;
        __threadfence();
        s_last = atomicInc(ticket, nblk - 1) == nblk - 1;          // self-resetting
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    double ts = 0.0;
    for (unsigned i = threadIdx.x; i < nblk; i += kCeThreads) ts += __ldcg(&partials[i]);       // fixed order
    ts = block_sum<double>(ts, s_sum);
    if (threadIdx.x == 0) *loss_out = (float)(ts * inv_count);
}

__global__ void __launch_bounds__(kCeThreads) mse_backward_kernel(const float *__restrict__ x, const float *__restrict__ y, int C, long long HW,
                                                                  float scale2, const float *__restrict__ grad_out, float *__restrict__ dx,
                                                                  FtGrad ft) {
    __shared__ float s_tapw[kTapMaxC];
    __shared__ float s_dw[kCeThreads / 32][kTapMaxC];
    const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long p0 = ((long long)blockIdx.x * kCeThreads + threadIdx.x) * 4;
    const bool live = p0 < HW;
    float tdz = 0.f;
    if (ft.w != nullptr) {
        if (threadIdx.x < C) s_tapw[threadIdx.x] = ft.w[threadIdx.x];
        __syncthreads();
        if (live) {
            const long long ty = p0 / ft.W;
            const int tx = (int)(p0 - ty * ft.W);
            if ((ty % ft.stride == 0) && (tx % ft.stride == 0))
                tdz = __ldg(ft.dz + ((long long)b * ft.Hf + ty / ft.stride) * ft.Wf + tx / ft.stride);
        }
    } else if (!live) {
        return;
    }
    const float sc = __ldg(grad_out) * scale2;                 // grad_out * 2 / count
    const float *xp = x + (size_t)b * C * HW + p0, *yp = y + (size_t)b * C * HW + p0;
    float *gp = dx + (size_t)b * C * HW + p0;
    for (int c = 0; c < C; ++c) {
        float x0 = 0.f;
        if (live) {
            const uint4 qx = ldg_stream_u4(xp), qy = ldg_stream_u4(yp);
            x0 = __uint_as_float(qx.x);
            float4 o = make_float4((x0 - __uint_as_float(qy.x)) * sc, (__uint_as_float(qx.y) - __uint_as_float(qy.y)) * sc,
                                   (__uint_as_float(qx.z) - __uint_as_float(qy.z)) * sc, (__uint_as_float(qx.w) - __uint_as_float(qy.w)) * sc);
            if (ft.w != nullptr) o.x = fmaf(tdz, s_tapw[c], o.x);
            __stcs(reinterpret_cast<float4 *>(gp), o);
        }
        if (ft.w != nullptr) {
            const float part = warp_sum(tdz * x0);
            if (lane == 0) s_dw[warp][c] = part;
        }
        xp += HW; yp += HW; gp += HW;
    }
    if (ft.w != nullptr) {
        __syncthreads();
        if (threadIdx.x < C) {
            float tot = 0.f;
#pragma unroll
            for (int wq = 0; wq < kCeThreads / 32; ++wq) tot += s_dw[wq][threadIdx.x];
            ft.dw_part[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * C + threadIdx.x] = tot;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// BatchNorm2d(1) + ReLU of the two transformers: statistics forward, everything backward (one 1024-thread CTA per
// transformer: the maps are B x 64 x 128 values)
// ---------------------------------------------------------------------------------------------------------------
struct BnParams {          // per transformer
    const float *z;        // (count) convolution output
    const float *gamma, *beta;          // BatchNorm2d(1) weight / bias (1 element each)
    float *run_mean, *run_var;          // running statistics (updated in training mode), 1 element each
};

// bn_out[t] = {a, b, mean, invstd}: the transformer's output is relu(a z + b) with a = gamma * invstd, b = beta - a * mean
__global__ void __launch_bounds__(1024) ft_bn_forward_kernel(BnParams p0, BnParams p1, long long count, float eps, float momentum, int training,
                                                            float *__restrict__ bn_out) {
    __shared__ double scratch[33];
    const BnParams p = blockIdx.x ? p1 : p0;
    double mean, var;
    if (training) {
        double s = 0.0;
        for (long long i = threadIdx.x; i < count; i += blockDim.x) s += (double)p.z[i];
        mean = block_sum<double>(s, scratch) / (double)count;
        double q = 0.0;
        for (long long i = threadIdx.x; i < count; i += blockDim.x) { const double d = (double)p.z[i] - mean; q += d * d; }
        var = block_sum<double>(q, scratch) / (double)count;            // biased: what normalises the batch (torch)
    } else {
        mean = (double)*p.run_mean;
        var = (double)*p.run_var;
    }
    if (threadIdx.x == 0) {
        const float invstd = (float)(1.0 / sqrt(var + (double)eps));
        const float a = *p.gamma * invstd;
        float *o = bn_out + 4 * blockIdx.x;
        o[0] = a; o[1] = *p.beta - a * (float)mean; o[2] = (float)mean; o[3] = invstd;
        if (training) {
            // torch: running = (1 - momentum) * running + momentum * batch statistic, the variance unbiased
            const double unbiased = count > 1 ? var * (double)count / (double)(count - 1) : var;
            *p.run_mean = (1.f - momentum) * *p.run_mean + momentum * (float)mean;
            *p.run_var = (1.f - momentum) * *p.run_var + momentum * (float)unbiased;
        }
    }
}

struct BnBackParams {
    const float *z, *dF;   // convolution output; gradient w.r.t. the transformer output (for a unit upstream gradient)
    float *dz;             // gradient w.r.t. the convolution output
    float *dgb;            // {dgamma, dbeta}
};

__global__ void __launch_bounds__(1024) ft_bn_backward_kernel(BnBackParams p0, BnBackParams p1, long long count, const float *__restrict__ bn,
                                                             const float *__restrict__ go, int training) {
    __shared__ double scratch[33];
    const BnBackParams p = blockIdx.x ? p1 : p0;
    const float a = bn[4 * blockIdx.x], bb = bn[4 * blockIdx.x + 1], mean = bn[4 * blockIdx.x + 2], invstd = bn[4 * blockIdx.x + 3];
    const float g = __ldg(go);
    double s1 = 0.0, s2 = 0.0;
    for (long long i = threadIdx.x; i < count; i += blockDim.x) {
        const float z = p.z[i];
        const float dt = fmaf(a, z, bb) > 0.f ? g * p.dF[i] : 0.f;
        s1 += (double)dt;
        s2 += (double)dt * (double)((z - mean) * invstd);
    }
    const double dbeta = block_sum<double>(s1, scratch), dgamma = block_sum<double>(s2, scratch);
    if (threadIdx.x == 0) { p.dgb[0] = (float)dgamma; p.dgb[1] = (float)dbeta; }
    const float mb = training ? (float)(dbeta / (double)count) : 0.f, mg = training ? (float)(dgamma / (double)count) : 0.f;
    for (long long i = threadIdx.x; i < count; i += blockDim.x) {
        const float z = p.z[i];
        const float dt = fmaf(a, z, bb) > 0.f ? g * p.dF[i] : 0.f;
        p.dz[i] = a * (dt - mb - (z - mean) * invstd * mg);
    }
}

}  // namespace
}  // namespace dsrl

extern "C" size_t dsrl_mse_workspace_bytes(int B, int C, int H, int W) {
    (void)B; (void)C; (void)H; (void)W;
    return (size_t)148 * 16 * sizeof(double) + 256;          // per-CTA partial sums of the persistent forward kernel
}

extern "C" int dsrl_mse_forward(const float *x, const float *y, int B, int C, int H, int W, float *loss_out, void *workspace,
                                size_t workspace_bytes, const float *tap_w, float *tap_z, int tap_stride, dsrl_stream_t stream) {
    if (!x || !y || !loss_out || !workspace) DSRL_FAIL(DSRL_ERR_BAD_ARG, "mse: null pointer");
    if (B < 1 || C < 1 || C > kTapMaxC || H < 1 || W < 1) DSRL_FAIL(DSRL_ERR_BAD_SHAPE, "mse: bad shape B=%d C=%d H=%d W=%d", B, C, H, W);
    const int64_t HW = (int64_t)H * W;
    if (HW % 4 != 0 || ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15))
        DSRL_FAIL(DSRL_ERR_UNSUPPORTED, "mse: needs H * W divisible by 4 and 16-byte aligned tensors");
    if (workspace_bytes < dsrl_mse_workspace_bytes(B, C, H, W)) DSRL_FAIL(DSRL_ERR_BAD_ARG, "mse: workspace too small");
    FtTap tap = {nullptr, nullptr, 1, 1, 1, 1};
    int rc;
    if (tap_w) {
        if (!tap_z || W % 4 != 0 || tap_stride % 4 != 0) DSRL_FAIL(DSRL_ERR_UNSUPPORTED, "mse: the tap needs W and stride divisible by 4");
        tap.w = tap_w; tap.z = tap_z; tap.stride = tap_stride; tap.W = W;
        if ((rc = tap_geom(HW, W, tap_stride, C, &tap.Hf, &tap.Wf))) return rc;
    }
    if ((rc = require_device())) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned *ticket = next_ticket_slot(st);
    if (!ticket) return DSRL_ERR_CUDA;
    const long long tiles = (HW + kCeThreads * 4 - 1) / (kCeThreads * 4) * B;
    long long cap = (long long)device_sm_count() * 8;
    if (cap > 148 * 16) cap = 148 * 16;
    const unsigned grid = (unsigned)(tiles < cap ? tiles : cap);
    mse_forward_kernel<<<grid, kCeThreads, 0, st>>>(x, y, B, C, (long long)HW, static_cast<double *>(workspace), ticket,
                                                    1.0 / ((double)B * C * (double)HW), loss_out, tap);
    DSRL_LAUNCH_CHECK();
    return DSRL_OK;
}

extern "C" int dsrl_mse_backward(const float *x, const float *y, int B, int C, int H, int W, const float *grad_out, float *dx,
                                 const float *tap_w, const float *tap_dz, float *tap_dw_part, int tap_stride, dsrl_stream_t stream) {
    if (!x || !y || !grad_out || !dx) DSRL_FAIL(DSRL_ERR_BAD_ARG, "mse backward: null pointer");
    if (B < 1 || B > 65535 || C < 1 || C > kTapMaxC || H < 1 || W < 1) DSRL_FAIL(DSRL_ERR_BAD_SHAPE, "mse backward: bad shape");
    const int64_t HW = (int64_t)H * W;
    if (HW % 4 != 0 || ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(dx)) & 15))
        DSRL_FAIL(DSRL_ERR_UNSUPPORTED, "mse backward: needs H * W divisible by 4 and 16-byte aligned tensors");
    FtGrad ft = {nullptr, nullptr, nullptr, 1, 1, 1, 1};
    int rc;
    if (tap_w) {
        if (!tap_dz || !tap_dw_part || W % 4 != 0 || tap_stride % 4 != 0) DSRL_FAIL(DSRL_ERR_UNSUPPORTED, "mse backward: bad tap");
        ft.w = tap_w; ft.dz = tap_dz; ft.dw_part = tap_dw_part; ft.stride = tap_stride; ft.W = W;
        if ((rc = tap_geom(HW, W, tap_stride, C, &ft.Hf, &ft.Wf))) return rc;
    }
    if ((rc = require_device())) return rc;
    const dim3 grid = ce_backward_grid(B, HW, true);            // the same (strips of 1024 pixels, B) grid: dsrl_tap_dw_blocks rows of dw_part
    mse_backward_kernel<<<grid, kCeThreads, 0, static_cast<cudaStream_t>(stream)>>>(x, y, C, (long long)HW, (float)(2.0 / ((double)B * C * (double)HW)),
                                                                                   grad_out, dx, ft);
    DSRL_LAUNCH_CHECK();
    return DSRL_OK;
}

extern "C" int dsrl_ft_bn_forward(const float *z1, const float *z2, int64_t count, const float *gamma1, const float *beta1,
                                  float *run_mean1, float *run_var1, const float *gamma2, const float *beta2, float *run_mean2,
                                  float *run_var2, float eps, float momentum, int training, float *bn_out, dsrl_stream_t stream) {
    if (!z1 || !z2 || !gamma1 || !beta1 || !gamma2 || !beta2 || !run_mean1 || !run_var1 || !run_mean2 || !run_var2 || !bn_out)
        DSRL_FAIL(DSRL_ERR_BAD_ARG, "transformer BatchNorm: null pointer");
    if (count < 1) DSRL_FAIL(DSRL_ERR_BAD_SHAPE, "transformer BatchNorm: empty map");
    int rc = require_device();
    if (rc) return rc;
    const BnParams p0 = {z1, gamma1, beta1, run_mean1, run_var1}, p1 = {z2, gamma2, beta2, run_mean2, run_var2};
    ft_bn_forward_kernel<<<2, 1024, 0, static_cast<cudaStream_t>(stream)>>>(p0, p1, (long long)count, eps, momentum, training, bn_out);
    DSRL_LAUNCH_CHECK();
    return DSRL_OK;
}

extern "C" int dsrl_ft_bn_backward(const float *z1, const float *dF1, float *dz1, float *dgb1, const float *z2, const float *dF2,
                                   float *dz2, float *dgb2, int64_t count, const float *bn, const float *grad_out, int training,
                                   dsrl_stream_t stream) {
    if (!z1 || !dF1 || !dz1 || !dgb1 || !z2 || !dF2 || !dz2 || !dgb2 || !bn || !grad_out)
        DSRL_FAIL(DSRL_ERR_BAD_ARG, "transformer BatchNorm backward: null pointer");
    if (count < 1) DSRL_FAIL(DSRL_ERR_BAD_SHAPE, "transformer BatchNorm backward: empty map");
    int rc = require_device();
    if (rc) return rc;
    const BnBackParams p0 = {z1, dF1, dz1, dgb1}, p1 = {z2, dF2, dz2, dgb2};
    ft_bn_backward_kernel<<<2, 1024, 0, static_cast<cudaStream_t>(stream)>>>(p0, p1, (long long)count, bn, grad_out, training);
    DSRL_LAUNCH_CHECK();
    return DSRL_OK;
}
