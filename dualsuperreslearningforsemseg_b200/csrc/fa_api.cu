// C-ABI entry points of the FA loss (argument validation + mode dispatch).  See include/dsrl_b200.h.
#include <string.h>

#include "common.cuh"

namespace dsrl {
size_t fa_ref_saved_bytes(int B, int C, int H, int W, int k);
size_t fa_ref_workspace_bytes(int B, int C, int H, int W, int k);
int fa_ref_forward(const float *x1, const float *x2, int B, int C, int H, int W, int k, int reduction, int need_grad,
                   float *loss_out, void *saved, size_t saved_bytes, void *ws, size_t ws_bytes, cudaStream_t st);
int fa_ref_backward(const void *saved, size_t saved_bytes, const float *grad_out, float *dx1, float *dx2, int B, int C,
                    int H, int W, int k, int reduction, void *ws, size_t ws_bytes, cudaStream_t st);

int fa_ref_forward_backward(const float *x1, const float *x2, int B, int C, int H, int W, int k, int reduction,
                            const float *grad_out, float *loss_out, float *dx1, float *dx2, void *saved, size_t saved_bytes,
                            void *ws, size_t ws_bytes, cudaStream_t st, const float *bn);

size_t fa_pos_saved_bytes(int precision, int B, int C1, int C2, int H, int W, int k);
size_t fa_pos_workspace_bytes(int precision, int B, int C1, int C2, int H, int W, int k);
int fa_pos_forward(int precision, const float *x1, const float *x2, int B, int C1, int C2, int H, int W, int k,
                   int reduction, int need_grad, float *loss_out, void *saved, size_t saved_bytes, void *ws,
                   size_t ws_bytes, cudaStream_t st);
int fa_pos_backward(int precision, const float *x1, const float *x2, const void *saved, size_t saved_bytes,
                    const float *grad_out, float *dx1, float *dx2, int B, int C1, int C2, int H, int W, int k,
                    int reduction, void *ws, size_t ws_bytes, cudaStream_t st);
int fa_pos_forward_backward(int precision, const float *x1, const float *x2, int B, int C1, int C2, int H, int W, int k,
                            int reduction, const float *grad_out, float *loss_out, float *dx1, float *dx2, void *saved,
                            size_t saved_bytes, void *ws, size_t ws_bytes, cudaStream_t st);
}  // namespace dsrl

using namespace dsrl;

// d *= *go, skipped entirely (one load per thread, no stores) when the upstream gradient is exactly 1
__global__ void __launch_bounds__(256) scale_grads_kernel(const float *__restrict__ go, float *__restrict__ d1, long long n1,
                                                          float *__restrict__ d2, long long n2) {
    const float s = __ldg(go);
    if (s == 1.f) return;
    const long long stride = (long long)gridDim.x * blockDim.x, t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (int which = 0; which < 2; ++which) {
        float *d = which ? d2 : d1;
        const long long n = which ? n2 : n1;
        if (!d) continue;
        if ((reinterpret_cast<uintptr_t>(d) & 15) == 0) {
            float4 *d4 = reinterpret_cast<float4 *>(d);
            for (long long i = t0; i < n / 4; i += stride) { float4 v = d4[i]; v.x *= s; v.y *= s; v.z *= s; v.w *= s; d4[i] = v; }
            for (long long i = n / 4 * 4 + t0; i < n; i += stride) d[i] *= s;
        } else {
            for (long long i = t0; i < n; i += stride) d[i] *= s;
        }
    }
}

extern "C" int dsrl_scale_grads(const float *grad_out, float *d1, int64_t n1, float *d2, int64_t n2, dsrl_stream_t stream) {
    if (!grad_out) DSRL_FAIL(DSRL_ERR_BAD_ARG, "scale_grads: null upstream gradient");
    if ((!d1 || n1 <= 0) && (!d2 || n2 <= 0)) return DSRL_OK;
    int rc = require_device();
    if (rc) return rc;
    const long long n = (long long)(d1 ? n1 : 0) + (long long)(d2 ? n2 : 0);
    long long blocks = (n / 4 + 255) / 256;
    const long long cap = (long long)device_sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    scale_grads_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(grad_out, d1, (long long)n1, d2, (long long)n2);
    DSRL_LAUNCH_CHECK();
    return DSRL_OK;
}

static int check_mode(int mode, int C1, int C2, int reduction) {
    if (mode != DSRL_FA_REFERENCE && mode != DSRL_FA_POSITION) DSRL_FAIL(DSRL_ERR_BAD_ARG, "FA: unknown mode %d", mode);
    if (reduction != DSRL_REDUCE_NONE && reduction != DSRL_REDUCE_MEAN && reduction != DSRL_REDUCE_SUM)
        DSRL_FAIL(DSRL_ERR_BAD_ARG, "FA: unknown reduction %d", reduction);
    // FALoss.py:20 -- "Feature map inputs to FALoss.forward() should be of same size"
    if (mode == DSRL_FA_REFERENCE && C1 != C2) DSRL_FAIL(DSRL_ERR_BAD_SHAPE, "FA(reference): inputs must have the same shape (C1=%d, C2=%d)", C1, C2);
    if (mode == DSRL_FA_POSITION && reduction == DSRL_REDUCE_NONE)
        DSRL_FAIL(DSRL_ERR_UNSUPPORTED, "FA(position): reduction='none' would materialise the N x N affinity; use mean or sum");
    return DSRL_OK;
}

extern "C" size_t dsrl_fa_saved_bytes(int mode, int precision, int B, int C1, int C2, int H, int W, int k) {
    if (mode == DSRL_FA_REFERENCE) return C1 == C2 ? fa_ref_saved_bytes(B, C1, H, W, k) : 0;
    if (mode == DSRL_FA_POSITION) return fa_pos_saved_bytes(precision, B, C1, C2, H, W, k);
    return 0;
}

extern "C" size_t dsrl_fa_workspace_bytes(int mode, int precision, int B, int C1, int C2, int H, int W, int k) {
    if (mode == DSRL_FA_REFERENCE) return C1 == C2 ? fa_ref_workspace_bytes(B, C1, H, W, k) : 0;
    if (mode == DSRL_FA_POSITION) return fa_pos_workspace_bytes(precision, B, C1, C2, H, W, k);
    return 0;
}

extern "C" int dsrl_fa_sign_stats(const void *saved, uint64_t *out, dsrl_stream_t stream) {
    if (!saved || !out) DSRL_FAIL(DSRL_ERR_BAD_ARG, "FA sign stats: null pointer");
    int rc = require_device();
    if (rc) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned long long h[4];
    DSRL_CUDA_TRY(cudaMemcpyAsync(h, static_cast<const unsigned char *>(saved) + 8, sizeof(h), cudaMemcpyDeviceToHost, st));
    DSRL_CUDA_TRY(cudaStreamSynchronize(st));
    float worst;
    const unsigned lo = (unsigned)(h[3] & 0xffffffffu);
    memcpy(&worst, &lo, 4);
    out[0] = h[0]; out[1] = h[1]; out[2] = h[2]; out[3] = (uint64_t)(worst * 1e6f);
    return DSRL_OK;
}

extern "C" int dsrl_fa_forward(int mode, int precision, const float *x1, const float *x2, int B, int C1, int C2, int H,
                               int W, int k, int reduction, int need_grad, float *loss_out, void *saved,
                               size_t saved_bytes, void *workspace, size_t workspace_bytes, dsrl_stream_t stream) {
    int rc = check_mode(mode, C1, C2, reduction);
    if (rc) return rc;
    if (!x1 || !x2 || !loss_out || !saved || !workspace) DSRL_FAIL(DSRL_ERR_BAD_ARG, "FA forward: null pointer");
    if ((rc = require_device())) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (mode == DSRL_FA_REFERENCE)
        return fa_ref_forward(x1, x2, B, C1, H, W, k, reduction, need_grad, loss_out, saved, saved_bytes, workspace, workspace_bytes, st);
    return fa_pos_forward(precision, x1, x2, B, C1, C2, H, W, k, reduction, need_grad, loss_out, saved, saved_bytes, workspace, workspace_bytes, st);
}

extern "C" int dsrl_fa_backward(int mode, int precision, const float *x1, const float *x2, const void *saved,
                                size_t saved_bytes, const float *grad_out, float *dx1, float *dx2, int B, int C1, int C2,
                                int H, int W, int k, int reduction, void *workspace, size_t workspace_bytes,
                                dsrl_stream_t stream) {
    int rc = check_mode(mode, C1, C2, reduction);
    if (rc) return rc;
    if (!saved || !grad_out) DSRL_FAIL(DSRL_ERR_BAD_ARG, "FA backward: null pointer");
    if (!workspace && mode == DSRL_FA_REFERENCE && reduction == DSRL_REDUCE_NONE)
        DSRL_FAIL(DSRL_ERR_BAD_ARG, "FA backward: reference mode with reduction 'none' needs the workspace");
    if (!dx1 && !dx2) return DSRL_OK;
    if ((rc = require_device())) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (mode == DSRL_FA_REFERENCE)
        return fa_ref_backward(saved, saved_bytes, grad_out, dx1, dx2, B, C1, H, W, k, reduction, workspace, workspace_bytes, st);
    return fa_pos_backward(precision, x1, x2, saved, saved_bytes, grad_out, dx1, dx2, B, C1, C2, H, W, k, reduction, workspace, workspace_bytes, st);
}

extern "C" int dsrl_fa_forward_backward(int mode, int precision, const float *x1, const float *x2, int B, int C1, int C2, int H,
                                        int W, int k, int reduction, const float *grad_out, float *loss_out, float *dx1,
                                        float *dx2, void *saved, size_t saved_bytes, void *workspace, size_t workspace_bytes,
                                        dsrl_stream_t stream) {
    int rc = check_mode(mode, C1, C2, reduction);
    if (rc) return rc;
    if (!x1 || !x2 || !grad_out || !loss_out || !saved || !workspace) DSRL_FAIL(DSRL_ERR_BAD_ARG, "FA forward_backward: null pointer");
    if ((rc = require_device())) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (mode == DSRL_FA_REFERENCE) {
        rc = fa_ref_forward_backward(x1, x2, B, C1, H, W, k, reduction, grad_out, loss_out, dx1, dx2, saved, saved_bytes, workspace,
                                     workspace_bytes, st, nullptr);
        if (rc != DSRL_ERR_UNSUPPORTED) return rc;       // DSRL_OK or a real error; otherwise fall through to the two-call path
    }
    if (mode == DSRL_FA_POSITION)
        return fa_pos_forward_backward(precision, x1, x2, B, C1, C2, H, W, k, reduction, grad_out, loss_out, dx1, dx2, saved, saved_bytes,
                                       workspace, workspace_bytes, st);
    rc = dsrl_fa_forward(mode, precision, x1, x2, B, C1, C2, H, W, k, reduction, 1, loss_out, saved, saved_bytes, workspace,
                         workspace_bytes, stream);
    if (rc) return rc;
    return dsrl_fa_backward(mode, precision, x1, x2, saved, saved_bytes, grad_out, dx1, dx2, B, C1, C2, H, W, k, reduction, workspace,
                            workspace_bytes, stream);
}

extern "C" int dsrl_fa_forward_backward_transformed(const float *z1, const float *z2, const float *bn, int B, int H, int W, int k,
                                                    int reduction, const float *grad_out, float *loss_out, float *df1, float *df2,
                                                    void *saved, size_t saved_bytes, void *workspace, size_t workspace_bytes,
                                                    dsrl_stream_t stream) {
    if (!z1 || !z2 || !bn || !grad_out || !loss_out || !df1 || !df2 || !saved || !workspace)
        DSRL_FAIL(DSRL_ERR_BAD_ARG, "FA (transformed): null pointer");
    if (reduction != DSRL_REDUCE_MEAN && reduction != DSRL_REDUCE_SUM) DSRL_FAIL(DSRL_ERR_UNSUPPORTED, "FA (transformed): reduction must be mean or sum");
    int rc = require_device();
    if (rc) return rc;
    rc = fa_ref_forward_backward(z1, z2, B, 1, H, W, k, reduction, grad_out, loss_out, df1, df2, saved, saved_bytes, workspace, workspace_bytes,
                                 static_cast<cudaStream_t>(stream), bn);
    if (rc == DSRL_ERR_UNSUPPORTED)
        set_last_error("FA (transformed): only the single-launch geometries (pooled map at most 32 x 16, H and W divisible by k, k divisible by 4)");
    return rc;
}
