// K4 -- segmentation counts for mIoU / mean accuracy (sm_100a).
//
// Replaces the host-side NumPy passes of the reference: `pred+1`, `target+1`, `pred*mask`,
// `pred*(pred==target)` and three `np.histogram` calls (metrices/mIoU.py:21-29), plus
// `((pred==target)*mask).sum()` and `mask.sum()` (metrices/Accuracy.py:19-20).  Output per update is the exact
// int64 row [area_pred | area_inter | area_target | correct | valid]; the float64 finish stays on the host
// with the reference's own NumPy expressions so the final numbers are bit-identical.
//
// Design (HBM-bound streaming kernel, 10 B/px at the reference dtypes int64/uint8/bool):
//  * one CTA per tile of <= 65536 consecutive pixels of ONE update (per-update rows are what the reference's
//    "mean over update() calls" needs, mIoU.py:35,40), so 500 updates cost one launch;
//  * 128-bit streaming loads (ld.global.nc.L1::no_allocate), 4 independent vectors in flight per thread;
//  * NO atomics in the pixel loop: every thread owns private 8-bit counters in shared memory (<= 130 pixels
//    per thread per tile, so they cannot overflow), four rows packed per 32-bit word at word index
//    (row/4)*BLOCK + tid -- the bank is the lane id, so the updates are bank-conflict free for ANY label
//    distribution.  Two increments per pixel:
//        row1 = mask ? 2*(pred in range ? pred : NC) + (pred == target) : dump
//        row2 = target in range ? K2 + target : dump
//    This is contention-free for any label distribution (a spatially coherent label map makes a warp-private
//    atomic histogram serialise 32-way; this layout does not care);
//  * tile flush: dp4a byte sums + warp shuffle per row, then <= 3*NC+2 u64 atomics per tile.
#include "common.cuh"

namespace dsrl {
namespace {

#ifndef DSRL_SEG_MINB
#define DSRL_SEG_MINB 4   /* measured on B200 (profiles/r01_seg_counts_sweep.md): 4 CTAs x 512 threads, 32 regs */
#endif
#ifndef DSRL_SEG_UNROLL
#define DSRL_SEG_UNROLL 1
#endif
constexpr int kPxPerThread = 128;  // vector-body pixels per thread per tile (+ <= 2 head/tail) -- must stay < 254
// The logits kernel moves 19x more bytes per pixel, so a much smaller tile already amortises the flush and gives one
// image (2 Mpx) 256 CTAs instead of 64 (measured: 64 CTAs per launch left the kernel at 17 % of the HBM peak).
constexpr int kLogitsPxPerThread = 32;

template <typename T, int N>
struct alignas(sizeof(T) * N > 16 ? 16 : sizeof(T) * N) Pack {
    T v[N];
};

template <typename T, int N>
__device__ __forceinline__ Pack<T, N> load_pack(const T *p) {
    constexpr int bytes = sizeof(T) * N;
    Pack<T, N> out;
    if constexpr (bytes % 16 == 0) {
#pragma unroll
        for (int q = 0; q < bytes / 16; ++q) {
            uint4 r = ldg_stream_u4(reinterpret_cast<const unsigned char *>(p) + 16 * q);
            memcpy(reinterpret_cast<unsigned char *>(&out) + 16 * q, &r, 16);
        }
    } else if constexpr (bytes == 8) {
        uint2 r = ldg_stream_u2(p);
        memcpy(&out, &r, 8);
    } else if constexpr (bytes == 4) {
        uint32_t r = ldg_stream_u32(p);
        memcpy(&out, &r, 4);
    } else if constexpr (bytes == 2) {
        uint16_t r = ldg_stream_u16(p);
        memcpy(&out, &r, 2);
    } else {
        static_assert(bytes == 1, "unsupported pack width");
        uint8_t r = ldg_stream_u8(p);
        memcpy(&out, &r, 1);
    }
    return out;
}

struct RowMap {
    int nc;     // number of classes
    int k2;     // first row of the target histogram = 2*(nc+1)
    int dump;   // row for "nothing to count" = k2 + nc
    int rows;   // dump + 1
};

template <int BLOCK>
__device__ __forceinline__ void bump(uint8_t *mine, unsigned r) {
    // byte (r & 3) of word (r >> 2) * BLOCK + tid
    uint8_t *c = mine + (((r & ~3u) * BLOCK) | (r & 3u));
    *c = (uint8_t)(*c + 1);
}

template <int BLOCK>
__device__ __forceinline__ void count_pixel(uint8_t *mine, const RowMap &rm, long long p, long long t, bool m) {
    const bool eq = (p == t);
    const bool pin = (unsigned long long)p < (unsigned long long)rm.nc;
    const bool tin = (unsigned long long)t < (unsigned long long)rm.nc;
    const unsigned r1 = m ? (((pin ? (unsigned)p : (unsigned)rm.nc) << 1) | (unsigned)eq) : (unsigned)rm.dump;
    const unsigned r2 = tin ? (unsigned)rm.k2 + (unsigned)t : (unsigned)rm.dump;
    bump<BLOCK>(mine, r1);
    bump<BLOCK>(mine, r2);
}

// Sums every counter row over the block's threads and adds the derived outputs to the update's global row.
template <int BLOCK>
__device__ __forceinline__ void flush_tile(const uint8_t *sm, int *rowsum, const RowMap &rm, unsigned long long *out_row) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    constexpr int NW = BLOCK / 32;
    const unsigned *words = reinterpret_cast<const unsigned *>(sm);
    const int nq = (rm.rows + 3) >> 2;
    __syncthreads();
    for (int q = wid; q < nq; q += NW) {
        unsigned s0 = 0, s1 = 0, s2 = 0, s3 = 0;
#pragma unroll
        for (int j = 0; j < BLOCK / 32; ++j) {
            const unsigned w = words[q * BLOCK + j * 32 + lane];
            s0 = __dp4a(w, 0x00000001u, s0);
            s1 = __dp4a(w, 0x00000100u, s1);
            s2 = __dp4a(w, 0x00010000u, s2);
            s3 = __dp4a(w, 0x01000000u, s3);
        }
        s0 = warp_sum((int)s0); s1 = warp_sum((int)s1); s2 = warp_sum((int)s2); s3 = warp_sum((int)s3);
        if (lane == 0) { rowsum[4 * q] = s0; rowsum[4 * q + 1] = s1; rowsum[4 * q + 2] = s2; rowsum[4 * q + 3] = s3; }
    }
    __syncthreads();
    const int nc = rm.nc;
    for (int o = threadIdx.x; o < 3 * nc + 2; o += BLOCK) {
        long long v;
        if (o < nc) v = rowsum[2 * o] + rowsum[2 * o + 1];
        else if (o < 2 * nc) v = rowsum[2 * (o - nc) + 1];
        else if (o < 3 * nc) v = rowsum[rm.k2 + (o - 2 * nc)];
        else {
            long long c = 0, a = 0;
            for (int s = 0; s <= nc; ++s) { c += rowsum[2 * s + 1]; a += rowsum[2 * s] + rowsum[2 * s + 1]; }
            v = (o == 3 * nc) ? c : a;
        }
        if (v) atomicAdd(out_row + o, (unsigned long long)v);
    }
}

template <int BLOCK>
__device__ __forceinline__ void zero_counters(uint8_t *sm, int bytes) {
    uint4 *p = reinterpret_cast<uint4 *>(sm);
    for (int i = threadIdx.x; i < bytes / 16; i += BLOCK) p[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
}

// ---- labels in, counts out ---------------------------------------------------------------------------------
template <typename PT, typename TT, bool HAS_MASK, int VEC, int BLOCK>
__global__ void __launch_bounds__(BLOCK, BLOCK == 512 ? DSRL_SEG_MINB : 4) seg_counts_kernel(const PT *__restrict__ pred, const TT *__restrict__ target,
                                                           const uint8_t *__restrict__ mask, long long npix,
                                                           int nc, int ignore_label, int tiles_per_update,
                                                           unsigned long long *__restrict__ counts) {
    constexpr long long TILE = (long long)BLOCK * kPxPerThread;
    extern __shared__ __align__(16) uint8_t sm[];
    const long long u = blockIdx.x / tiles_per_update;
    const int tt = blockIdx.x % tiles_per_update;
    const long long s = u * npix, e = s + npix;
    const long long cell = s / TILE + tt;
    const long long lo = max(cell * TILE, s), hi = min((cell + 1) * TILE, e);
    if (lo >= hi) return;

    RowMap rm{nc, 2 * (nc + 1), 2 * (nc + 1) + nc, 2 * (nc + 1) + nc + 1};
    const int counter_bytes = ((rm.rows + 3) >> 2) * BLOCK * 4;
    int *rowsum = reinterpret_cast<int *>(sm + counter_bytes);
    zero_counters<BLOCK>(sm, counter_bytes);
    uint8_t *mine = sm + threadIdx.x * 4;

    auto one = [&](long long g) {
        const long long p = (long long)pred[g], t = (long long)target[g];
        const bool m = HAS_MASK ? (mask[g] != 0) : (t != (long long)ignore_label);
        count_pixel<BLOCK>(mine, rm, p, t, m);
    };

    // scalar head / tail so the vector body is aligned in all three streams
    const long long lo_al = min(hi, (lo + VEC - 1) / VEC * VEC);
    const int nv = (int)((hi - lo_al) / VEC);             // <= TILE / VEC: 32-bit indexing inside the tile
    const long long hi_al = lo_al + (long long)nv * VEC;
    if (lo + threadIdx.x < lo_al) one(lo + threadIdx.x);
    if (hi_al + threadIdx.x < hi) one(hi_al + threadIdx.x);

    const PT *__restrict__ pb = pred + lo_al;
    const TT *__restrict__ tb = target + lo_al;
    const uint8_t *__restrict__ mb = HAS_MASK ? mask + lo_al : nullptr;
    constexpr int UNROLL = VEC >= 8 ? 1 : (VEC == 4 ? 2 : DSRL_SEG_UNROLL);  // ~64-80 B of loads in flight per thread
    for (int v0 = threadIdx.x; v0 < nv; v0 += UNROLL * BLOCK) {
        Pack<PT, VEC> pp[UNROLL];
        Pack<TT, VEC> tp[UNROLL];
        Pack<uint8_t, VEC> mp[UNROLL];
#pragma unroll
        for (int j = 0; j < UNROLL; ++j) {
            const int v = v0 + j * BLOCK;
            if (v < nv) {
                pp[j] = load_pack<PT, VEC>(pb + v * VEC);
                tp[j] = load_pack<TT, VEC>(tb + v * VEC);
                if (HAS_MASK) mp[j] = load_pack<uint8_t, VEC>(mb + v * VEC);
            }
        }
#pragma unroll
        for (int j = 0; j < UNROLL; ++j) {
            if (v0 + j * BLOCK < nv) {
#pragma unroll
                for (int q = 0; q < VEC; ++q) {
                    const long long p = (long long)pp[j].v[q], t = (long long)tp[j].v[q];
                    const bool m = HAS_MASK ? (mp[j].v[q] != 0) : (t != (long long)ignore_label);
                    count_pixel<BLOCK>(mine, rm, p, t, m);
                }
            }
        }
    }
    flush_tile<BLOCK>(sm, rowsum, rm, counts + (size_t)u * (3 * nc + 2));
}

// ---- logits in (fused argmax), counts out --------------------------------------------------------------------
// Reads each fp32 logit exactly once (NC*4 B/px) and never materialises the int64 prediction map unless asked.
template <typename TT, bool HAS_MASK, int VEC, int BLOCK>
__global__ void __launch_bounds__(BLOCK) seg_counts_logits_kernel(const float *__restrict__ logits, const TT *__restrict__ target,
                                                                  const uint8_t *__restrict__ mask, long long batch, long long hw,
                                                                  int nc, int ignore_label, int tiles_per_image,
                                                                  unsigned long long *__restrict__ counts,
                                                                  long long *__restrict__ pred_out) {
    constexpr long long TILE = (long long)BLOCK * kLogitsPxPerThread;
    extern __shared__ __align__(16) uint8_t sm[];
    const long long img = blockIdx.x / tiles_per_image;  // global image index = u*batch + i
    const int tt = blockIdx.x % tiles_per_image;
    const long long u = img / batch;
    const long long lo = (long long)tt * TILE, hi = min(lo + TILE, hw);
    if (lo >= hi) return;

    RowMap rm{nc, 2 * (nc + 1), 2 * (nc + 1) + nc, 2 * (nc + 1) + nc + 1};
    const int counter_bytes = ((rm.rows + 3) >> 2) * BLOCK * 4;
    int *rowsum = reinterpret_cast<int *>(sm + counter_bytes);
    zero_counters<BLOCK>(sm, counter_bytes);
    uint8_t *mine = sm + threadIdx.x * 4;

    const float *lg = logits + img * nc * hw;
    const TT *tg = target + img * hw;
    const uint8_t *mk = HAS_MASK ? mask + img * hw : nullptr;
    long long *po = pred_out ? pred_out + img * hw : nullptr;

    const long long nv = (hi - lo + VEC - 1) / VEC;  // VEC > 1 only when hw % VEC == 0, so no ragged vector
    for (long long v = threadIdx.x; v < nv; v += BLOCK) {
        const long long g = lo + v * VEC;
        float best[VEC];
        int idx[VEC];
        {
            Pack<float, VEC> x = load_pack<float, VEC>(lg + g);
#pragma unroll
            for (int q = 0; q < VEC; ++q) { best[q] = x.v[q]; idx[q] = 0; }
        }
#pragma unroll 6
        for (int c = 1; c < nc; ++c) {
            Pack<float, VEC> x = load_pack<float, VEC>(lg + (long long)c * hw + g);
#pragma unroll
            for (int q = 0; q < VEC; ++q) {
                const float xv = x.v[q];
                // first maximum wins; a NaN is "the maximum" and the first NaN sticks (numpy / torch argmax)
                const bool take = (xv > best[q]) || ((xv != xv) && (best[q] == best[q]));
                best[q] = take ? xv : best[q];
                idx[q] = take ? c : idx[q];
            }
        }
        Pack<TT, VEC> tp = load_pack<TT, VEC>(tg + g);
        Pack<uint8_t, VEC> mp;
        if (HAS_MASK) mp = load_pack<uint8_t, VEC>(mk + g);
#pragma unroll
        for (int q = 0; q < VEC; ++q) {
            const long long t = (long long)tp.v[q];
            const bool m = HAS_MASK ? (mp.v[q] != 0) : (t != (long long)ignore_label);
            count_pixel<BLOCK>(mine, rm, (long long)idx[q], t, m);
            if (po) po[g + q] = idx[q];
        }
    }
    flush_tile<BLOCK>(sm, rowsum, rm, counts + (size_t)u * (3 * nc + 2));
}

// ---- host dispatch -----------------------------------------------------------------------------------------
inline bool aligned_to(const void *p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

template <typename K>
int prepare_smem(K kernel, size_t smem) {
    if (smem > 48 * 1024) DSRL_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    return DSRL_OK;
}

template <typename PT, typename TT, bool HAS_MASK, int VEC, int BLOCK>
int launch_counts(const void *pred, const void *target, const uint8_t *mask, int64_t num_updates, int64_t npix,
                  int nc, int ignore_label, int64_t *counts, cudaStream_t st) {
    constexpr long long TILE = (long long)BLOCK * kPxPerThread;
    const int rows = 3 * nc + 3;
    const size_t smem = (size_t)((rows + 3) / 4) * BLOCK * 4 + (size_t)(rows + 4) * sizeof(int);
    auto kern = seg_counts_kernel<PT, TT, HAS_MASK, VEC, BLOCK>;
    int rc = prepare_smem(kern, smem);
    if (rc) return rc;
    const long long tiles_per_update = (npix + TILE - 1) / TILE + 1;  // +1: tiles sit on the absolute TILE grid
    const long long grid = tiles_per_update * num_updates;
    if (grid > 0x7fffffffLL) DSRL_FAIL(DSRL_ERR_UNSUPPORTED, "seg_counts: too many tiles (%lld)", grid);
    kern<<<(unsigned)grid, BLOCK, smem, st>>>(static_cast<const PT *>(pred), static_cast<const TT *>(target), mask,
                                               (long long)npix, nc, ignore_label, (int)tiles_per_update,
                                               reinterpret_cast<unsigned long long *>(counts));
    DSRL_LAUNCH_CHECK();
    return DSRL_OK;
}

template <typename PT, typename TT, int BLOCK>
int dispatch_vec(const void *pred, const void *target, const uint8_t *mask, int64_t num_updates, int64_t npix, int nc,
                 int ignore_label, int64_t *counts, cudaStream_t st) {
    constexpr int W = sizeof(PT) > sizeof(TT) ? sizeof(PT) : sizeof(TT);
    constexpr int VEC = 16 / W;
    const bool ok = aligned_to(pred, sizeof(PT) * VEC) && aligned_to(target, sizeof(TT) * VEC) &&
                    (mask == nullptr || aligned_to(mask, VEC));
    if (ok) {
        return mask ? launch_counts<PT, TT, true, VEC, BLOCK>(pred, target, mask, num_updates, npix, nc, ignore_label, counts, st)
                    : launch_counts<PT, TT, false, VEC, BLOCK>(pred, target, mask, num_updates, npix, nc, ignore_label, counts, st);
    }
    return mask ? launch_counts<PT, TT, true, 1, BLOCK>(pred, target, mask, num_updates, npix, nc, ignore_label, counts, st)
                : launch_counts<PT, TT, false, 1, BLOCK>(pred, target, mask, num_updates, npix, nc, ignore_label, counts, st);
}

template <typename PT, int BLOCK>
int dispatch_target(const void *pred, const void *target, int target_dtype, const uint8_t *mask, int64_t num_updates,
                    int64_t npix, int nc, int ignore_label, int64_t *counts, cudaStream_t st) {
    switch (target_dtype) {
        case DSRL_U8: return dispatch_vec<PT, uint8_t, BLOCK>(pred, target, mask, num_updates, npix, nc, ignore_label, counts, st);
        case DSRL_I32: return dispatch_vec<PT, int32_t, BLOCK>(pred, target, mask, num_updates, npix, nc, ignore_label, counts, st);
        case DSRL_I64: return dispatch_vec<PT, long long, BLOCK>(pred, target, mask, num_updates, npix, nc, ignore_label, counts, st);
    }
    DSRL_FAIL(DSRL_ERR_BAD_DTYPE, "seg_counts: unknown target dtype %d", target_dtype);
}

template <int BLOCK>
int dispatch_pred(const void *pred, int pred_dtype, const void *target, int target_dtype, const uint8_t *mask,
                  int64_t num_updates, int64_t npix, int nc, int ignore_label, int64_t *counts, cudaStream_t st) {
    switch (pred_dtype) {
        case DSRL_U8: return dispatch_target<uint8_t, BLOCK>(pred, target, target_dtype, mask, num_updates, npix, nc, ignore_label, counts, st);
        case DSRL_I32: return dispatch_target<int32_t, BLOCK>(pred, target, target_dtype, mask, num_updates, npix, nc, ignore_label, counts, st);
        case DSRL_I64: return dispatch_target<long long, BLOCK>(pred, target, target_dtype, mask, num_updates, npix, nc, ignore_label, counts, st);
    }
    DSRL_FAIL(DSRL_ERR_BAD_DTYPE, "seg_counts: unknown pred dtype %d", pred_dtype);
}

template <typename TT, int VEC, int BLOCK>
int launch_logits(const float *logits, const void *target, const uint8_t *mask, int64_t num_updates, int64_t batch,
                  int64_t hw, int nc, int ignore_label, int64_t *counts, int64_t *pred_out, cudaStream_t st) {
    constexpr long long TILE = (long long)BLOCK * kLogitsPxPerThread;
    const int rows = 3 * nc + 3;
    const size_t smem = (size_t)((rows + 3) / 4) * BLOCK * 4 + (size_t)(rows + 4) * sizeof(int);
    const long long tiles_per_image = (hw + TILE - 1) / TILE;
    const long long grid = tiles_per_image * batch * num_updates;
    if (grid > 0x7fffffffLL) DSRL_FAIL(DSRL_ERR_UNSUPPORTED, "seg_counts_from_logits: too many tiles (%lld)", grid);
    if (mask) {
        auto kern = seg_counts_logits_kernel<TT, true, VEC, BLOCK>;
        int rc = prepare_smem(kern, smem);
        if (rc) return rc;
        kern<<<(unsigned)grid, BLOCK, smem, st>>>(logits, static_cast<const TT *>(target), mask, batch, hw, nc, ignore_label,
                                                   (int)tiles_per_image, reinterpret_cast<unsigned long long *>(counts),
                                                   reinterpret_cast<long long *>(pred_out));
    } else {
        auto kern = seg_counts_logits_kernel<TT, false, VEC, BLOCK>;
        int rc = prepare_smem(kern, smem);
        if (rc) return rc;
        kern<<<(unsigned)grid, BLOCK, smem, st>>>(logits, static_cast<const TT *>(target), mask, batch, hw, nc, ignore_label,
                                                   (int)tiles_per_image, reinterpret_cast<unsigned long long *>(counts),
                                                   reinterpret_cast<long long *>(pred_out));
    }
    DSRL_LAUNCH_CHECK();
    return DSRL_OK;
}

template <typename TT>
int dispatch_logits(const float *logits, const void *target, const uint8_t *mask, int64_t num_updates, int64_t batch,
                    int64_t hw, int nc, int ignore_label, int64_t *counts, int64_t *pred_out, cudaStream_t st) {
    const bool vec4 = (hw % 4 == 0) && aligned_to(logits, 16) && aligned_to(target, sizeof(TT) * 4) &&
                      (mask == nullptr || aligned_to(mask, 4));
    if (vec4) return launch_logits<TT, 4, 256>(logits, target, mask, num_updates, batch, hw, nc, ignore_label, counts, pred_out, st);
    return launch_logits<TT, 1, 256>(logits, target, mask, num_updates, batch, hw, nc, ignore_label, counts, pred_out, st);
}

int check_common(const void *a, const void *b, const void *counts, int64_t num_updates, int64_t npix, int nc) {
    if (!a || !b || !counts) DSRL_FAIL(DSRL_ERR_BAD_ARG, "seg_counts: null pointer");
    if (num_updates < 0 || npix < 0) DSRL_FAIL(DSRL_ERR_BAD_SHAPE, "seg_counts: negative size");
    // uint8 labels: `x + 1` wraps 255 to 0 in the reference (mIoU.py:21-22), so class ids above 254 cannot be
    // represented identically; 8-bit shared-memory counters also bound the row count.
    if (nc < 1 || nc > 254) DSRL_FAIL(DSRL_ERR_UNSUPPORTED, "seg_counts: num_classes must be in [1, 254], got %d", nc);
    return DSRL_OK;
}

}  // namespace
}  // namespace dsrl

using namespace dsrl;

extern "C" int dsrl_seg_counts(const void *pred, int pred_dtype, const void *target, int target_dtype, const uint8_t *mask,
                               int64_t num_updates, int64_t npix_per_update, int num_classes, int ignore_label,
                               int64_t *counts, dsrl_stream_t stream) {
    int rc = check_common(pred, target, counts, num_updates, npix_per_update, num_classes);
    if (rc) return rc;
    if ((rc = require_device())) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (num_updates == 0) return DSRL_OK;
    DSRL_CUDA_TRY(cudaMemsetAsync(counts, 0, sizeof(int64_t) * (size_t)num_updates * DSRL_SEG_ROW_LEN(num_classes), st));
    count_launch();
    if (npix_per_update == 0) return DSRL_OK;
    if (num_classes <= 120)
        return dispatch_pred<512>(pred, pred_dtype, target, target_dtype, mask, num_updates, npix_per_update, num_classes, ignore_label, counts, st);
    return dispatch_pred<128>(pred, pred_dtype, target, target_dtype, mask, num_updates, npix_per_update, num_classes, ignore_label, counts, st);
}

extern "C" int dsrl_seg_counts_from_logits(const float *logits, const void *target, int target_dtype, const uint8_t *mask,
                                           int64_t num_updates, int64_t batch, int64_t hw, int num_classes, int ignore_label,
                                           int64_t *counts, int64_t *pred_out, dsrl_stream_t stream) {
    int rc = check_common(logits, target, counts, num_updates, hw, num_classes);
    if (rc) return rc;
    if (batch < 0) DSRL_FAIL(DSRL_ERR_BAD_SHAPE, "seg_counts_from_logits: negative batch");
    if (num_classes > 120) DSRL_FAIL(DSRL_ERR_UNSUPPORTED, "seg_counts_from_logits: num_classes > 120");
    if ((rc = require_device())) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (num_updates == 0) return DSRL_OK;
    DSRL_CUDA_TRY(cudaMemsetAsync(counts, 0, sizeof(int64_t) * (size_t)num_updates * DSRL_SEG_ROW_LEN(num_classes), st));
    count_launch();
    if (batch == 0 || hw == 0) return DSRL_OK;
    switch (target_dtype) {
        case DSRL_U8: return dispatch_logits<uint8_t>(logits, target, mask, num_updates, batch, hw, num_classes, ignore_label, counts, pred_out, st);
        case DSRL_I32: return dispatch_logits<int32_t>(logits, target, mask, num_updates, batch, hw, num_classes, ignore_label, counts, pred_out, st);
        case DSRL_I64: return dispatch_logits<long long>(logits, target, mask, num_updates, batch, hw, num_classes, ignore_label, counts, pred_out, st);
    }
    DSRL_FAIL(DSRL_ERR_BAD_DTYPE, "seg_counts_from_logits: unknown target dtype %d", target_dtype);
}
