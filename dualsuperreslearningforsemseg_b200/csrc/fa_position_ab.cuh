// FA loss, position semantics: the SYMMETRIC TWO-PASS form of the FP16 tile engine (included by fa_position.cu inside
// namespace dsrl { namespace { ... } }; it uses that file's geometry, argument structs, epilogue_finish and expand_signs).
//
// D = S1 - S2 is symmetric, and the gradient needs nothing of D but the SIGN of every entry.  So instead of one fused kernel
// that computes every D tile next to the gradient accumulator it feeds (8 C N^2 executed FLOP per sample, and a cluster of
// four CTAs at C = 256 because accumulator + D tiles do not fit one SM's tensor memory -- only 132 of the 148 SMs can hold
// clusters of four), the work is split into two tensor-bound passes over CTA pairs (clusters of two, every SM busy):
//
//   pass A  fa_pos_dsign   D(i, j) for the tiles j >= i only (2 C N^2): |D| sums (the loss, off-diagonal tiles counted twice),
//                          near-tie lists (exact signs), and the sign / is-zero BITS of the tile written into the "sign
//                          planes", block (i, j).  Four D buffers in tensor memory (all 512 columns) decouple the MMA
//                          issuer from the conversion warps, which are what bounds this pass (issue slots): everything that
//                          can wait is left to pass B.
//   resolve fa_pos_resolve (exact signs) re-decides the listed near ties from the unrounded features and SETS THE BITS,
//                          so no accumulator is ever corrected and no raw rows are kept.
//   pass B  fa_pos_grad    O_i = sum_j sign(D_ij) Fh_j (4 C N^2): per column tile the conversion warps read 4 KB of bits --
//                          block (i, j) for j >= i, block (j, i) bit-TRANSPOSED across the warp (five shuffle butterflies)
//                          for j < i -- and expand them into the packed FP16 sign tile in tensor memory (A-from-TMEM
//                          operand); the issuer runs the gradient MMAs against the V_j boxes (tensor bound: the conversion
//                          warps have the slack pass A lacks); normalisation Jacobian and dX / dP in the epilogue.
//
// Executed tensor work: 6 C N^2 per sample (4 for the gradient, 2 for the upper triangle of D) instead of 8.
// Sign planes: per sample T x T blocks of 4 KB, block (I, J) = [column half h][row r] x 16 bytes {neg s0, zero s0, neg s1,
// zero s1}: the sign and is-zero bits of row r of tile I against the two 32-column strips s of half h of tile J (only I <= J is
// stored, as a packed upper triangle of blocks: the lower triangle is read transposed), bit e of a
// word = entry 2e, bit 16 + e = entry 2e + 1 of the strip (the order expand_signs() wants).  N^2 / 8 bytes per sample
// (134 MB at N = 32768), written once and read twice per channel group; geometries whose planes would exceed the cap
// fall back to the fused kernels (kSignPlaneCapBytes).

constexpr size_t kSignBlock = kSignBlockBytes;       // bytes per (row tile, column tile)
constexpr int kDBufs = 4;                            // pass A: D tiles in flight (4 x 128 tensor-memory columns)
constexpr int kSBufs = 4;                            // pass B: packed sign tiles in flight (4 x 64 columns at kColS)
constexpr uint32_t kColS = 256;
constexpr int kDsignEpiWarps = 16;                   // pass A: one conversion warp per (lane quarter, 32-column strip)
constexpr int kDsignEpiThreads = kDsignEpiWarps * 32;
constexpr int kDsignThreads = 64 + kDsignEpiThreads;
constexpr int kGradConvWarps = 16;                   // pass B: two sets of eight conversion warps (lane quarter x column half) that alternate over the tiles; warps 2..9 also run the epilogue
constexpr int kGradThreads = 64 + kGradConvWarps * 32;
constexpr int kGradFBoxes = 4;                       // pass B: boxes (128 rows x 64 channels) of the CTA's own feature rows kept for the epilogue

// 32 x 32 bit transpose across a warp: lane l enters with row l, leaves with column l (bit b = row b's bit l)
__device__ __forceinline__ uint32_t transpose32(uint32_t x, int lane) {
#pragma unroll
    for (int s = 0; s < 5; ++s) {
        const int k = 16 >> s;
        const uint32_t m = s == 0 ? 0x0000ffffu : s == 1 ? 0x00ff00ffu : s == 2 ? 0x0f0f0f0fu : s == 3 ? 0x33333333u : 0x55555555u;   // bits b with (b & k) == 0
        const uint32_t y = __shfl_xor_sync(0xffffffffu, x, k);
        x = (lane & k) ? ((x & ~m) | ((y >> k) & m)) : ((x & m) | ((y << k) & ~m));
    }
    return x;
}
// natural bit order (bit e = entry e) -> the paired order of the sign planes (bit e = entry 2e, bit 16 + e = entry 2e + 1)
__device__ __forceinline__ uint32_t pair_order(uint32_t x) {
    uint32_t ev = x & 0x55555555u, od = (x >> 1) & 0x55555555u;
    ev = (ev | (ev >> 1)) & 0x33333333u; od = (od | (od >> 1)) & 0x33333333u;
    ev = (ev | (ev >> 2)) & 0x0f0f0f0fu; od = (od | (od >> 2)) & 0x0f0f0f0fu;
    ev = (ev | (ev >> 4)) & 0x00ff00ffu; od = (od | (od >> 4)) & 0x00ff00ffu;
    ev = (ev | (ev >> 8)) & 0x0000ffffu; od = (od | (od >> 8)) & 0x0000ffffu;
    return ev | (od << 16);
}

// Only the blocks I <= J exist: they are stored as the packed upper triangle, row by row.
__host__ __device__ inline size_t sign_blocks(int T) { return (size_t)T * (T + 1) / 2; }                       // per sample
__host__ __device__ inline size_t sign_block_index(int T, int I, int J) { return (size_t)I * (2 * T - I + 1) / 2 + (size_t)(J - I); }

// work unit of pass A: (pair of row tiles p, column chunk) -- column tiles [2p + chunk * Lc, ... + Lc) of the rows [256 p, 256 p + 256)
__host__ __device__ inline int dsign_chunks(int T, int Lc, int p) { return (T - 2 * p + Lc - 1) / Lc; }

// ---------------------------------------------------------------------------------------------------------------
// pass A: upper-triangle D tiles -> loss, near-tie lists, sign planes
// ---------------------------------------------------------------------------------------------------------------
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kDsignThreads, 1)
fa_pos_dsign(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k, const PosGeom g, const PosArgs a) {
    extern __shared__ unsigned char smraw[];
    const uint32_t raw = smem_u32(smraw);
    unsigned char *sm = smraw + (((raw + 1023u) & ~1023u) - raw);
    constexpr int kElems = 64;                                      // FP16 operand elements per 128-byte row
    constexpr int kUnitBytes = kPairKBox;                           // this CTA's 64 rows of one 64-channel chunk of K_j
    constexpr int kStageBytes = 4 * kPairKBox;                      // 32 KB (stages of 16 KB were measured: 2165 -> 2727 cycles per tile --
    constexpr int kUPS = 4;                                         // twice the barrier round trips for the same bytes in flight)
    const int S = g.a_stages, nkc = g.nkh;
    unsigned char *qreg = sm;
    unsigned char *ring = sm + (size_t)nkc * kBoxBytes;
    uint64_t *full = reinterpret_cast<uint64_t *>(ring + (size_t)S * kStageBytes);
    uint64_t *empty = full + S;
    uint64_t *q_full = empty + S;
    uint64_t *d_full = q_full + 1;          // [kDBufs]
    uint64_t *d_empty = d_full + kDBufs;    // [kDBufs] (in the leader: the conversion warps of both CTAs have read the buffer)
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(d_empty + kDBufs);
    double *red = reinterpret_cast<double *>(d_empty + kDBufs + 1);
    int *flag = reinterpret_cast<int *>(red + 2 * kDsignEpiWarps);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int T = g.tiles, Lc = g.a_chunk, b = blockIdx.z;
    int p = (int)(blockIdx.x >> 1), chunk = 0;
    if (Lc < T) {                                                    // units are numbered pair by pair, chunk by chunk
        int rem = p;
        for (p = 0;; ++p) {
            const int nch = dsign_chunks(T, Lc, p);
            if (rem < nch) break;
            rem -= nch;
        }
        chunk = rem;
    }
    const int itile = 2 * p + (int)rank;
    const int jbeg = 2 * p + chunk * Lc, nt = min(Lc, T - jbeg);
    const int row_q = b * g.Npad + itile * kTile;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tm_q);
        prefetch_tmap(&tm_k);
        for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(q_full, 1);
        for (int i = 0; i < kDBufs; ++i) { mbar_init(&d_full[i], 1); mbar_init(&d_empty[i], 2 * kDsignEpiWarps); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc2(tmem_slot, kTmemCols);
    fence_before_sync();
    cluster_sync();
    fence_after_sync();
    const uint32_t tmem = *tmem_slot;

    int slot = 0;
    uint32_t ph = 0;
#define RING_ADVANCE() do { if (++slot == S) { slot = 0; ph ^= 1; } } while (0)

    if (warp == 0) {
        // ===================================== TMA producer (both CTAs, each for its own shared memory) =====================================
        if (elect_one()) {
            if (leader) mbar_arrive_expect_tx(q_full, 2u * (uint32_t)nkc * kBoxBytes);
            for (int kc = 0; kc < nkc; ++kc) tma_load_2d_pair(qreg + (size_t)kc * kBoxBytes, &tm_q, q_full, kc * kElems, row_q);
        }
        __syncwarp();
        for (int jj = 0; jj < nt; ++jj) {
            const int row_k = b * g.Npad + (jbeg + jj) * kTile + (int)rank * (kTile / 2);     // this CTA's 64 rows of K_j
            for (int kc0 = 0; kc0 < nkc; kc0 += kUPS) {
                const int nu = min(kUPS, nkc - kc0);
                mbar_wait(&empty[slot], ph ^ 1, 31);
                if (elect_one()) {
                    unsigned char *dst = ring + (size_t)slot * kStageBytes;
                    if (leader) mbar_arrive_expect_tx(&full[slot], 2u * (uint32_t)(nu * kUnitBytes));
                    for (int u = 0; u < nu; ++u) tma_load_2d_pair(dst + (size_t)u * kUnitBytes, &tm_k, &full[slot], (kc0 + u) * kElems, row_k);
                }
                __syncwarp();
                RING_ADVANCE();
            }
        }
    } else if (warp == 1) {
        // ===================================== MMA issuer (leader CTA only) =====================================
        if (leader) {
            constexpr uint64_t kBoxDesc = kBoxBytes >> 4, kStageDesc = kStageBytes >> 4, kUnitDesc = kUnitBytes >> 4;
            const uint64_t ring_desc = smem_desc_sw128(smem_u32(ring)), q_desc = smem_desc_sw128(smem_u32(qreg));
            const uint32_t id_pos = idesc_f16(2 * kTile, kTile, false), id_neg = idesc_f16(2 * kTile, kTile, true);
            const int q_neg = g.C1p / 16;                           // first K step (16 channels) of branch 2 (subtracted)
            long long w_full = 0, w_de = 0;
            const long long t_begin = clock64();
            mbar_wait(q_full, 0, 32);
            const long long t_q = clock64();
            for (int jj = 0; jj < nt; ++jj) {
                const int buf = jj & (kDBufs - 1);
                const uint32_t dcol = tmem + (uint32_t)buf * kTile;
                TWAIT(w_de, mbar_wait(&d_empty[buf], ((uint32_t)(jj / kDBufs) & 1u) ^ 1u, 33));      // the tile that used this buffer has been read
                fence_after_sync();
                for (int kc0 = 0; kc0 < nkc; kc0 += kUPS) {
                    TWAIT(w_full, mbar_wait(&full[slot], ph, 34));
                    const uint64_t sd = ring_desc + (uint64_t)slot * kStageDesc;
                    const int ss = slot;
                    RING_ADVANCE();
                    fence_after_sync();
                    if (elect_one()) {
#pragma unroll
                        for (int u = 0; u < kUPS; ++u) {
                            const int kc = kc0 + u;
                            if (kc < nkc) {
                                const uint64_t ad = q_desc + (uint64_t)kc * kBoxDesc, bd = sd + (uint64_t)u * kUnitDesc;
#pragma unroll
                                for (int ks = 0; ks < 4; ++ks)
                                    mma_f16_ss_pair(dcol, ad + 2 * ks, bd + 2 * ks, 4 * kc + ks >= q_neg ? id_neg : id_pos, (kc | ks) != 0);
                            }
                        }
                        umma_commit_pair(&empty[ss]);
                        if (kc0 + kUPS >= nkc) umma_commit_pair(&d_full[buf]);
                    }
                    __syncwarp();
                }
            }
#ifdef DSRL_POS_TIMING
            if (blockIdx.x == 0 && blockIdx.z == 0 && lane == 0) {
                long long *tm = reinterpret_cast<long long *>(a.sum_out) + 8;      // saved[64..256): role clocks of CTA (0,0,0)
                tm[0] = t_q - t_begin; tm[1] = clock64() - t_q; tm[2] = w_full; tm[3] = w_de; tm[4] = nt;
            }
#else
            (void)t_begin; (void)t_q; (void)w_full; (void)w_de;
#endif
        }
    } else {
        // ===================================== conversion warps =====================================
        // Sixteen warps, one per (TMEM lane quarter q, 32-column strip cg) of a tile: the conversion is a chain of dependent
        // short instructions, so what bounds it is the latency of one warp's chain per tile, not the issue slots -- eight warps
        // converting two strips each ran pass A at 49 % of the tensor peak (ncu r02c), about 960 instructions per warp and tile.
        const int q = warp & 3, r = q * 32 + lane, cg = (warp - 2) >> 2;
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        const bool listing = g.exact != 0;
        float tau = 0.f;
        if (listing) {
            // the tie threshold of this sample, while the operand rows load: the arithmetic of fa_pos_tau, operation for operation
            // (every CTA of the sample derives the same bits; the fused kernels use that kernel, and the near-tie lists of the two
            // forms are compared entry for entry by the tests)
            const int et = threadIdx.x - 64, M = g.N < 64 ? g.N : 64, step = g.N / M;
            float mu = 0.f;
            if (et < g.Kc) {
                const double *inv = a.inv64 + ((size_t)b * 2 + (et >= g.C1p)) * g.Npad;
                const float *pc = a.Ppm + (size_t)b * g.Npad * g.Kc + et;
                float m0 = 0.f, m1 = 0.f;
#pragma unroll 8
                for (int m = 0; m < M; m += 2) { const int i = m * step; const float f = pc[(size_t)i * g.Kc] * (float)inv[i]; m0 = fmaf(f, f, m0); }
#pragma unroll 8
                for (int m = 1; m < M; m += 2) { const int i = m * step; const float f = pc[(size_t)i * g.Kc] * (float)inv[i]; m1 = fmaf(f, f, m1); }
                mu = (m0 + m1) / (float)M;
            }
            float *redf = reinterpret_cast<float *>(red);
            const float wsum = warp_sum(mu * mu);
            if (lane == 0) redf[et >> 5] = wsum;
            asm volatile("bar.sync 1, %0;" ::"n"(kDsignEpiThreads) : "memory");
            if (et < 32) {
                const float t = warp_sum(lane < kDsignEpiWarps ? redf[lane] : 0.f);
                if (lane == 0) {
                    const float tv = a.tau_ksigma * sqrtf(2.f * a.tau_r2 * t + a.tau_floor2);
                    redf[32] = tv;
                    if (blockIdx.x == 0) a.tau_out[b] = tv;          // for the resolve pass's statistics
                }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(kDsignEpiThreads) : "memory");
            tau = redf[32];
            asm volatile("bar.sync 1, %0;" ::"n"(kDsignEpiThreads) : "memory");      // red is reused by the loss reduction
        }
        const size_t gsub = ((size_t)b * g.Npad + (size_t)itile * kTile + r) * g.fnsub + (size_t)(4 * chunk + cg);
        uint2 *srow = reinterpret_cast<uint2 *>(a.sb + ((size_t)b * sign_blocks(T) + sign_block_index(T, itile, max(jbeg, itile))) * (kSignBlock / 16) + (size_t)(cg >> 1) * kTile + r) + (cg & 1);
        unsigned nlisted = 0;
        double acc = 0.0;
        float facc = 0.f;
        long long w_d = 0;
        const long long t_begin = clock64();
        // tile sums are gathered in FP32 over eight tiles before they enter the FP64 total.  The FP64 add sits in an OUTER loop:
        // written as `if ((jj & 7) == 7) acc += facc` the compiler predicates it and issues a DADD per tile and warp, which
        // was the top stall of this kernel (FP64 pipe, ncu r02d: 13 % of the samples)
        for (int jj0 = 0; jj0 < nt; jj0 += 8) {
        const int jj1 = min(jj0 + 8, nt);
#pragma unroll 1
        for (int jj = jj0; jj < jj1; ++jj) {
            const int buf = jj & (kDBufs - 1), j = jbeg + jj;
            TWAIT(w_d, mbar_wait(&d_full[buf], (uint32_t)(jj / kDBufs) & 1u, 35));
            fence_after_sync();
            const bool diag = j == itile, lower = j < itile;       // lower: tile (2p + 1, 2p), supplied by the transpose of (2p, 2p + 1)
            uint32_t v[32];
            tmem_ld32(tmem + lane_addr + (uint32_t)(buf * kTile + cg * 32), v);
            tmem_ld_wait();
            fence_before_sync();                                   // the strip is in registers: the buffer may be overwritten
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(&d_empty[buf]);
            if (lower) continue;                                   // warp-uniform
            if (diag && cg == q) {                                 // S_ii = 1 in both branches: a structural tie
#pragma unroll
                for (int e = 0; e < 32; ++e) if (e == lane) v[e] = 0u;
            }
            float tsum = 0.f, zmin = 3.0e38f;
#pragma unroll
            for (int e = 0; e < 32; ++e) {
                const float x = __uint_as_float(v[e]);
                tsum += fabsf(x);
                zmin = fminf(zmin, fabsf(x));
            }
            // sign / is-zero words in the paired order (bit e = entry 2e, bit 16 + e = entry 2e + 1): one funnel shift per entry
            // pushes its sign bit into the word of its parity
            uint32_t M, Z = 0u;
            {
                uint32_t ev = 0u, od = 0u;
#pragma unroll
                for (int e = 15; e >= 0; --e) { ev = __funnelshift_l(v[2 * e], ev, 1); od = __funnelshift_l(v[2 * e + 1], od, 1); }
                M = ev | (od << 16);
            }
            if (zmin == 0.f) {                                     // exact zeros (the forced diagonal, padding, identical branches): sign 0
#pragma unroll
                for (int e = 0; e < 16; ++e)
                    Z |= ((v[2 * e] & 0x7fffffffu) ? 0u : (1u << e)) | ((v[2 * e + 1] & 0x7fffffffu) ? 0u : (1u << (16 + e)));
                M &= ~Z;
            }
            if (listing && zmin < tau) {
                // near ties of this strip, upper triangle only (the resolve pass sets both mirror bits); exact zeros keep sign 0.
                // A warp takes this path when ANY of its rows has a near tie in the strip (about 70 % of the strips at 3.5 sigma),
                // so the scan is two instructions per entry: a compare that yields all ones, funnel-shifted into the mask.
                uint32_t m = 0u;
#pragma unroll
                for (int e = 31; e >= 0; --e) {
                    uint32_t c;
                    asm("set.lt.u32.f32 %0, %1, %2;" : "=r"(c) : "f"(fabsf(__uint_as_float(v[e]))), "f"(tau));
                    m = __funnelshift_l(c, m, 1);                  // entry e ends at bit e
                }
                if (zmin == 0.f) {
#pragma unroll
                    for (int e = 0; e < 32; ++e) m &= (v[e] & 0x7fffffffu) ? 0xffffffffu : ~(1u << e);
                }
                if (diag) m = cg > q ? m : (cg == q ? (m & ~((2u << lane) - 1u)) : 0u);       // columns right of the diagonal
                while (m) {
                    const int e = __ffs(m) - 1;
                    m &= m - 1;
                    const uint32_t ng = (M >> ((e >> 1) + ((e & 1) << 4))) & 1u;
                    if (nlisted < (unsigned)g.fsub) a.fent[gsub * g.fsub + nlisted] = (uint32_t)(j * kTile + cg * 32 + e) | (ng << 31);
                    ++nlisted;
                }
            }
            // block (itile, j): this thread's row, its strip's two words (consecutive column tiles are one block apart)
            srow[(size_t)(j - max(jbeg, itile)) * (kSignBlock / 8)] = make_uint2(M, Z);
            facc += diag ? tsum : 2.f * tsum;
        }
        acc += (double)facc;
        facc = 0.f;
        }
        if (listing) a.fcnt[gsub] = nlisted;
#ifdef DSRL_POS_TIMING
        if (blockIdx.x == 0 && blockIdx.z == 0 && threadIdx.x == 64) {
            long long *tm = reinterpret_cast<long long *>(a.sum_out) + 8;
            tm[5] = clock64() - t_begin; tm[6] = w_d;
        }
#else
        (void)t_begin; (void)w_d;
#endif
        // loss: per-CTA partial, finished in a fixed order by the last CTA to arrive (deterministic)
        {
            const int et = threadIdx.x - 64;
            const int part_index = (int)(blockIdx.z * gridDim.x + blockIdx.x), nparts = (int)(gridDim.x * gridDim.z);
            const double tot = warp_sum(acc);
            if (lane == 0) red[et >> 5] = tot;
            asm volatile("bar.sync 1, %0;" ::"n"(kDsignEpiThreads) : "memory");
            if (et == 0) {
                double sum = 0.0;
                for (int i = 0; i < kDsignEpiWarps; ++i) sum += red[i];
                a.partials[part_index] = sum;
                __threadfence();
                *flag = atomicInc(a.ticket, (unsigned)nparts - 1u) == (unsigned)nparts - 1u;      // self-resetting
            }
            asm volatile("bar.sync 1, %0;" ::"n"(kDsignEpiThreads) : "memory");
            if (*flag) {
                __threadfence();
                double sum = 0.0;
                for (int i = et; i < nparts; i += kDsignEpiThreads) sum += __ldcg(a.partials + i);
                sum = warp_sum(sum);
                if (lane == 0) red[kDsignEpiWarps + (et >> 5)] = sum;
                asm volatile("bar.sync 1, %0;" ::"n"(kDsignEpiThreads) : "memory");
                if (et == 0) {
                    double all = 0.0;
                    for (int i = 0; i < kDsignEpiWarps; ++i) all += red[kDsignEpiWarps + i];
                    *a.sum_out = all;
                    *a.loss_out = (float)(all / a.loss_div);
                }
            }
        }
    }
#undef RING_ADVANCE

    fence_before_sync();
    cluster_sync();                         // the leader's MMAs read the peer's shared / tensor memory until the very end
    if (warp == 1) tmem_dealloc2(tmem, kTmemCols);
}

// Epilogue of pass B when the CTA's feature rows are in shared memory: normalisation Jacobian of the accumulator rows,
//   dF_i = (O_i - Fh_i <Fh_i, O_i>) * (+-2 / Z) / ||F_i||    per branch,
// by all sixteen conversion warps (TMEM lane quarter q x channel part: 32-channel chunks part, part + 4, ...), the features
// read with conflict-free 16-byte shared-memory loads (a quarter warp covers the eight swizzled chunks of a 128-byte row).
// Stored channel-major -- dP (padded) or, fused forward + backward without pooling, dX itself scaled by *go -- coalesced
// along positions.  Same arithmetic, same bits as epilogue_finish.
constexpr int kGradParts = 4;
__device__ __forceinline__ void grad_epilogue_smem(const PosGeom &g, const PosArgs &a, const unsigned char *fsm, float *projbuf /* [parts][2][128] */,
                                                   uint32_t tmem, int itile, int b, int gN, int gbeg) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = warp & 3, r = q * 32 + lane, part = (warp - 2) >> 2;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const int row = itile * kTile + r;
    auto feat8 = [&](int cl, float (&f)[8]) {                      // channels gbeg + cl .. + 7 (cl a multiple of 8) of this thread's row
        const uint4 w = *reinterpret_cast<const uint4 *>(fsm + (size_t)(cl >> 6) * kBoxBytes + r * 128 + ((((cl & 63) >> 3) ^ (r & 7)) << 4));
        const float2 f0 = __half22float2(*reinterpret_cast<const __half2 *>(&w.x)), f1 = __half22float2(*reinterpret_cast<const __half2 *>(&w.y));
        const float2 f2 = __half22float2(*reinterpret_cast<const __half2 *>(&w.z)), f3 = __half22float2(*reinterpret_cast<const __half2 *>(&w.w));
        f[0] = f0.x; f[1] = f0.y; f[2] = f1.x; f[3] = f1.y; f[4] = f2.x; f[5] = f2.y; f[6] = f3.x; f[7] = f3.y;
    };
    float proj[2] = {0.f, 0.f};
    for (int c0 = 32 * part; c0 < gN; c0 += 32 * kGradParts) {
        uint32_t v[32];
        tmem_ld32(tmem + lane_addr + (uint32_t)c0, v);
        tmem_ld_wait();
        float p = 0.f;
#pragma unroll
        for (int e8 = 0; e8 < 4; ++e8) {
            float f[8];
            feat8(c0 + 8 * e8, f);
#pragma unroll
            for (int e = 0; e < 8; ++e) p = fmaf(f[e], __uint_as_float(v[e8 * 8 + e]), p);
        }
        if (gbeg + c0 >= g.C1p) proj[1] += p; else proj[0] += p;
    }
    projbuf[(part * 2 + 0) * kTile + r] = proj[0];
    projbuf[(part * 2 + 1) * kTile + r] = proj[1];
    asm volatile("bar.sync 1, %0;" ::"n"(kGradConvWarps * 32) : "memory");
    float scale[2];
#pragma unroll
    for (int br = 0; br < 2; ++br) {
        float sp = 0.f;
#pragma unroll
        for (int i = 0; i < kGradParts; ++i) sp += projbuf[(i * 2 + br) * kTile + r];
        const float n = a.nrm[((size_t)b * 2 + br) * g.Npad + row];
        proj[br] = n > 1e-12f ? sp : 0.f;                           // F / eps branch of the clamp: no projection
        scale[br] = (br ? -a.grad_scale : a.grad_scale) / fmaxf(n, 1e-12f);
    }
    const float gmul = a.direct ? __ldg(a.go) : 1.f;               // applied as a second multiply: the bits fa_pos_unpool would produce
    for (int c0 = 32 * part; c0 < gN; c0 += 32 * kGradParts) {
        uint32_t v[32];
        tmem_ld32(tmem + lane_addr + (uint32_t)c0, v);
        tmem_ld_wait();
        const int c = gbeg + c0, br = c >= g.C1p, cb = br ? g.C1p : 0, Cr = br ? g.C2 : g.C1;
        const size_t pitch = a.direct ? (size_t)g.N : (size_t)g.Npad;
        float *dst = a.direct ? a.dx[br] + ((size_t)b * Cr + (c - cb)) * pitch + row : a.dP + ((size_t)b * g.Kc + c) * pitch + row;
        const int nreal = a.direct ? (row < g.N ? Cr - (c - cb) : 0) : 32;       // channels of this chunk that exist in dX
        const float pr = proj[br], sc = scale[br];
#pragma unroll
        for (int e8 = 0; e8 < 4; ++e8) {
            float f[8];
            feat8(c0 + 8 * e8, f);
#pragma unroll
            for (int e = 0; e < 8; ++e)
                if (e8 * 8 + e < nreal) dst[(size_t)(e8 * 8 + e) * pitch] = (__uint_as_float(v[e8 * 8 + e]) - f[e] * pr) * sc * gmul;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// pass B: sign planes -> gradient contraction -> normalisation Jacobian -> dX / dP (or raw partial rows, jsplit > 1)
// ---------------------------------------------------------------------------------------------------------------
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGradThreads, 1)
fa_pos_grad(const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_q, const PosGeom g, const PosArgs a) {
    extern __shared__ unsigned char smraw[];
    const uint32_t raw = smem_u32(smraw);
    unsigned char *sm = smraw + (((raw + 1023u) & ~1023u) - raw);
    constexpr int kElems = 64;
    constexpr int kStageBytes = kBoxBytes;                          // one V box: this CTA's half of the group's channels x 64 positions
    const int S = g.b_stages;
    unsigned char *ring = sm;
    unsigned char *fbuf = ring + (size_t)S * kStageBytes;           // kGradFBoxes boxes: the CTA's own feature rows (epilogue operand)
    uint64_t *full = reinterpret_cast<uint64_t *>(fbuf + (size_t)kGradFBoxes * kBoxBytes);
    uint64_t *empty = full + S;
    uint64_t *p_full = empty + S;           // [kSBufs] (in the leader: both CTAs' conversion warps have written the sign tile)
    uint64_t *p_empty = p_full + kSBufs;    // [kSBufs] the gradient MMAs of the tile that used this buffer have completed
    uint64_t *o_full = p_empty + kSBufs;
    uint64_t *f_full = o_full + 1;          // this CTA's own feature rows have landed in fbuf
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(o_full + 2);
    double *red = reinterpret_cast<double *>(o_full + 3);
    int *flag = reinterpret_cast<int *>(red + 2 * kEpiWarps);
    float *projbuf = reinterpret_cast<float *>(flag + 2);           // [kGradParts][2][128]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    // grid.x = clusters in the order (pair of row tiles, channel group): the channel groups of a pair run on neighbouring SM pairs
    // at the same time and read the same sign blocks, which then come from DRAM once per pair instead of once per group
    const int T = g.tiles, cl = (int)(blockIdx.x >> 1), grp = cl % g.G, xlin = 2 * (cl / g.G) + (int)(blockIdx.x & 1);
    const int js = xlin / T, itile = xlin - js * T, b = blockIdx.z;                    // T is even: the pair shares js
    const int nt = T / g.jsplit, j0 = js * nt;
    const int gN = g.gcnt[grp], gbeg = g.gbeg[grp], vrows = gN / 2;
    const uint32_t vbytes = (uint32_t)vrows * 128u;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tm_v);
        for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int i = 0; i < kSBufs; ++i) { mbar_init(&p_full[i], 2 * kEpiWarps); mbar_init(&p_empty[i], 1); }          // the eight warps of the tile's set, both CTAs
        mbar_init(o_full, 1);
        mbar_init(f_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc2(tmem_slot, kTmemCols);
    fence_before_sync();
    cluster_sync();
    fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    // the epilogue multiplies the accumulator rows with the CTA's own normalised feature rows.  Read per thread from global
    // memory that is one 128-byte line per lane and load (a row is 1 KB): the L1 wavefronts of those loads made the epilogue
    // 9 % of a CTA's life (role clocks r02g: 8.6 k + 14.5 k cycles for the two passes).  The rows are instead fetched by TMA at
    // the start of the kernel: gN / 64 boxes of 128 rows x 64 channels, 128-byte swizzle.
    const int fboxes = gN / 64;
    const bool f_smem = !g.raw_o && gN % 64 == 0 && fboxes <= kGradFBoxes;

    int slot = 0;
    uint32_t ph = 0;
#define RING_ADVANCE() do { if (++slot == S) { slot = 0; ph ^= 1; } } while (0)

    if (warp == 0) {
        // ===================================== TMA producer: the V_j boxes =====================================
        const int row_v = b * g.Kc + gbeg + (int)rank * vrows;                        // this CTA's half of the group's channels
        if (f_smem) {
            if (elect_one()) {
                mbar_arrive_expect_tx(f_full, (uint32_t)fboxes * kBoxBytes);
                for (int i = 0; i < fboxes; ++i)
                    tma_load_2d(fbuf + (size_t)i * kBoxBytes, &tm_q, f_full, gbeg + i * kElems, b * g.Npad + itile * kTile);
            }
            __syncwarp();
        }
        for (int jj = 0; jj < nt; ++jj) {
            for (int jc = 0; jc < 2; ++jc) {
                mbar_wait(&empty[slot], ph ^ 1, 41);
                if (elect_one()) {
                    if (leader) mbar_arrive_expect_tx(&full[slot], 2u * vbytes);
                    tma_load_2d_pair(ring + (size_t)slot * kStageBytes, &tm_v, &full[slot], (j0 + jj) * kTile + jc * kElems, row_v);
                }
                __syncwarp();
                RING_ADVANCE();
            }
        }
    } else if (warp == 1) {
        // ===================================== MMA issuer (leader CTA only) =====================================
        if (leader) {
            constexpr uint64_t kStageDesc = kStageBytes >> 4;
            const uint64_t ring_desc = smem_desc_sw128(smem_u32(ring));
            const uint32_t id_g = idesc_f16(2 * kTile, gN, false);
            long long w_full = 0, w_p = 0, t_first = 0;
            const long long t_begin = clock64();
            for (int jj = 0; jj < nt; ++jj) {
                const int buf = jj & (kSBufs - 1);
                const uint32_t pcol = tmem + kColS + (uint32_t)buf * 64u;
                TWAIT(w_p, mbar_wait(&p_full[buf], (uint32_t)(jj / kSBufs) & 1u, 42));
                fence_after_sync();
                for (int jc = 0; jc < 2; ++jc) {
                    TWAIT(w_full, mbar_wait(&full[slot], ph, 43));
                    if (jj == 0 && jc == 0) t_first = clock64();
                    const uint64_t sd = ring_desc + (uint64_t)slot * kStageDesc;
                    const int ss = slot;
                    RING_ADVANCE();
                    fence_after_sync();
                    if (elect_one()) {
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks)
                            mma_f16_ts_pair(tmem, pcol + (uint32_t)(jc * 32 + ks * 8), sd + 2 * ks, id_g, (jj | jc | ks) != 0);
                        umma_commit_pair(&empty[ss]);
                        if (jc == 1) umma_commit_pair(&p_empty[buf]);
                        if (jc == 1 && jj == nt - 1) umma_commit_pair(o_full);
                    }
                    __syncwarp();
                }
            }
#ifdef DSRL_POS_TIMING
            if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0) {
                long long *tm = reinterpret_cast<long long *>(a.sum_out) + 8;
                tm[12] = t_first - t_begin; tm[13] = clock64() - t_first; tm[14] = w_full; tm[15] = w_p; tm[16] = nt;
            }
            if (blockIdx.x == gridDim.x - 2 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0) {     // a row tile whose column tiles are all below the diagonal
                long long *tm = reinterpret_cast<long long *>(a.sum_out) + 8;
                tm[8] = clock64() - t_first; tm[9] = w_full; tm[10] = w_p;
            }
#else
            (void)t_begin; (void)t_first; (void)w_full; (void)w_p;
#endif
        }
    } else {
        // ===================================== conversion warps: bits -> packed FP16 sign tiles =====================================
        // Two sets of eight warps take the column tiles in turn (set s: tiles with jj % 2 == s), so a warp has two tiles of MMA
        // time for one conversion -- with the transposed read of the lower triangle a single set fell behind the tensor pipe on
        // the row tiles that are mostly below the diagonal.  Inside a set: one warp per (TMEM lane quarter q, column half).
        //   column tiles j >= itile: this thread's row of block (itile, j), the 16 bytes of its half;
        //   column tiles j <  itile: only the mirror block (j, itile) exists -- for each strip cg of its half, lane l loads the
        //     words of ITS ROW c = 32 cg + l there (bits = this warp's 32 rows: strip q % 2 of half q / 2), the warp transposes
        //     them (transpose32) and each lane picks up the word of its own row: bits over the 32 columns of the strip.
        const int cw = warp - 2, set = cw >> 3;
        const int q = warp & 3, r = q * 32 + lane, half = (cw >> 2) & 1;
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        const uint4 *blocks = a.sb + (size_t)b * sign_blocks(T) * (kSignBlock / 16);
        const uint4 *up = blocks + sign_block_index(T, itile, itile) * (kSignBlock / 16) + (size_t)half * kTile + r;         // + (j - itile) blocks
        const uint2 *lo = reinterpret_cast<const uint2 *>(blocks + (size_t)(q >> 1) * kTile + 64 * half + lane) + (q & 1);      // + block (j, itile); strip 1: 32 units on
        auto bits_of = [&](int j) {
            if (j >= itile) return ldg_stream_u4(up + (size_t)(j - itile) * (kSignBlock / 16));
            const uint2 *pl = lo + sign_block_index(T, j, itile) * (kSignBlock / 8);
            const uint2 s0 = ldg_stream_u2(pl), s1 = ldg_stream_u2(pl + 64);
            return make_uint4(s0.x, s0.y, s1.x, s1.y);
        };
        const int src_lane = (lane >> 1) + ((lane & 1) << 4);       // the lane that holds this lane's row after the transpose (paired order)
        constexpr int kAhead = 2;                                      // this set's tiles in flight: 4 tiles of MMA time cover an L2 / HBM round trip
        uint4 pre[kAhead];
#pragma unroll
        for (int i = 0; i < kAhead; ++i) pre[i] = set + 2 * i < nt ? bits_of(j0 + set + 2 * i) : make_uint4(0u, 0u, 0u, 0u);
        long long w_pe = 0;
        const long long t_begin = clock64();
        for (int jj = set; jj < nt; jj += 2) {
            const int buf = jj & (kSBufs - 1), j = j0 + jj;
            uint4 cur = pre[0];
#pragma unroll
            for (int i = 0; i + 1 < kAhead; ++i) pre[i] = pre[i + 1];
            if (jj + 2 * kAhead < nt) pre[kAhead - 1] = bits_of(j + 2 * kAhead);
            if (j < itile) {                                         // warp-uniform
                cur.x = pair_order(__shfl_sync(0xffffffffu, transpose32(cur.x, lane), src_lane));
                cur.z = pair_order(__shfl_sync(0xffffffffu, transpose32(cur.z, lane), src_lane));
                if (__any_sync(0xffffffffu, (cur.y | cur.w) != 0u)) {
                    cur.y = pair_order(__shfl_sync(0xffffffffu, transpose32(cur.y, lane), src_lane));
                    cur.w = pair_order(__shfl_sync(0xffffffffu, transpose32(cur.w, lane), src_lane));
                }
            }
            TWAIT(w_pe, mbar_wait(&p_empty[buf], ((uint32_t)(jj / kSBufs) & 1u) ^ 1u, 44));
            fence_after_sync();
            uint32_t pk[16];
            expand_signs(cur.x, cur.y, pk);
            tmem_st16(tmem + lane_addr + kColS + (uint32_t)(buf * 64 + half * 32), pk);
            expand_signs(cur.z, cur.w, pk);
            tmem_st16(tmem + lane_addr + kColS + (uint32_t)(buf * 64 + half * 32 + 16), pk);
            tmem_st_wait();
            fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(&p_full[buf]);
        }
#ifdef DSRL_POS_TIMING
        const long long t_conv = clock64();
#endif
        if (f_smem) {
            // every conversion warp: accumulator complete, feature rows landed
            mbar_wait(o_full, 0, 47);
            fence_after_sync();
            mbar_wait(f_full, 0, 46);
            grad_epilogue_smem(g, a, fbuf, projbuf, tmem, itile, b, gN, gbeg);
        } else if (warp < 2 + kEpiWarps) {
            EpiCtx c;
            c.d_full = nullptr; c.p_full = p_full; c.o_full = o_full; c.red = red; c.flag = flag; c.proj = projbuf; c.tmem = tmem;
            c.itile = itile; c.js = js; c.grp = grp; c.b = b; c.j0 = j0; c.nt = nt; c.gN = gN; c.gbeg = gbeg; c.T = T;
            c.part_index = -1;                   // the loss was finished by pass A
            c.nparts = 0;
            c.sub = 0; c.nsub = g.fnsub;
            epilogue_finish<true, true>(g, a, c, 0.0);
        }
#ifdef DSRL_POS_TIMING
        if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 64) {
            long long *tm = reinterpret_cast<long long *>(a.sum_out) + 8;
            tm[17] = t_conv - t_begin; tm[18] = w_pe; tm[19] = clock64() - t_conv;
        }
#endif
        (void)t_begin; (void)w_pe;
    }
#undef RING_ADVANCE

    fence_before_sync();
    cluster_sync();
    if (warp == 1) tmem_dealloc2(tmem, kTmemCols);
}

// One entry of the sign planes set to an exactly known sign (resolve pass, one lane): entry (i, j) of a sample whose planes
// start at `sbw`; the caller passes entries of existing blocks only (tile of i <= tile of j).  Integer atomics on different
// bits commute, so the planes are bit-repeatable.
__device__ __forceinline__ void sign_plane_set(uint32_t *sbw, int T, int i, int j, int s) {
    const size_t unit = (sign_block_index(T, i >> 7, j >> 7) * 2 + ((j >> 6) & 1)) * kTile + (i & 127);
    uint32_t *w = sbw + unit * 4 + 2 * ((j >> 5) & 1);
    const uint32_t bit = 1u << (((j & 31) >> 1) + 16 * (j & 1));
    if (s < 0) atomicOr(w, bit); else atomicAnd(w, ~bit);
    if (s == 0) atomicOr(w + 1, bit);
}
