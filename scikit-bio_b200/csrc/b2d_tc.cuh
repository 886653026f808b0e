// Unweighted UniFrac, the DENSE node rows on the 5th-generation tensor cores (tcgen05, int8).
//
// Over the rows that stay dense (nodes with many tips below), with 61-bit fixed-point lengths,
//   sum_q len_q [u_qi xor u_qj] = L_i + L_j - 2 M(i,j),   M(i,j) = sum_q len_q u_qi u_qj,
// and M IS a contraction: M = sum_l 2^(8l) (A B_l^T)(i,j) with A[i][q] = u_qi (0/1 bytes) and
// B_l[j][q] = u_qj * byte l of len_q - unsigned 8-bit operands, int32 accumulators, exact (a sum of
// at most 2^23 products below 2^8).  The accuracy objection to the AND form (it
// cancels for near-identical samples) does not apply once everything is an integer: L_i + L_j - 2M
// is exact, identical samples give exactly 0, and the near-duplicate guard of b2d_tips.cuh covers
// the rounding of the lengths themselves.
//
// One CTA = a 128 x 64 half tile of pairs.  The eight length limbs are stacked along N
// (n = limb * 64 + column sample), so a 128-row k-block costs two N = 256 instructions per 32-row
// k-step into ONE 128-lane x 512-column accumulator - all of tensor memory.
//   * warps 0..15 expand: bits -> operand bytes, written straight into the K-major no-swizzle
//     canonical layout of the instruction (8-row x 16-byte core matrices), two 80 KB stages;
//     fence.proxy.async, then an mbarrier arrive per warp.
//   * warp 16, one lane, issues tcgen05.mma.kind::i8 (M 128, N 256, K 32) and commits to the
//     stage's "empty" barrier (the operands may be overwritten once the instructions have read
//     them) and finally to the "accumulator complete" barrier.
//   * epilogue: every expansion warp reads its TMEM lane quadrant (tcgen05.ld 32x32b), recombines
//     the limbs into 64-bit sums and emits (L_i + L_j - 2M) * 2^(e-61).
// This is an experiment against the look-up-table kernel (k_pair_unweighted_lut, bound by 16
// shared-memory gathers per clock = 128 terms/clk/SM); the int8 peak is 8 192 MAC/clk/SM = 1 024
// terms/clk/SM with eight limbs.  Kept behind the "udense_tc" option until it wins everywhere.
#pragma once

#include "b2d_common.cuh"
#include "b2d_pair.cuh"

namespace b2d {

constexpr int TC_KB = 128;                       // dense rows per k-block (one bit stripe)
constexpr int TC_NCOL = 64;                      // column samples per CTA
constexpr int TC_LIMBS = 8;
constexpr int TC_N = TC_LIMBS * TC_NCOL;         // 512 stacked columns
constexpr int TC_XWARPS = 16;                    // expansion (and epilogue) warps
constexpr int TC_THREADS = (TC_XWARPS + 1) * 32; // + the issuing warp
constexpr int TC_STAGES = 2;
constexpr int TC_A_BYTES = TILE * TC_KB;         // 16 KB
constexpr int TC_B_BYTES = TC_N * TC_KB;         // 64 KB
constexpr int TC_STAGE_BYTES = TC_A_BYTES + TC_B_BYTES;
constexpr size_t TC_SMEM = (size_t)TC_STAGES * TC_STAGE_BYTES + 1024 /* limb table */ + 64 /* barriers, tmem address */;

// ---- tcgen05 PTX ---------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// shared-memory matrix descriptor, K-major, no swizzle: 8-row x 16-byte core matrices; lbo = byte
// distance between the two 16-byte K halves of an instruction, sbo = between 8-row groups
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) | ((uint64_t)1 << 46);   // version 1 (Blackwell), layout 0
}
// instruction descriptor: D = s32, A = B = unsigned 8 bit, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
__device__ __forceinline__ constexpr uint32_t tc_idesc_u8(int m, int n) {
    return (2u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void tc_mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 16 consecutive 32-bit columns of this warp's lane quadrant
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// four presence bits -> four bytes of 0x00 / 0xff
__device__ __forceinline__ uint32_t nibble_mask(uint32_t nib) {
    return (((nib & 0xfu) * 0x00204081u) & 0x01010101u) * 0xffu;
}

// ---- preparation ---------------------------------------------------------------------------------
// scale of the dense part: 2^e >= sum of all dense lengths >= any sample's L
__global__ void __launch_bounds__(256)
k_udense_scale(const double* __restrict__ lend, int64_t nd_pad, double* __restrict__ scale) {
    __shared__ double sh[256];
    double s = 0.0;
    for (int64_t r = threadIdx.x; r < nd_pad; r += 256) s += lend[r];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o; o >>= 1) {
        if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        int e = 61;
        if (sh[0] > 0.0) frexp(sh[0], &e);
        scale[0] = ldexp(1.0, 61 - e);
        scale[1] = ldexp(1.0, e - 61);
    }
}
__global__ void k_udense_lenfx(const double* __restrict__ lend, const double* __restrict__ scale, int64_t nd_pad,
                               long long* __restrict__ lenfx) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < nd_pad) lenfx[r] = __double2ll_rn(lend[r] * scale[0]);
}
// per sample: L (fixed point, exact), the same in double, and the number of dense rows present
__global__ void __launch_bounds__(128)
k_udense_sums(const uint32_t* __restrict__ bitsd, const long long* __restrict__ lenfx, int64_t NSd,
              long long* __restrict__ Lfx, double* __restrict__ Ld, unsigned* __restrict__ Kd) {
    const int64_t b = blockIdx.x;
    const int s = threadIdx.x;
    long long acc = 0;
    unsigned cnt = 0;
    for (int64_t st = 0; st < NSd; ++st) {
        const uint4 w = *reinterpret_cast<const uint4*>(bitsd + ((b * NSd + st) * TILE + s) * 4);
        const uint32_t x[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t m = x[k];
            cnt += __popc(m);
            while (m) {
                const int bit = __ffs(m) - 1;
                m &= m - 1u;
                acc += lenfx[st * 128 + k * 32 + bit];
            }
        }
    }
    Lfx[b * TILE + s] = acc;
    Ld[b * TILE + s] = (double)acc;   // in units of the dense scale
    Kd[b * TILE + s] = cnt;
}

// ---- the tile kernel -----------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS, 1)
k_pair_udense_tc(const uint32_t* __restrict__ bitsd, const long long* __restrict__ lenfx, int64_t NSd,
                 const long long* __restrict__ Lfx, const double* __restrict__ scale, PairArgs a) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* stage0 = smem_raw;
    unsigned char* limb_tab = smem_raw + (size_t)TC_STAGES * TC_STAGE_BYTES;   // [8 limbs][128 rows] of the k-block in flight... per stage below
    uint64_t* bars = reinterpret_cast<uint64_t*>(limb_tab + 1024);
    uint64_t* full = bars;                  // [TC_STAGES] operands written (TC_XWARPS arrivals)
    uint64_t* empty = bars + TC_STAGES;     // [TC_STAGES] operands consumed (tcgen05.commit)
    uint64_t* done = bars + 2 * TC_STAGES;  // accumulator complete
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * TC_STAGES + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tile = blockIdx.x >> 1, half = blockIdx.x & 1;
    const int bi = a.tiles[2 * tile], bj = a.tiles[2 * tile + 1];
    const int nkb = (int)NSd;   // k-blocks = packed stripes

    if (tid == 0) {
        for (int s = 0; s < TC_STAGES; ++s) {
            mbar_init(&full[s], TC_XWARPS);
            mbar_init(&empty[s], 1);
        }
        mbar_init(done, 1);
        mbar_fence_init();
    }
    if (warp == TC_XWARPS) {   // the issuing warp owns the tensor-memory allocation: all 512 columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == TC_XWARPS) {
        // ------------------------------- MMA issuer -------------------------------
        if (lane == 0) {
            constexpr uint32_t IDESC = tc_idesc_u8(TILE, 256);
            int s = 0;
            uint32_t ph = 0;
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(&full[s], ph);
                tc_fence_after();
                const uint32_t abase = smem_u32(stage0 + (size_t)s * TC_STAGE_BYTES);
                const uint32_t bbase = abase + TC_A_BYTES;
#pragma unroll
                for (int ks = 0; ks < TC_KB / 32; ++ks) {
                    // operand layout per 32-byte k-step: [16-byte half][8-row group][8 rows][16 bytes]
                    const uint64_t adesc = tc_smem_desc(abase + ks * (2 * TILE * 16), TILE * 16, 128);
                    const uint64_t b0 = tc_smem_desc(bbase + ks * (2 * TC_N * 16), TC_N * 16, 128);
                    const uint64_t b1 = tc_smem_desc(bbase + ks * (2 * TC_N * 16) + 256 * 16, TC_N * 16, 128);
                    const uint32_t acc = (kb | ks) ? 1u : 0u;
                    tc_mma_i8(tmem, adesc, b0, IDESC, acc);          // limbs 0..3 -> columns   0..255
                    tc_mma_i8(tmem + 256, adesc, b1, IDESC, acc);    // limbs 4..7 -> columns 256..511
                }
                tc_commit(&empty[s]);   // arrives when the instructions above have read their operands
                if (++s == TC_STAGES) {
                    s = 0;
                    ph ^= 1u;
                }
            }
            tc_commit(done);
        }
    } else {
        // ------------------------------- expansion -------------------------------
        // A: 128 row samples x 128 k bytes (0 / 1); B: (8 limbs x 64 column samples) x 128 k bytes.
        // Thread t of 512: A - sample t / 4, k bytes 32 * (t % 4) .. + 31; B - column sample t / 8,
        // k bytes 16 * (t % 8) .. + 15, all eight limbs.
        const uint32_t* Ag = bitsd + ((int64_t)bi * NSd * TILE) * 4;
        const uint32_t* Bg = bitsd + ((int64_t)bj * NSd * TILE + half * TC_NCOL) * 4;
        const int ai = tid >> 2, aw = tid & 3;      // A: sample, 32-bit word of the stripe
        const int bjx = tid >> 3, bh = tid & 7;     // B: column sample, 16-bit piece of the stripe
        int s = 0;
        uint32_t ph = 0;
        for (int kb = 0; kb < nkb; ++kb) {
            const uint32_t wa = Ag[((int64_t)kb * TILE + ai) * 4 + aw];
            const uint32_t wb = (Bg[((int64_t)kb * TILE + bjx) * 4 + (bh >> 1)] >> (16 * (bh & 1))) & 0xffffu;
            // lengths of the 16 rows this thread expands for B: bytes l of lenfx[kb * 128 + 16 bh + r]
            const long long* lf = lenfx + (int64_t)kb * TC_KB + 16 * bh;
            if (kb >= TC_STAGES) mbar_wait(&empty[s], ph ^ 1u);
            unsigned char* A = stage0 + (size_t)s * TC_STAGE_BYTES;
            unsigned char* B = A + TC_A_BYTES;
            // A rows: k bytes 32 aw .. 32 aw + 31 = k-step aw, both halves
            {
                unsigned char* dst = A + aw * (2 * TILE * 16) + (ai >> 3) * 128 + (ai & 7) * 16;
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const uint32_t bits16 = (wa >> (16 * hh)) & 0xffffu;
                    uint4 v;
                    v.x = nibble_mask(bits16) & 0x01010101u;
                    v.y = nibble_mask(bits16 >> 4) & 0x01010101u;
                    v.z = nibble_mask(bits16 >> 8) & 0x01010101u;
                    v.w = nibble_mask(bits16 >> 12) & 0x01010101u;
                    *reinterpret_cast<uint4*>(dst + hh * (TILE * 16)) = v;
                }
            }
            // B rows: k bytes 16 bh .. + 15 = k-step bh / 2, half bh % 2; limb l of the 16 lengths
            {
                const uint32_t m0 = nibble_mask(wb), m1 = nibble_mask(wb >> 4), m2 = nibble_mask(wb >> 8),
                               m3 = nibble_mask(wb >> 12);
                // transpose 16 lengths x 8 bytes into 8 limbs x 16 bytes: 4 x 4 byte transposes (PRMT)
                uint32_t limb[TC_LIMBS][4];
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    uint32_t lo[4], hi[4];
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const unsigned long long v = (unsigned long long)lf[4 * q4 + r];
                        lo[r] = (uint32_t)v;
                        hi[r] = (uint32_t)(v >> 32);
                    }
                    uint32_t t0 = __byte_perm(lo[0], lo[1], 0x5140), t1 = __byte_perm(lo[0], lo[1], 0x7362);
                    uint32_t u0 = __byte_perm(lo[2], lo[3], 0x5140), u1 = __byte_perm(lo[2], lo[3], 0x7362);
                    limb[0][q4] = __byte_perm(t0, u0, 0x5410);
                    limb[1][q4] = __byte_perm(t0, u0, 0x7632);
                    limb[2][q4] = __byte_perm(t1, u1, 0x5410);
                    limb[3][q4] = __byte_perm(t1, u1, 0x7632);
                    t0 = __byte_perm(hi[0], hi[1], 0x5140), t1 = __byte_perm(hi[0], hi[1], 0x7362);
                    u0 = __byte_perm(hi[2], hi[3], 0x5140), u1 = __byte_perm(hi[2], hi[3], 0x7362);
                    limb[4][q4] = __byte_perm(t0, u0, 0x5410);
                    limb[5][q4] = __byte_perm(t0, u0, 0x7632);
                    limb[6][q4] = __byte_perm(t1, u1, 0x5410);
                    limb[7][q4] = __byte_perm(t1, u1, 0x7632);
                }
                unsigned char* dstb = B + (bh >> 1) * (2 * TC_N * 16) + (bh & 1) * (TC_N * 16);
#pragma unroll
                for (int l = 0; l < TC_LIMBS; ++l) {
                    const int n = l * TC_NCOL + bjx;
                    uint4 v;
                    v.x = limb[l][0] & m0;
                    v.y = limb[l][1] & m1;
                    v.z = limb[l][2] & m2;
                    v.w = limb[l][3] & m3;
                    *reinterpret_cast<uint4*>(dstb + (n >> 3) * 128 + (n & 7) * 16) = v;
                }
            }
            fence_proxy_async_smem();   // generic-proxy stores -> visible to the tensor core's async proxy
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[s]);
            if (++s == TC_STAGES) {
                s = 0;
                ph ^= 1u;
            }
        }
        // ------------------------------- epilogue -------------------------------
        mbar_wait(done, 0);
        tc_fence_after();
        const int quad = warp & 3, part = warp >> 2;   // TMEM lanes 32 quad .. + 31; column samples 16 part .. + 15
        const int row = 32 * quad + lane;
        unsigned long long m[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) m[c] = 0ull;
#pragma unroll
        for (int l = 0; l < TC_LIMBS; ++l) {
            uint32_t v[16];
            tc_ld16(tmem + ((uint32_t)(32 * quad) << 16) + (uint32_t)(l * TC_NCOL + 16 * part), v);
            tc_wait_ld();
#pragma unroll
            for (int c = 0; c < 16; ++c) m[c] += (unsigned long long)v[c] << (8 * l);
        }
        const double inv = scale[1];
        const int64_t base = a.row_base[bi];
        const int64_t i0 = (int64_t)bi * TILE, i = i0 + row;
        if (i < a.n) {
            const long long li = Lfx[i];
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                const int64_t j = (int64_t)bj * TILE + half * TC_NCOL + 16 * part + c;
                if (i < j && j < a.n) emit(a, base, i0, i, j, (double)(li + Lfx[j] - 2 * (long long)m[c]) * inv);
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == TC_XWARPS) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
    }
}

}  // namespace b2d
