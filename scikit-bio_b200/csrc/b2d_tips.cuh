// Tip rows of the weighted kernel as a sparse intersection (tree problems, zero-block skipping on).
//
// Half of the node rows are tips, and a tip row is the sparsest kind of row: sample i is
// non-zero there only if it contains that very taxon.  With w_ti = len_t * p_ti,
//   sum_t len_t |p_ti - p_tj| = U_i + U_j - 2 * sum_{t in both} min(w_ti, w_tj),   U_i = sum_t w_ti
// and only tips present in BOTH samples need pair work (about 40 of 2 x 2000 per pair at config 2).
// To keep the subtraction harmless the w are quantised once to 61-bit fixed point (scale = a
// power of two from max_i U_i), after which U, the intersection sums and the combination are
// exact integers: accumulation order does not matter (shared-memory integer atomics are
// deterministic), identical samples give exactly 0, and the only error left is the rounding of
// each w itself (2^-53 relative), i.e. at most 2.3e-16 * (U_i + U_j) absolute per pair.  Pairs
// whose total weighted UniFrac is smaller than 2.5e-4 * (U_i + U_j) - near-duplicates, where that
// absolute error could exceed 1e-12 relative - are counted by k_tip_guard, and the caller then
// recomputes with the tips back in the dense kernel.
//
// The pair kernel then never visits a tip row (its bits are cleared in the row-major bitmaps)
// and the closed form of b2d_skip.cuh uses root distances that leave the tips' own branches out.
//
// Layout and kernel follow b2d_sparse.cuh: CSR entries on tips are bucketed by (sample block of
// 128, chunk of 4096 node positions), sorted by position inside a bucket; a CTA owns one tile,
// stages the column block's bucket with its per-position run boundaries and lets every lane walk
// the run of one row entry, adding min(w, w') into a 128 x 128 int64 tile in shared memory.
#pragma once

#include <cuda_pipeline.h>

#include "b2d_common.cuh"
#include "b2d_pair.cuh"

namespace b2d {

constexpr int TP_CHUNK = 4096;
constexpr int TP_MAXE = 2048;      // column entries resident per batch
constexpr int TP_MAXR = 2048;      // row entries resident per batch
constexpr int TP_SORT_MAX = 8192;  // largest bucket the shared-memory counting sort accepts
constexpr int TP_THREADS = 512;
struct TipEntry {
    uint32_t key;  // position_in_chunk << 7 | sample_in_block
    uint32_t pad;
    long long val; // fixed-point len * proportion
};
constexpr size_t TP_SMEM = (size_t)TILE * TILE * sizeof(long long) + 2 * TP_CHUNK * sizeof(unsigned short) +
                           (size_t)(TP_MAXE + TP_MAXR) * sizeof(TipEntry);
constexpr size_t TP_SORT_SMEM = (size_t)TP_SORT_MAX * sizeof(TipEntry) + (TP_CHUNK + 1) * sizeof(int);

__device__ __forceinline__ double tip_weight(const int64_t* values, const long long tot, const double* len,
                                             int32_t q, int64_t e) {
    double pv = values ? (double)values[e] : 1.0;
    if (tot > 0) pv = pv / (double)tot;
    return len[q] * pv;
}

// 1. U_i in double (only to find the scale) and its maximum (positive doubles order like ints)
__global__ void k_tip_usum(const int64_t* __restrict__ indptr, const int32_t* __restrict__ qidx,
                           const int64_t* __restrict__ values, const long long* __restrict__ total,
                           const uint8_t* __restrict__ flags, const double* __restrict__ len, int64_t n,
                           double* __restrict__ U, unsigned long long* __restrict__ maxbits) {
    const int64_t s = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (s >= n) return;
    const long long tot = total[s];
    double u = 0.0;
    for (int64_t e = indptr[s] + lane; e < indptr[s + 1]; e += 32) {
        const int32_t q = qidx[e];
        if (!(flags[q] & NF_HAS_CHILDREN)) u += tip_weight(values, tot, len, q, e);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) u += __shfl_xor_sync(0xffffffffu, u, o);
    if (lane == 0) {
        U[s] = u;
        atomicMax(maxbits, (unsigned long long)__double_as_longlong(u));
    }
}

// 2. scale[0] = 2^(61 - e), scale[1] = 2^(e - 61) with 2^e >= max U (1 when there is nothing)
__global__ void k_tip_scale(const unsigned long long* __restrict__ maxbits, double* __restrict__ scale) {
    const double mx = __longlong_as_double((long long)*maxbits);
    int e = 0;
    if (mx > 0.0) {
        frexp(mx, &e);  // mx = f * 2^e, f in [0.5, 1)
    } else {
        e = 61;
    }
    scale[0] = ldexp(1.0, 61 - e);
    scale[1] = ldexp(1.0, e - 61);
}

// 3. bucket histogram (tip entries only)
__global__ void k_tip_count(const int64_t* __restrict__ indptr, const int32_t* __restrict__ qidx,
                            const uint8_t* __restrict__ flags, int64_t n, int64_t nch,
                            unsigned* __restrict__ counts) {
    const int64_t s = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (s >= n) return;
    const int64_t blk = s / TILE;
    for (int64_t e = indptr[s] + lane; e < indptr[s + 1]; e += 32) {
        const int32_t q = qidx[e];
        if (!(flags[q] & NF_HAS_CHILDREN)) atomicAdd(&counts[blk * nch + q / TP_CHUNK], 1u);
    }
}

// 4. exclusive scan of the bucket sizes (one CTA) + "a bucket is too large" flag
__global__ void __launch_bounds__(1024)
k_tip_scan(const unsigned* __restrict__ counts, int64_t nbuckets, int64_t* __restrict__ ptr,
           int* __restrict__ status) {
    __shared__ long long part[1024];
    const int tid = threadIdx.x;
    const int64_t per = (nbuckets + 1023) / 1024;
    const int64_t b0 = tid * per, b1 = min(nbuckets, b0 + per);
    long long sum = 0;
    bool big = false;
    for (int64_t b = b0; b < b1; ++b) {
        sum += counts[b];
        big |= counts[b] > (unsigned)TP_SORT_MAX;
    }
    if (big) atomicOr(status, 1);
    part[tid] = sum;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
        const long long v = tid >= off ? part[tid - off] : 0;
        __syncthreads();
        part[tid] += v;
        __syncthreads();
    }
    long long run = part[tid] - sum;
    for (int64_t b = b0; b < b1; ++b) {
        ptr[b] = run;
        run += counts[b];
    }
    if (tid == 1023) ptr[nbuckets] = part[1023];
}

// 5. scatter the quantised entries into their buckets; U_i in fixed point (exact integer sum)
__global__ void k_tip_scatter(const int64_t* __restrict__ indptr, const int32_t* __restrict__ qidx,
                              const int64_t* __restrict__ values, const long long* __restrict__ total,
                              const uint8_t* __restrict__ flags, const double* __restrict__ len,
                              const double* __restrict__ scale, int64_t n, int64_t nch,
                              const int64_t* __restrict__ bucket_ptr, unsigned* __restrict__ cursor,
                              TipEntry* __restrict__ entries, long long* __restrict__ Ufx) {
    const int64_t s = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (s >= n) return;
    const int64_t blk = s / TILE;
    const uint32_t sl = (uint32_t)(s % TILE);
    const long long tot = total[s];
    const double sc = scale[0];
    long long u = 0;
    for (int64_t e = indptr[s] + lane; e < indptr[s + 1]; e += 32) {
        const int32_t q = qidx[e];
        if (flags[q] & NF_HAS_CHILDREN) continue;
        const long long w = __double2ll_rn(tip_weight(values, tot, len, q, e) * sc);
        u += w;
        const int64_t b = blk * nch + q / TP_CHUNK;
        const unsigned pos = atomicAdd(&cursor[b], 1u);
        TipEntry en;
        en.key = ((uint32_t)(q % TP_CHUNK) << 7) | sl;
        en.pad = 0;
        en.val = w;
        entries[bucket_ptr[b] + pos] = en;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) u += __shfl_xor_sync(0xffffffffu, u, o);
    if (lane == 0) Ufx[s] = u;
}

// 6. sort every bucket by position (counting sort in shared memory, one CTA per bucket)
__global__ void __launch_bounds__(256)
k_tip_sort(TipEntry* __restrict__ entries, const int64_t* __restrict__ bucket_ptr) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TipEntry* buf = reinterpret_cast<TipEntry*>(smem_raw);
    int* cnt = reinterpret_cast<int*>(buf + TP_SORT_MAX);  // [TP_CHUNK + 1]
    const int64_t b0 = bucket_ptr[blockIdx.x];
    const int nb = (int)(bucket_ptr[blockIdx.x + 1] - b0);
    if (nb <= 1 || nb > TP_SORT_MAX) return;
    const int tid = threadIdx.x;
    for (int x = tid; x <= TP_CHUNK; x += 256) cnt[x] = 0;
    for (int x = tid; x < nb; x += 256) buf[x] = entries[b0 + x];
    __syncthreads();
    for (int x = tid; x < nb; x += 256) atomicAdd(&cnt[(buf[x].key >> 7) + 1], 1);
    __syncthreads();
    __shared__ int part[256];
    int local[16], sum = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        local[k] = cnt[1 + tid * 16 + k];
        sum += local[k];
    }
    part[tid] = sum;
    __syncthreads();
    for (int off = 1; off < 256; off <<= 1) {
        const int v = tid >= off ? part[tid - off] : 0;
        __syncthreads();
        part[tid] += v;
        __syncthreads();
    }
    int run = part[tid] - sum;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        cnt[1 + tid * 16 + k] = run;
        run += local[k];
    }
    __syncthreads();
    for (int x = tid; x < nb; x += 256) {
        const TipEntry e = buf[x];
        const int pos = atomicAdd(&cnt[(e.key >> 7) + 1], 1);
        entries[b0 + pos] = e;
    }
}

// 7. the tile kernel: T(i,j) = U_i + U_j - 2 * sum_{t in both} min(w_ti, w_tj), exact in fixed point
__global__ void __launch_bounds__(TP_THREADS, 1)
k_pair_tips(const TipEntry* __restrict__ entries, const int64_t* __restrict__ bucket_ptr, int64_t nch,
            const long long* __restrict__ Ufx, const double* __restrict__ scale, PairArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long* acc = reinterpret_cast<unsigned long long*>(smem_raw);  // [128][128]
    unsigned short* fstart = reinterpret_cast<unsigned short*>(smem_raw + (size_t)TILE * TILE * sizeof(long long));
    unsigned short* fend = fstart + TP_CHUNK;
    TipEntry* centry = reinterpret_cast<TipEntry*>(fend + TP_CHUNK);
    TipEntry* rentry = centry + TP_MAXE;

    const int tid = threadIdx.x;
    const int bi = a.tiles[2 * blockIdx.x], bj = a.tiles[2 * blockIdx.x + 1];
    for (int x = tid; x < TILE * TILE; x += TP_THREADS) acc[x] = 0ull;

    for (int64_t ch = 0; ch < nch; ++ch) {
        const int64_t r0 = bucket_ptr[(int64_t)bi * nch + ch], r1 = bucket_ptr[(int64_t)bi * nch + ch + 1];
        const int64_t c0 = bucket_ptr[(int64_t)bj * nch + ch], c1 = bucket_ptr[(int64_t)bj * nch + ch + 1];
        if (r1 == r0 || c1 == c0) continue;  // uniform across the CTA
        for (int64_t cb = c0; cb < c1; cb += TP_MAXE) {
            const int nc = (int)((c1 - cb) < TP_MAXE ? (c1 - cb) : TP_MAXE);
            __syncthreads();  // previous batch fully consumed
            for (int x = tid; x < TP_CHUNK; x += TP_THREADS) {
                fstart[x] = 0;
                fend[x] = 0;
            }
            for (int x = tid; x < nc; x += TP_THREADS)
                __pipeline_memcpy_async(&centry[x], &entries[cb + x], sizeof(TipEntry));
            __pipeline_commit();
            __pipeline_wait_prior(0);
            __syncthreads();
            for (int x = tid; x < nc; x += TP_THREADS) {
                const uint32_t f = centry[x].key >> 7;
                if (x == 0 || (centry[x - 1].key >> 7) != f) fstart[f] = (unsigned short)x;
                if (x == nc - 1 || (centry[x + 1].key >> 7) != f) fend[f] = (unsigned short)(x + 1);
            }
            for (int64_t rb = r0; rb < r1; rb += TP_MAXR) {
                const int nr = (int)((r1 - rb) < TP_MAXR ? (r1 - rb) : TP_MAXR);
                __syncthreads();  // boundaries complete / previous R batch consumed
                for (int x = tid; x < nr; x += TP_THREADS)
                    __pipeline_memcpy_async(&rentry[x], &entries[rb + x], sizeof(TipEntry));
                __pipeline_commit();
                __pipeline_wait_prior(0);
                __syncthreads();
                for (int x = tid; x < nr; x += TP_THREADS) {
                    const TipEntry re = rentry[x];
                    const uint32_t f = re.key >> 7, r = re.key & 127u;
                    const int xe = fend[f];
                    for (int e = fstart[f]; e < xe; ++e) {
                        const TipEntry ce = centry[e];
                        const long long add = re.val < ce.val ? re.val : ce.val;
                        atomicAdd(&acc[r * TILE + (ce.key & 127u)], (unsigned long long)add);
                    }
                }
            }
        }
    }
    __syncthreads();
    const double inv = scale[1];
    const int64_t base = a.row_base[bi];
    const int64_t i0 = (int64_t)bi * TILE;
    for (int x = tid; x < TILE * TILE; x += TP_THREADS) {
        const int rr = x / TILE, c = x % TILE;
        const int64_t i = i0 + rr, j = (int64_t)bj * TILE + c;
        if (i >= j || j >= a.n) continue;
        const long long t = Ufx[i] + Ufx[j] - 2 * (long long)acc[x];
        emit(a, base, i0, i, j, (double)t * inv);
    }
}

// rows of tips are dead for the pair kernel: nz[g][w] &= has-children bits of the resident rows
__global__ void k_mask_tip_rows(uint32_t* __restrict__ nz, const uint32_t* __restrict__ hc, int64_t ngroups,
                                int64_t words, int64_t nwords) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ngroups * nwords) return;
    const int64_t g = t / nwords, w = t % nwords;
    nz[g * words + w] &= hc[w];
}

// near-duplicate pairs: 0 < x(i,j) < 2.5e-4 * (U_i + U_j) on the raw sums (before the epilogue)
__global__ void __launch_bounds__(256)
k_tip_guard(const double* __restrict__ out, const double* __restrict__ out2,
            const int64_t* __restrict__ owned, const int64_t* __restrict__ row_base, int64_t n,
            const double* __restrict__ U, unsigned long long* __restrict__ flagged) {
    const int64_t b = owned[blockIdx.x / TILE];
    const int64_t i0 = b * TILE, i = i0 + (blockIdx.x % TILE);
    if (i >= n - 1) return;
    const int64_t off = row_base[b] + (cond_row_start(i, n) - cond_row_start(i0, n));
    const int64_t cnt = n - 1 - i;
    const double ui = U[i];
    unsigned c = 0;
    for (int64_t t = threadIdx.x; t < cnt; t += blockDim.x) {
        double x = out[off + t];
        if (out2) x += out2[off + t];
        if (x != 0.0 && x < 2.5e-4 * (ui + U[i + 1 + t])) ++c;
    }
    if (c) atomicAdd(flagged, (unsigned long long)c);
}

}  // namespace b2d
