// Sparse pair kernels for feature tables with integer counts (Bray-Curtis, Jaccard).
//
// For non-negative integer counts  sum|a-b| = S_a + S_b - 2*I(a,b)  with
// I(a,b) = sum_f min(a_f, b_f)  and  |a xor b| = c_a + c_b - 2*|a and b|.  All quantities are
// exact integers, so evaluating only the INTERSECTION of two samples' non-zero features is
// bit-identical to SciPy's dense loops (numerator and denominator are the same integers, one
// fp64 division).  At the densities of BASELINE config 3 (~1 % non-zero) a 128x128 sample
// tile has ~1.6 matches per feature instead of 16 384 dense terms.
//
// Layout: the CSR entries are bucketed by (sample block of 128, feature chunk of 4096); an
// entry is {key = feature_in_chunk << 7 | sample_in_block, value}.  A CTA owns one tile
// (row block R, column block C); per feature chunk it threads the C bucket into per-feature
// linked lists in shared memory (head[4096], next[]), then streams the R bucket and, for each
// R entry, walks the list of its feature and adds min(v, w) (or 1) into a 128x128 int32
// accumulator tile held in shared memory.
#pragma once

#include <cuda_pipeline.h>

#include "b2d_common.cuh"
#include "b2d_pair.cuh"

namespace b2d {

constexpr int SP_CHUNK = 4096;     // features per chunk (12 bits)
constexpr int SP_MAXE = 8192;      // C entries resident per batch
constexpr int SP_SORT_MAX = 20480; // largest bucket the shared-memory counting sort accepts
constexpr int SP_THREADS = 512;    // 128 rows x 4 column groups
struct SpEntry {
    uint32_t key;   // feature_in_chunk << 7 | sample_in_block
    int32_t val;
};
constexpr size_t SP_SMEM = (size_t)TILE * TILE * sizeof(int) + 2 * SP_CHUNK * sizeof(unsigned short) +
                           (size_t)SP_MAXE * sizeof(SpEntry);
constexpr size_t SP_SORT_SMEM = (size_t)SP_SORT_MAX * sizeof(SpEntry) + (SP_CHUNK + 1) * sizeof(int);

// 1. histogram of entries per (block, chunk)
__global__ void k_sp_count(const int64_t* __restrict__ indptr, const int32_t* __restrict__ feat,
                           int64_t n, int64_t nch, unsigned* __restrict__ counts) {
    int64_t s = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (s >= n) return;
    int64_t blk = s / TILE;
    for (int64_t e = indptr[s] + lane; e < indptr[s + 1]; e += 32)
        atomicAdd(&counts[blk * nch + feat[e] / SP_CHUNK], 1u);
}

// 2. scatter entries into their buckets
template <typename V>
__global__ void k_sp_scatter(const int64_t* __restrict__ indptr, const int32_t* __restrict__ feat,
                             const V* __restrict__ values, int64_t n, int64_t nch,
                             const int64_t* __restrict__ bucket_ptr, unsigned* __restrict__ cursor,
                             SpEntry* __restrict__ entries) {
    int64_t s = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (s >= n) return;
    int64_t blk = s / TILE;
    uint32_t sl = (uint32_t)(s % TILE);
    for (int64_t e = indptr[s] + lane; e < indptr[s + 1]; e += 32) {
        int32_t f = feat[e];
        int64_t b = blk * nch + f / SP_CHUNK;
        unsigned pos = atomicAdd(&cursor[b], 1u);
        SpEntry en;
        en.key = ((uint32_t)(f % SP_CHUNK) << 7) | sl;
        en.val = values ? (int32_t)values[e] : 1;
        entries[bucket_ptr[b] + pos] = en;
    }
}

// 3. sort every bucket by feature (counting sort in shared memory, one CTA per bucket)
__global__ void __launch_bounds__(256)
k_sp_sort(SpEntry* __restrict__ entries, const int64_t* __restrict__ bucket_ptr,
          unsigned short* __restrict__ runs /*[nbuckets][SP_CHUNK + 8] run start of every feature*/) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SpEntry* buf = reinterpret_cast<SpEntry*>(smem_raw);
    int* cnt = reinterpret_cast<int*>(buf + SP_SORT_MAX);  // [SP_CHUNK + 1]
    const int64_t b0 = bucket_ptr[blockIdx.x];
    const int nb = (int)(bucket_ptr[blockIdx.x + 1] - b0);
    if (nb == 0) return;   // an empty bucket is never staged: its run table stays unwritten
    const int tid = threadIdx.x;
    unsigned short* rs = runs + (size_t)blockIdx.x * (SP_CHUNK + 8);
    for (int x = tid; x <= SP_CHUNK; x += 256) cnt[x] = 0;
    for (int x = tid; x < nb; x += 256) buf[x] = entries[b0 + x];
    __syncthreads();
    for (int x = tid; x < nb; x += 256) atomicAdd(&cnt[(buf[x].key >> 7) + 1], 1);
    __syncthreads();
    // exclusive scan of 4096 counters: 16 per thread + block scan of the partial sums
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
        int v = tid >= off ? part[tid - off] : 0;
        __syncthreads();
        part[tid] += v;
        __syncthreads();
    }
    int run = part[tid] - sum;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        cnt[1 + tid * 16 + k] = run;   // start offset of feature tid*16+k (stored at index f+1)
        rs[tid * 16 + k] = (unsigned short)run;
        run += local[k];
    }
    if (tid == 255) rs[SP_CHUNK] = (unsigned short)nb;
    __syncthreads();
    for (int x = tid; x < nb; x += 256) {
        const SpEntry e = buf[x];
        const int pos = atomicAdd(&cnt[(e.key >> 7) + 1], 1);
        entries[b0 + pos] = e;
    }
}

// values must be non-negative integers below 2^31 and row sums below 2^31 for the int32 tile
__global__ void k_sp_check(const double* __restrict__ values, int64_t nnz, int* __restrict__ bad) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nnz) return;
    double v = values[e];
    if (!(v >= 0.0) || v >= 2147483648.0 || v != floor(v)) *bad = 1;
}

// 5. the tile kernel.  binary = 1: count common features (Jaccard), else sum of minima.
// Per feature chunk the column block's feature-sorted bucket is staged in shared memory with
// its run boundaries (fstart/fend per feature, plain stores); the row block's bucket is
// staged too and every lane takes one row entry, walks the run of its feature and adds into
// the 128x128 int32 accumulator tile with shared-memory atomics (two row entries of one
// sample may hit the same cell).  Bound by the shared-memory atomic rate (~2 cycles per lane).
constexpr int SP_MAXR = 4096;
constexpr size_t SP_SMEM2 = SP_SMEM + (size_t)SP_MAXR * sizeof(SpEntry);

__global__ void __launch_bounds__(SP_THREADS, 1)
k_pair_sparse(const SpEntry* __restrict__ entries, const int64_t* __restrict__ bucket_ptr,
              int64_t nch, int binary, PairArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int* acc = reinterpret_cast<int*>(smem_raw);                                  // [128][128]
    unsigned short* fstart = reinterpret_cast<unsigned short*>(smem_raw + (size_t)TILE * TILE * sizeof(int));
    unsigned short* fend = fstart + SP_CHUNK;
    SpEntry* centry = reinterpret_cast<SpEntry*>(fend + SP_CHUNK);
    SpEntry* rentry = centry + SP_MAXE;

    const int tid = threadIdx.x;
    const int bi = a.tiles[2 * blockIdx.x], bj = a.tiles[2 * blockIdx.x + 1];
    for (int x = tid; x < TILE * TILE; x += SP_THREADS) acc[x] = 0;

    for (int64_t ch = 0; ch < nch; ++ch) {
        const int64_t r0 = bucket_ptr[(int64_t)bi * nch + ch], r1 = bucket_ptr[(int64_t)bi * nch + ch + 1];
        const int64_t c0 = bucket_ptr[(int64_t)bj * nch + ch], c1 = bucket_ptr[(int64_t)bj * nch + ch + 1];
        if (r1 == r0 || c1 == c0) continue;  // uniform across the CTA
        for (int64_t cb = c0; cb < c1; cb += SP_MAXE) {
            const int nc = (int)((c1 - cb) < SP_MAXE ? (c1 - cb) : SP_MAXE);
            __syncthreads();  // previous batch fully consumed
            for (int x = tid; x < SP_CHUNK; x += SP_THREADS) {
                fstart[x] = 0;
                fend[x] = 0;
            }
            for (int x = tid; x < nc; x += SP_THREADS)
                __pipeline_memcpy_async(&centry[x], &entries[cb + x], sizeof(SpEntry));
            __pipeline_commit();
            __pipeline_wait_prior(0);
            __syncthreads();
            // run boundaries of the feature-sorted batch
            for (int x = tid; x < nc; x += SP_THREADS) {
                const uint32_t f = centry[x].key >> 7;
                if (x == 0 || (centry[x - 1].key >> 7) != f) fstart[f] = (unsigned short)x;
                if (x == nc - 1 || (centry[x + 1].key >> 7) != f) fend[f] = (unsigned short)(x + 1);
            }
            for (int64_t rb = r0; rb < r1; rb += SP_MAXR) {
                const int nr = (int)((r1 - rb) < SP_MAXR ? (r1 - rb) : SP_MAXR);
                __syncthreads();  // boundaries complete / previous R batch consumed
                for (int x = tid; x < nr; x += SP_THREADS)
                    __pipeline_memcpy_async(&rentry[x], &entries[rb + x], sizeof(SpEntry));
                __pipeline_commit();
                __pipeline_wait_prior(0);
                __syncthreads();
                for (int x = tid; x < nr; x += SP_THREADS) {
                    const SpEntry re = rentry[x];
                    const uint32_t f = re.key >> 7, r = re.key & 127u;
                    const int xe = fend[f];
                    for (int e = fstart[f]; e < xe; ++e) {
                        const SpEntry ce = centry[e];
                        const int add = binary ? 1 : (re.val < ce.val ? re.val : ce.val);
                        atomicAdd(&acc[r * TILE + (ce.key & 127u)], add);
                    }
                }
            }
        }
    }
    __syncthreads();
    const int64_t base = a.row_base[bi];
    const int64_t i0 = (int64_t)bi * TILE;
    for (int x = tid; x < TILE * TILE; x += SP_THREADS) {
        const int rr = x / TILE, c = x % TILE;
        emit(a, base, i0, i0 + rr, (int64_t)bj * TILE + c, (double)acc[x]);
    }
}

}  // namespace b2d
