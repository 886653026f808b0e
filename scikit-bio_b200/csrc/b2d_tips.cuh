// Sparse node rows as an intersection of sorted entry lists (tree problems, UniFrac).
//
// Half of the node rows are tips, and a tip row is the sparsest kind of row: sample i is
// non-zero there only if it contains that very taxon; the rows of nodes with a few tips below
// are hardly denser.  With w_qi = len_q * p_qi,
//   sum_q len_q |p_qi - p_qj| = U_i + U_j - 2 * sum_{q in both} min(w_qi, w_qj),   U_i = sum_q w_qi
// over those rows, and only rows non-zero in BOTH samples need pair work.  To keep the
// subtraction harmless the w are quantised once to 61-bit fixed point (scale = a power of two
// from max_i U_i), after which U, the intersection sums and the combination are exact integers:
// accumulation order does not matter (shared-memory integer atomics are deterministic), identical
// samples give exactly 0, and the only error left is the rounding of each w itself (2^-53
// relative), i.e. at most 2.3e-16 * (U_i + U_j) absolute per pair.  Pairs whose whole distance is
// smaller than 2.5e-4 * (U_i + U_j) - near-duplicates, where that absolute error could exceed
// 1e-12 relative - are counted by k_tip_guard, and the caller then recomputes with every row in
// the dense kernel.  Needs len >= 0 (the caller checks).
//
// This file, in order:
//   * tips only, entries from the CSR input (used when the embedding is built in several passes):
//     k_tip_usum / k_tip_scale / k_tip_count / k_tip_scan / k_tip_scatter / k_tip_sort;
//   * k_pair_tips: one CTA per tile, the column block's bucket and run starts staged in shared
//     memory, one thread per row entry walking the run of its position, min(w, w') added into a
//     128 x 128 tile of 64-bit sums (two 32-bit halves, XOR-swizzled cells); k_tip_guard;
//   * rows of nodes with up to S0 tips below, entries = non-zeros of those embedding rows
//     (whole embedding resident): k_rows_count / k_rows_ureduce / k_rows_scatter (position-sorted
//     by construction, no sort), k_compact_rows (the rows left to the dense kernel, packed);
//   * unweighted UniFrac on the presence bits: bare keys, fixed-point length table, packed bit
//     stripes: k_ubits_count / k_make_lenfx / k_ubits_scatter / k_pair_utips / k_pack_bits.
// Buckets: (sample block of 128, chunk of TP_CHUNK node positions), entries sorted by position,
// the start of every position's run kept beside them.
#pragma once

#include <cuda_pipeline.h>

#include "b2d_common.cuh"
#include "b2d_pair.cuh"

namespace b2d {

constexpr int TP_CHUNK = 2048;     // node positions per bucket
constexpr int TP_MAXE = 5120;      // column entries resident per batch (normally a whole bucket)
constexpr int TP_SORT_MAX = 8192;  // largest bucket the shared-memory counting sort accepts
constexpr int TP_THREADS = 1024;
struct TipEntry {
    uint32_t key;  // position_in_chunk << 7 | sample_in_block
    uint32_t pad;
    long long val; // fixed-point len * proportion
};
constexpr size_t TP_SMEM = (size_t)TILE * TILE * sizeof(long long) + (size_t)TP_MAXE * sizeof(TipEntry) +
                           (TP_CHUNK + 8) * sizeof(unsigned short);
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
           int* __restrict__ status, unsigned limit) {
    __shared__ long long part[1024];
    const int tid = threadIdx.x;
    const int64_t per = (nbuckets + 1023) / 1024;
    const int64_t b0 = tid * per, b1 = min(nbuckets, b0 + per);
    long long sum = 0;
    bool big = false;
    for (int64_t b = b0; b < b1; ++b) {
        sum += counts[b];
        big |= counts[b] > limit;
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
                              TipEntry* __restrict__ entries, long long* __restrict__ Ufx,
                              unsigned* __restrict__ nent) {
    const int64_t s = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (s >= n) return;
    const int64_t blk = s / TILE;
    const uint32_t sl = (uint32_t)(s % TILE);
    const long long tot = total[s];
    const double sc = scale[0];
    long long u = 0;
    unsigned cnt = 0;
    for (int64_t e = indptr[s] + lane; e < indptr[s + 1]; e += 32) {
        const int32_t q = qidx[e];
        if (flags[q] & NF_HAS_CHILDREN) continue;
        const long long w = __double2ll_rn(tip_weight(values, tot, len, q, e) * sc);
        u += w;
        ++cnt;
        const int64_t b = blk * nch + q / TP_CHUNK;
        const unsigned pos = atomicAdd(&cursor[b], 1u);
        TipEntry en;
        en.key = ((uint32_t)(q % TP_CHUNK) << 7) | sl;
        en.pad = 0;
        en.val = w;
        entries[bucket_ptr[b] + pos] = en;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        u += __shfl_xor_sync(0xffffffffu, u, o);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    if (lane == 0) {
        Ufx[s] = u;
        nent[s] = cnt;
    }
}

// 6. sort every bucket by position (counting sort in shared memory, one CTA per bucket) and keep
// the run starts: runs[bucket][f] = first entry of position f, runs[bucket][TP_CHUNK] = size
__global__ void __launch_bounds__(256)
k_tip_sort(TipEntry* __restrict__ entries, const int64_t* __restrict__ bucket_ptr,
           unsigned short* __restrict__ runs /*[nbuckets][TP_CHUNK + 8]*/) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TipEntry* buf = reinterpret_cast<TipEntry*>(smem_raw);
    int* cnt = reinterpret_cast<int*>(buf + TP_SORT_MAX);  // [TP_CHUNK + 1]
    const int64_t b0 = bucket_ptr[blockIdx.x];
    const int nb = (int)(bucket_ptr[blockIdx.x + 1] - b0);
    if (nb == 0 || nb > TP_SORT_MAX) return;
    const int tid = threadIdx.x;
    constexpr int PER = TP_CHUNK / 256;
    for (int x = tid; x <= TP_CHUNK; x += 256) cnt[x] = 0;
    for (int x = tid; x < nb; x += 256) buf[x] = entries[b0 + x];
    __syncthreads();
    for (int x = tid; x < nb; x += 256) atomicAdd(&cnt[(buf[x].key >> 7) + 1], 1);
    __syncthreads();
    __shared__ int part[256];
    int local[PER], sum = 0;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        local[k] = cnt[1 + tid * PER + k];
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
    unsigned short* rs = runs + (size_t)blockIdx.x * (TP_CHUNK + 8);
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        cnt[1 + tid * PER + k] = run;
        rs[tid * PER + k] = (unsigned short)run;
        run += local[k];
    }
    if (tid == 255) rs[TP_CHUNK] = (unsigned short)nb;
    __syncthreads();
    for (int x = tid; x < nb; x += 256) {
        const TipEntry e = buf[x];
        const int pos = atomicAdd(&cnt[(e.key >> 7) + 1], 1);
        entries[b0 + pos] = e;
    }
}

// 7. the tile kernel: T(i,j) = U_i + U_j - 2 * sum_{t in both} min(w_ti, w_tj), exact in fixed point.
// Per position chunk the column block's bucket and its run starts are staged in shared memory
// (one batch unless the bucket is larger than TP_MAXE); the row block's entries are read straight
// from global memory, one per thread, and each walks the run of its position.
__global__ void __launch_bounds__(TP_THREADS, 1)
k_pair_tips(const TipEntry* __restrict__ entries, const int64_t* __restrict__ bucket_ptr,
            const unsigned short* __restrict__ runs, int64_t nch, const long long* __restrict__ Ufx,
            const double* __restrict__ scale, unsigned long long* __restrict__ n_matches, PairArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // 64-bit sums as two 32-bit halves: shared memory has native 32-bit atomic adds only (a 64-bit
    // atomicAdd is a compare-and-swap loop); the add that wraps the low half carries into the high
    uint32_t* acc_lo = reinterpret_cast<uint32_t*>(smem_raw);   // [128][128]
    uint32_t* acc_hi = acc_lo + TILE * TILE;                    // [128][128]
    TipEntry* centry = reinterpret_cast<TipEntry*>(smem_raw + (size_t)TILE * TILE * sizeof(long long));
    unsigned short* fstart = reinterpret_cast<unsigned short*>(centry + TP_MAXE);  // [TP_CHUNK + 1]

    const int tid = threadIdx.x;
    const int bi = a.tiles[2 * blockIdx.x], bj = a.tiles[2 * blockIdx.x + 1];
    for (int x = tid; x < 2 * TILE * TILE; x += TP_THREADS) acc_lo[x] = 0u;
    unsigned nm = 0;  // matches this thread accumulated

    for (int64_t ch = 0; ch < nch; ++ch) {
        const int64_t rbk = (int64_t)bi * nch + ch, cbk = (int64_t)bj * nch + ch;
        const int64_t r0 = bucket_ptr[rbk], r1 = bucket_ptr[rbk + 1];
        const int64_t c0 = bucket_ptr[cbk], c1 = bucket_ptr[cbk + 1];
        if (r1 == r0 || c1 == c0) continue;  // uniform across the CTA
        const unsigned short* crun = runs + (size_t)cbk * (TP_CHUNK + 8);
        for (int64_t cb = c0; cb < c1; cb += TP_MAXE) {
            const int nc = (int)((c1 - cb) < TP_MAXE ? (c1 - cb) : TP_MAXE);
            const int lo = (int)(cb - c0), hi = lo + nc;  // entry range of this batch within the bucket
            __syncthreads();  // previous batch fully consumed
            for (int x = tid; x < nc; x += TP_THREADS)
                __pipeline_memcpy_async(&centry[x], &entries[cb + x], sizeof(TipEntry));
            if (cb == c0)
                for (int x = tid; x < (TP_CHUNK + 8) / 8; x += TP_THREADS)
                    __pipeline_memcpy_async(fstart + 8 * x, crun + 8 * x, 16);
            __pipeline_commit();
            __pipeline_wait_prior(0);
            __syncthreads();
            constexpr int RU = 4;  // row entries in flight per thread
            for (int64_t xb = r0 + tid; xb < r1; xb += (int64_t)RU * TP_THREADS) {
                TipEntry re[RU];
#pragma unroll
                for (int u = 0; u < RU; ++u) {
                    const int64_t x = xb + (int64_t)u * TP_THREADS;
                    if (x < r1) re[u] = entries[x];
                    else re[u].key = 0xffffffffu;
                }
#pragma unroll
                for (int u = 0; u < RU; ++u) {
                    if (re[u].key == 0xffffffffu) continue;
                    const uint32_t f = re[u].key >> 7, r = re[u].key & 127u;
                    const int e0 = max((int)fstart[f], lo), e1 = min((int)fstart[f + 1], hi);
                    nm += (unsigned)max(e1 - e0, 0);
#pragma unroll 4
                    for (int e = e0; e < e1; ++e) {
                        const TipEntry ce = centry[e - lo];
                        // min of two non-negative 61-bit integers: their bit patterns order like doubles
                        const unsigned long long add = (unsigned long long)__double_as_longlong(
                            fmin(__longlong_as_double(re[u].val), __longlong_as_double(ce.val)));
                        // lanes of a warp mostly share the position and the column entry and differ
                        // in the row sample: without the XOR they would all hit one bank
                        const uint32_t cell = r * TILE + ((ce.key & 127u) ^ (r & 31u));
                        const uint32_t alo = (uint32_t)add;
                        const uint32_t old = atomicAdd(&acc_lo[cell], alo);
                        atomicAdd(&acc_hi[cell], (uint32_t)(add >> 32) + (old + alo < old ? 1u : 0u));
                    }
                }
            }
        }
    }
    __syncthreads();
    if (n_matches) {
#pragma unroll
        for (int o = 16; o; o >>= 1) nm += __shfl_xor_sync(0xffffffffu, nm, o);
        if ((tid & 31) == 0 && nm) atomicAdd(n_matches, (unsigned long long)nm);
    }
    const double inv = scale[1];
    const int64_t base = a.row_base[bi];
    const int64_t i0 = (int64_t)bi * TILE;
    for (int x = tid; x < TILE * TILE; x += TP_THREADS) {
        const int rr = x / TILE, c = x % TILE;
        const int64_t i = i0 + rr, j = (int64_t)bj * TILE + c;
        if (i >= j || j >= a.n) continue;
        const int xs = rr * TILE + (c ^ (rr & 31));
        const long long m = (long long)(((unsigned long long)acc_hi[xs] << 32) | acc_lo[xs]);
        const long long t = Ufx[i] + Ufx[j] - 2 * m;
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

// Pairs whose fixed-point result cannot be trusted to 1e-12 relative, on the raw sums (before the
// epilogue): (a) near-duplicates, 0 < x(i,j) < 2.5e-4 * (U_i + U_j), where the 2^-53 relative
// rounding of every w (at most 2.3e-16 * (U_i + U_j) in all) could exceed 1e-12 * x; (b) pairs
// where the quantisation itself could: every entry is rounded to a multiple of q = scale[1], so
// U_i + U_j - 2 * M is off by at most (K_i + K_j) * q with K the entries of a sample (a sample
// whose U is far below max U has coarse entries).  Such pairs are appended to `list` (at most
// `cap`; `flagged` counts all of them) and k_fix_pairs_* recomputes just those pairs from the
// resident embedding, every row, in floating point - no recomputation of anything else.
struct FlaggedPair {
    int32_t i, j;
    int64_t off;   // position in the shard's compact result
};
constexpr int64_t FLAG_CAP = 1 << 20;

__global__ void __launch_bounds__(256)
k_tip_guard(const double* __restrict__ out, const double* __restrict__ out2,
            const int64_t* __restrict__ owned, const int64_t* __restrict__ row_base, int64_t n,
            const double* __restrict__ U, const unsigned* __restrict__ nent, const double* __restrict__ scale,
            unsigned long long* __restrict__ flagged, FlaggedPair* __restrict__ list,
            const double* __restrict__ Ld = nullptr, const unsigned* __restrict__ Kd = nullptr,
            const double* __restrict__ scale_d = nullptr) {
    // Ld / Kd / scale_d: a second fixed-point part of the same form (the dense rows of unweighted
    // UniFrac on the tensor cores, b2d_tc.cuh): its per-sample sums (in units of its own quantum),
    // entry counts and quantum join the two conditions
    const int64_t b = owned[blockIdx.x / TILE];
    const int64_t i0 = b * TILE, i = i0 + (blockIdx.x % TILE);
    if (i >= n - 1) return;
    const int64_t off = row_base[b] + (cond_row_start(i, n) - cond_row_start(i0, n));
    const int64_t cnt = n - 1 - i;
    const double q = scale[1], qd = Ld ? scale_d[1] : 0.0;
    const double ui = U[i] + (Ld ? Ld[i] * qd : 0.0);
    const double ki = (double)nent[i] * q + (Ld ? (double)Kd[i] * qd : 0.0);
    for (int64_t t = threadIdx.x; t < cnt; t += blockDim.x) {
        double x = out[off + t];
        if (out2) x += out2[off + t];
        const int64_t j = i + 1 + t;
        const double uj = U[j] + (Ld ? Ld[j] * qd : 0.0);
        const double kj = (double)nent[j] * q + (Ld ? (double)Kd[j] * qd : 0.0);
        if (x != 0.0 && (x < 2.5e-4 * (ui + uj) || ki + kj > 1e-12 * x)) {
            const unsigned long long slot = atomicAdd(flagged, 1ull);
            if (list && slot < (unsigned long long)FLAG_CAP) {
                FlaggedPair fp;
                fp.i = (int32_t)i;
                fp.j = (int32_t)j;
                fp.off = off + t;
                list[slot] = fp;
            }
        }
    }
}

// An embedding element as a proportion.  With `tot` < 0 the element already is one; otherwise it still holds
// the int64 subtree count of the propagation and the division k_counts_to_proportions would do (one correctly
// rounded count / total, the count itself when the total is 0: beta/_unifrac.py:399-410) is taken here, for
// the non-zero elements only - the separate pass over the whole embedding is then not needed.
__device__ __forceinline__ double emb_proportion(double raw, long long tot) {
    if (tot < 0 || raw == 0.0) return raw;
    double v = (double)__double_as_longlong(raw);
    if (tot > 0) v = v / (double)tot;
    return v;
}

// deterministic block sum (fixed tree), result valid in thread 0
__device__ __forceinline__ double block_sum_256(double v, double* sh) {
    const int tid = threadIdx.x;
    sh[tid] = v;
    __syncthreads();
    for (int o = 128; o; o >>= 1) {
        if (tid < o) sh[tid] += sh[tid + o];
        __syncthreads();
    }
    return sh[0];
}

// flagged pairs again, straight from the resident fp64 proportions: x = sum over ALL node rows of
// len_q * |p_qi - p_qj| (the reference's own per-term arithmetic, no subtraction of large sums).
// Fixed grid; a CTA takes pairs blockIdx.x, blockIdx.x + gridDim.x, ...; nothing to do -> exits.
__global__ void __launch_bounds__(256)
k_fix_pairs_weighted(const double* __restrict__ emb, const double* __restrict__ len, int64_t QR, int64_t rows,
                     const FlaggedPair* __restrict__ list, const unsigned long long* __restrict__ flagged,
                     double* __restrict__ out, double* __restrict__ out2,
                     const long long* __restrict__ totals = nullptr) {
    __shared__ double sh[256];
    const unsigned long long total = *flagged;
    const int64_t cnt = total < (unsigned long long)FLAG_CAP ? (int64_t)total : FLAG_CAP;
    for (int64_t k = blockIdx.x; k < cnt; k += gridDim.x) {
        const FlaggedPair fp = list[k];
        const double* pa = emb + ((int64_t)(fp.i / TILE) * QR) * TILE + fp.i % TILE;
        const double* pb = emb + ((int64_t)(fp.j / TILE) * QR) * TILE + fp.j % TILE;
        // the embedding holds proportions, or (totals given) still the int64 subtree counts
        const long long ta = totals ? totals[fp.i] : -1, tb = totals ? totals[fp.j] : -1;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        int64_t q = threadIdx.x;
        auto term = [&](int64_t r) {
            return fabs(emb_proportion(pa[r * TILE], ta) - emb_proportion(pb[r * TILE], tb));
        };
        for (; q + 768 < rows; q += 1024) {
            s0 = fma(term(q), len[q], s0);
            s1 = fma(term(q + 256), len[q + 256], s1);
            s2 = fma(term(q + 512), len[q + 512], s2);
            s3 = fma(term(q + 768), len[q + 768], s3);
        }
        for (; q < rows; q += 256) s0 = fma(term(q), len[q], s0);
        const double s = block_sum_256((s0 + s1) + (s2 + s3), sh);
        if (threadIdx.x == 0) {
            out[fp.off] = s;
            if (out2) out2[fp.off] = 0.0;
        }
        __syncthreads();
    }
}

// the same on the presence bits: x = sum over all node rows of len_q * [u_qi xor u_qj]
__global__ void __launch_bounds__(256)
k_fix_pairs_unweighted(const uint4* __restrict__ bits, const double* __restrict__ len, int64_t NS,
                       const FlaggedPair* __restrict__ list, const unsigned long long* __restrict__ flagged,
                       double* __restrict__ out, double* __restrict__ out2) {
    __shared__ double sh[256];
    const unsigned long long total = *flagged;
    const int64_t cnt = total < (unsigned long long)FLAG_CAP ? (int64_t)total : FLAG_CAP;
    for (int64_t k = blockIdx.x; k < cnt; k += gridDim.x) {
        const FlaggedPair fp = list[k];
        const uint4* ra = bits + ((int64_t)(fp.i / TILE) * NS) * TILE + fp.i % TILE;
        const uint4* rb = bits + ((int64_t)(fp.j / TILE) * NS) * TILE + fp.j % TILE;
        double s = 0.0;
        for (int64_t st = threadIdx.x; st < NS; st += 256) {
            const uint4 u = ra[st * TILE], v = rb[st * TILE];
            const uint32_t x[4] = {u.x ^ v.x, u.y ^ v.y, u.z ^ v.z, u.w ^ v.w};
            const double* l = len + st * NODE_STRIPE;
            double t = 0.0;
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                uint32_t m = x[w];
                while (m) {
                    const int bit = __ffs(m) - 1;
                    m &= m - 1u;
                    t += l[w * 32 + bit];
                }
            }
            s += t;
        }
        s = block_sum_256(s, sh);
        if (threadIdx.x == 0) {
            out[fp.off] = s;
            if (out2) out2[fp.off] = 0.0;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------
// Sparse rows beyond the tips (whole embedding resident).  The rows of nodes with at most S0
// tips below them (whole small subtrees; `sparse` = bitmap over node positions chosen by the
// host) are as sparse as a few tips, so they go through the same intersection kernel.  Their
// entries are the non-zeros of those embedding rows; one CTA owns one bucket (128-sample block b,
// chunk of TP_CHUNK positions), its four warps a quarter of the chunk each, a lane four samples.
// A warp walks its rows in position order, so the entries come out sorted by position and the
// run starts are written on the way - no sort.
// ---------------------------------------------------------------------------------------
constexpr int TP_QROWS = TP_CHUNK / 4;


// pass 1: entries per (bucket, quarter) and per-sample partial sums of len * p (double)
template <bool LAZY>   // LAZY: the embedding holds counts, `total` given (see emb_proportion)
__global__ void __launch_bounds__(128)
k_rows_count(const double* __restrict__ emb, const double* __restrict__ len, const uint32_t* __restrict__ sparse,
             int64_t rows, int64_t QR, int64_t nch, int64_t npad, unsigned* __restrict__ qcount /*[nb*nch][4]*/,
             double* __restrict__ upart /*[nch*4][npad]*/, const long long* __restrict__ total = nullptr) {
    const int64_t b = blockIdx.y, ch = blockIdx.x;  // chunks (up to 2^19) in x, sample blocks in y
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t qa = ch * TP_CHUNK + (int64_t)warp * TP_QROWS, qb = min(rows, qa + TP_QROWS);
    const double* p = emb + (b * QR) * TILE + lane;
    double u[4] = {0.0, 0.0, 0.0, 0.0};
    long long tot[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) tot[k] = LAZY ? total[b * TILE + 32 * k + lane] : -1;
    unsigned cnt = 0;
    for (int64_t q = qa; q < qb; ++q) {
        if (!((sparse[q >> 5] >> (q & 31)) & 1u)) continue;  // warp-uniform
        const double l = len[q];
        double raw[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) raw[k] = p[q * TILE + 32 * k];   // four loads in flight, whatever follows
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const double v = LAZY ? emb_proportion(raw[k], tot[k]) : raw[k];
            cnt += __popc(__ballot_sync(0xffffffffu, v != 0.0));
            u[k] += l * v;
        }
    }
    if (lane == 0) qcount[(b * nch + ch) * 4 + warp] = cnt;
#pragma unroll
    for (int k = 0; k < 4; ++k) upart[(ch * 4 + warp) * npad + b * TILE + 32 * k + lane] = u[k];
}

// bucket sizes from the quarter counts
__global__ void k_rows_bucket_counts(const unsigned* __restrict__ qcount, int64_t nbuckets,
                                     unsigned* __restrict__ counts) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nbuckets) counts[t] = qcount[4 * t] + qcount[4 * t + 1] + qcount[4 * t + 2] + qcount[4 * t + 3];
}

// U_i = sum of the partials in a fixed order; maximum for the scale
__global__ void k_rows_ureduce(const double* __restrict__ upart, int64_t nparts, int64_t npad, int64_t count,
                               double* __restrict__ U, unsigned long long* __restrict__ maxbits) {
    // npad: stride between the partials of one sample; count: samples to reduce (from the pointers on)
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= count) return;
    double u = 0.0;
    for (int64_t k = 0; k < nparts; ++k) u += upart[k * npad + s];
    U[s] = u;
    atomicMax(maxbits, (unsigned long long)__double_as_longlong(u));
}

// pass 2: the entries (sorted by position by construction), run starts, U_i in fixed point
template <bool LAZY>
__global__ void __launch_bounds__(128)
k_rows_scatter(const double* __restrict__ emb, const double* __restrict__ len, const uint32_t* __restrict__ sparse,
               const double* __restrict__ scale, int64_t rows, int64_t QR, int64_t nch,
               const unsigned* __restrict__ qcount, const int64_t* __restrict__ bucket_ptr,
               TipEntry* __restrict__ entries, unsigned short* __restrict__ runs,
               unsigned long long* __restrict__ Ufx, unsigned* __restrict__ nent,
               const long long* __restrict__ total = nullptr) {
    const int64_t b = blockIdx.y, ch = blockIdx.x;  // chunks (up to 2^19) in x, sample blocks in y
    const int64_t bucket = b * nch + ch;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t lt = (1u << lane) - 1u;
    unsigned run = 0;  // entries of the bucket before this warp's quarter
    for (int w = 0; w < warp; ++w) run += qcount[bucket * 4 + w];
    const int64_t qa = ch * TP_CHUNK + (int64_t)warp * TP_QROWS, qb = min(rows, qa + TP_QROWS);
    const double* p = emb + (b * QR) * TILE + lane;
    TipEntry* out = entries + bucket_ptr[bucket];
    unsigned short* rs = runs + (size_t)bucket * (TP_CHUNK + 8);
    const double sc = scale[0];
    long long ufx[4] = {0, 0, 0, 0};
    unsigned kcnt[4] = {0u, 0u, 0u, 0u};
    long long tot[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) tot[k] = LAZY ? total[b * TILE + 32 * k + lane] : -1;
    for (int64_t q = qa; q < qa + TP_QROWS; ++q) {
        if (lane == 0) rs[q - ch * TP_CHUNK] = (unsigned short)run;
        if (q >= qb || !((sparse[q >> 5] >> (q & 31)) & 1u)) continue;  // warp-uniform
        const double l = len[q];
        double raw[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) raw[k] = p[q * TILE + 32 * k];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const double v = LAZY ? emb_proportion(raw[k], tot[k]) : raw[k];
            const uint32_t m = __ballot_sync(0xffffffffu, v != 0.0);
            if (v != 0.0) {
                const long long wv = __double2ll_rn(l * v * sc);
                ufx[k] += wv;
                ++kcnt[k];
                TipEntry en;
                en.key = ((uint32_t)(q - ch * TP_CHUNK) << 7) | (uint32_t)(32 * k + lane);
                en.pad = 0;
                en.val = wv;
                out[run + __popc(m & lt)] = en;
            }
            run += __popc(m);
        }
    }
    if (warp == 3 && lane == 0) rs[TP_CHUNK] = (unsigned short)run;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (kcnt[k]) {
            atomicAdd(&Ufx[b * TILE + 32 * k + lane], (unsigned long long)ufx[k]);
            atomicAdd(&nent[b * TILE + 32 * k + lane], kcnt[k]);
        }
}

// The rows that stay with the dense pair kernel, packed: embd[block][r][128] = emb[block][pos[r] - q0][128]
// (zero rows and zero lengths in the padding), so that the pair kernel streams only them.  pos
// holds absolute node positions, q0 is the first resident row of the pass.
// With `total` given the embedding still holds the int64 subtree counts of the propagation and the
// proportion count / total (the division k_counts_to_proportions does) is taken here, for the
// packed rows only.
__global__ void __launch_bounds__(128)
k_compact_rows(const double* __restrict__ emb, const double* __restrict__ len, const int32_t* __restrict__ pos,
               int64_t nd, int64_t nd_pad, int64_t QR, int64_t q0, const long long* __restrict__ total,
               double* __restrict__ embd, double* __restrict__ lend) {
    const int64_t r = blockIdx.x, b = blockIdx.y;
    const bool real = r < nd;
    const int64_t q = real ? pos[r] : 0;
    double v = real ? emb[(b * QR + (q - q0)) * TILE + threadIdx.x] : 0.0;
    if (total && real) {
        const long long c = __double_as_longlong(v);
        const long long t = total[b * TILE + threadIdx.x];
        v = (double)c;
        if (t > 0) v = v / (double)t;
    }
    embd[(b * nd_pad + r) * TILE + threadIdx.x] = v;
    if (b == 0 && threadIdx.x == 0) lend[r] = real ? len[q] : 0.0;
}

// =======================================================================================
// Unweighted UniFrac: the same split on the presence bits.  sum_q len_q [u_q xor v_q] over the
// sparse rows = U_i + U_j - 2 * sum_{q in both} len_q with U_i = sum_q len_q [u_q], and here the
// weight of a match depends on the position only: entries are bare keys (4 bytes) and the
// fixed-point lengths lenfx[q] = round(len_q * scale) live in a table.  The rows that stay with
// the look-up-table kernel are packed into their own bit stripes.
// bits layout (b2d_embed.cuh): word ((blk*NS + q/128)*128 + sample)*4 + (q%128)/32, bit q%32.
// =======================================================================================
constexpr int UT_MAXE = 16384;  // column keys resident per batch
constexpr size_t UT_SMEM = (size_t)TILE * TILE * sizeof(long long) + (size_t)UT_MAXE * sizeof(uint32_t) +
                           2 * (TP_CHUNK + 8) * sizeof(unsigned short) + (size_t)TP_CHUNK * sizeof(long long);

// pass 1: entries per (bucket, quarter) and per-sample partial sums of the lengths (double)
__global__ void __launch_bounds__(128)
k_ubits_count(const uint32_t* __restrict__ bits, const double* __restrict__ len, const uint32_t* __restrict__ sparse,
              int64_t mpad, int64_t NS, int64_t nch, int64_t npad, unsigned* __restrict__ qcount,
              double* __restrict__ upart) {
    const int64_t b = blockIdx.y, ch = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t w0 = (ch * TP_CHUNK + (int64_t)warp * TP_QROWS) / 32;
    const int64_t w1 = min(mpad / 32, w0 + TP_QROWS / 32);
    double u[4] = {0.0, 0.0, 0.0, 0.0};
    unsigned cnt = 0;
    for (int64_t W = w0; W < w1; ++W) {
        const uint32_t mask = sparse[W];
        if (mask == 0u) continue;  // warp-uniform
        const uint32_t* src = bits + ((b * NS + W / 4) * TILE + lane) * 4 + (W & 3);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t x = src[(int64_t)32 * k * 4] & mask;
            cnt += __popc(x);
            while (x) {
                const int bit = __ffs(x) - 1;
                x &= x - 1u;
                u[k] += len[W * 32 + bit];
            }
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) qcount[(b * nch + ch) * 4 + warp] = cnt;
#pragma unroll
    for (int k = 0; k < 4; ++k) upart[(ch * 4 + warp) * npad + b * TILE + 32 * k + lane] = u[k];
}

__global__ void k_make_lenfx(const double* __restrict__ len, const double* __restrict__ scale, int64_t mpad,
                             long long* __restrict__ lenfx) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q < mpad) lenfx[q] = __double2ll_rn(len[q] * scale[0]);
}

// pass 2: the keys (sorted by position by construction), run starts, U_i in fixed point
__global__ void __launch_bounds__(128)
k_ubits_scatter(const uint32_t* __restrict__ bits, const long long* __restrict__ lenfx,
                const uint32_t* __restrict__ sparse, int64_t mpad, int64_t NS, int64_t nch,
                const unsigned* __restrict__ qcount, const int64_t* __restrict__ bucket_ptr,
                uint32_t* __restrict__ keys, unsigned short* __restrict__ runs,
                unsigned long long* __restrict__ Ufx, unsigned* __restrict__ nent) {
    const int64_t b = blockIdx.y, ch = blockIdx.x;
    const int64_t bucket = b * nch + ch;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t lt = (1u << lane) - 1u;
    unsigned run = 0;
    for (int w = 0; w < warp; ++w) run += qcount[bucket * 4 + w];
    const int64_t q0 = ch * TP_CHUNK + (int64_t)warp * TP_QROWS;
    uint32_t* out = keys + bucket_ptr[bucket];
    unsigned short* rs = runs + (size_t)bucket * (TP_CHUNK + 8);
    long long ufx[4] = {0, 0, 0, 0};
    unsigned kcnt[4] = {0u, 0u, 0u, 0u};
    for (int64_t W = q0 / 32; W < (q0 + TP_QROWS) / 32; ++W) {
        const bool in = W < mpad / 32;
        const uint32_t mask = in ? sparse[W] : 0u;
        uint32_t x[4] = {0u, 0u, 0u, 0u};
        if (mask != 0u) {
            const uint32_t* src = bits + ((b * NS + W / 4) * TILE + lane) * 4 + (W & 3);
#pragma unroll
            for (int k = 0; k < 4; ++k) x[k] = src[(int64_t)32 * k * 4] & mask;
        }
        for (int bit = 0; bit < 32; ++bit) {
            const int64_t q = W * 32 + bit;
            if (lane == 0) rs[q - ch * TP_CHUNK] = (unsigned short)run;
            if (!((mask >> bit) & 1u)) continue;  // warp-uniform
            const long long lf = lenfx[q];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const bool on = (x[k] >> bit) & 1u;
                const uint32_t m = __ballot_sync(0xffffffffu, on);
                if (on) {
                    ufx[k] += lf;
                    ++kcnt[k];
                    out[run + __popc(m & lt)] = ((uint32_t)(q - ch * TP_CHUNK) << 7) | (uint32_t)(32 * k + lane);
                }
                run += __popc(m);
            }
        }
    }
    if (warp == 3 && lane == 0) rs[TP_CHUNK] = (unsigned short)run;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (kcnt[k]) {
            atomicAdd(&Ufx[b * TILE + 32 * k + lane], (unsigned long long)ufx[k]);
            atomicAdd(&nent[b * TILE + 32 * k + lane], kcnt[k]);
        }
}

// the tile kernel: T(i,j) = U_i + U_j - 2 * sum_{q in both} lenfx[q]
__global__ void __launch_bounds__(TP_THREADS, 1)
k_pair_utips(const uint32_t* __restrict__ keys, const int64_t* __restrict__ bucket_ptr,
             const unsigned short* __restrict__ runs, const long long* __restrict__ lenfx, int64_t nch,
             const long long* __restrict__ Ufx, const double* __restrict__ scale,
             unsigned long long* __restrict__ n_matches, PairArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t* acc_lo = reinterpret_cast<uint32_t*>(smem_raw);   // [128][128], see k_pair_tips
    uint32_t* acc_hi = acc_lo + TILE * TILE;
    uint32_t* ckey = reinterpret_cast<uint32_t*>(smem_raw + (size_t)TILE * TILE * sizeof(long long));
    unsigned short* fstart = reinterpret_cast<unsigned short*>(ckey + UT_MAXE);
    unsigned short* rstart = fstart + TP_CHUNK + 8;
    long long* lf = reinterpret_cast<long long*>(rstart + TP_CHUNK + 8);  // [TP_CHUNK] lengths of the chunk

    const int tid = threadIdx.x;
    const int bi = a.tiles[2 * blockIdx.x], bj = a.tiles[2 * blockIdx.x + 1];
    for (int x = tid; x < 2 * TILE * TILE; x += TP_THREADS) acc_lo[x] = 0u;
    unsigned nm = 0;

    for (int64_t ch = 0; ch < nch; ++ch) {
        const int64_t rbk = (int64_t)bi * nch + ch, cbk = (int64_t)bj * nch + ch;
        const int64_t r0 = bucket_ptr[rbk], r1 = bucket_ptr[rbk + 1];
        const int64_t c0 = bucket_ptr[cbk], c1 = bucket_ptr[cbk + 1];
        if (r1 == r0 || c1 == c0) continue;  // uniform across the CTA
        const unsigned short* crun = runs + (size_t)cbk * (TP_CHUNK + 8);
        const unsigned short* rrun = runs + (size_t)rbk * (TP_CHUNK + 8);
        for (int64_t cb = c0; cb < c1; cb += UT_MAXE) {
            const int nc = (int)((c1 - cb) < UT_MAXE ? (c1 - cb) : UT_MAXE);
            const int lo = (int)(cb - c0), hi = lo + nc;
            __syncthreads();  // previous batch fully consumed
            for (int x = tid; x < nc; x += TP_THREADS) ckey[x] = keys[cb + x];
            if (cb == c0) {
                for (int x = tid; x < (TP_CHUNK + 8) / 8; x += TP_THREADS) {
                    reinterpret_cast<uint4*>(fstart)[x] = reinterpret_cast<const uint4*>(crun)[x];
                    reinterpret_cast<uint4*>(rstart)[x] = reinterpret_cast<const uint4*>(rrun)[x];
                }
                for (int x = tid; x < TP_CHUNK; x += TP_THREADS) lf[x] = lenfx[ch * TP_CHUNK + x];
            }
            __syncthreads();
            // one thread per (position, row sample) entry of the row bucket
            for (int64_t x = r0 + tid; x < r1; x += TP_THREADS) {
                const uint32_t rk = keys[x];
                const uint32_t f = rk >> 7, r = rk & 127u;
                const int e0 = max((int)fstart[f], lo), e1 = min((int)fstart[f + 1], hi);
                if (e1 <= e0) continue;
                nm += (unsigned)(e1 - e0);
                const unsigned long long add = (unsigned long long)lf[f];
                const uint32_t alo = (uint32_t)add, ahi = (uint32_t)(add >> 32);
                const uint32_t rbase = r * TILE, rx = r & 31u;
#pragma unroll 4
                for (int e = e0; e < e1; ++e) {
                    const uint32_t cell = rbase + ((ckey[e - lo] & 127u) ^ rx);
                    const uint32_t old = atomicAdd(&acc_lo[cell], alo);
                    atomicAdd(&acc_hi[cell], ahi + (old + alo < old ? 1u : 0u));
                }
            }
        }
    }
    __syncthreads();
    if (n_matches) {
#pragma unroll
        for (int o = 16; o; o >>= 1) nm += __shfl_xor_sync(0xffffffffu, nm, o);
        if ((tid & 31) == 0 && nm) atomicAdd(n_matches, (unsigned long long)nm);
    }
    const double inv = scale[1];
    const int64_t base = a.row_base[bi];
    const int64_t i0 = (int64_t)bi * TILE;
    for (int x = tid; x < TILE * TILE; x += TP_THREADS) {
        const int rr = x / TILE, c = x % TILE;
        const int64_t i = i0 + rr, j = (int64_t)bj * TILE + c;
        if (i >= j || j >= a.n) continue;
        const int xs = rr * TILE + (c ^ (rr & 31));
        const long long m = (long long)(((unsigned long long)acc_hi[xs] << 32) | acc_lo[xs]);
        const long long t = Ufx[i] + Ufx[j] - 2 * m;
        emit(a, base, i0, i, j, (double)t * inv);
    }
}

// the rows left to the look-up-table kernel, packed into their own stripes:
// bit r of bitsd (same layout, NSd stripes) = bit pos[r] of bits; lend[r] = len[pos[r]]
__global__ void __launch_bounds__(128)
k_pack_bits(const uint32_t* __restrict__ bits, const double* __restrict__ len, const int32_t* __restrict__ pos,
            int64_t nd, int64_t NS, int64_t NSd, uint32_t* __restrict__ bitsd, double* __restrict__ lend) {
    const int64_t sd = blockIdx.x, b = blockIdx.y;
    const int sample = threadIdx.x;
    const uint32_t* src = bits + (b * NS * TILE + sample) * 4;
    uint32_t wv[4];
#pragma unroll
    for (int w = 0; w < 4; ++w) {
        uint32_t word = 0;
        for (int bit = 0; bit < 32; ++bit) {
            const int64_t r = sd * 128 + w * 32 + bit;
            if (r >= nd) break;
            const int64_t q = pos[r];
            word |= ((src[(q / 128) * TILE * 4 + (q % 128) / 32] >> (q % 32)) & 1u) << bit;
        }
        wv[w] = word;
    }
    *reinterpret_cast<uint4*>(bitsd + ((b * NSd + sd) * TILE + sample) * 4) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
    if (b == 0) {
        const int64_t r = sd * 128 + sample;
        lend[r] = r < nd ? len[pos[r]] : 0.0;
    }
}

}  // namespace b2d
