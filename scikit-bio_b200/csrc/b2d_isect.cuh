// Intersection tile kernel for the sparse node rows (weighted and unweighted UniFrac), second
// generation: a warp-specialised producer / consumer pipeline instead of barriers around a
// single staging buffer (k_pair_tips / k_pair_utips in b2d_tips.cuh, kept as cross-checks).
//
//   T(i,j) = U_i + U_j - 2 * sum_{q in both} min(w_qi, w_qj)       (weighted,  b2d_tips.cuh)
//   T(i,j) = U_i + U_j - 2 * sum_{q in both} lenfx_q                (unweighted)
//   I(i,j) = sum_{f in both} min(a_f, b_f)  or  |a and b|           (integer feature tables:
//                                                                    Bray-Curtis / Jaccard, b2d_sparse.cuh)
//
// exact in 61-bit fixed point (32-bit integers for the feature tables).  One CTA per 128 x 128 tile of sample pairs:
//   * warp 31 is the PRODUCER.  Per (tile, position chunk) it cuts the bucket pair into batches at
//     position boundaries (both buckets whole when they fit; a two-step 32-lane search over the
//     run tables otherwise), and for every batch issues 1-D bulk copies (cp.async.bulk, SASS
//     UBLKCP) of the column entries, the ROW entries, the slice of the column run table and -
//     unweighted - the fixed-point lengths into one of the stages, with a small header (counts,
//     offsets, position range).  full / empty mbarriers per stage; the next batch lands while the
//     current one is consumed.
//   * warps 0..30 are CONSUMERS and touch shared memory only.  A warp claims eight row entries at
//     a time through a counter in the stage header (dynamic: the cost of a row entry is the length
//     of its column run, anything static leaves warps waiting at every stage boundary); G = 4
//     lanes share one row entry and stride over the column run of its position, so a run of k
//     column entries costs ceil(k / 4) steps instead of k (the one-thread-per-row-entry walk of
//     the first generation left 43 % of the lanes idle: run lengths differ from lane to lane).
//     Every match adds into a 128 x 128 tile of 64-bit sums kept as two 32-bit halves (shared
//     memory has native 32-bit atomic adds only; the add that wraps the low half carries into the
//     high one), cells XOR-swizzled by the row.
//   * no __syncthreads between the prologue (zeroing the tile) and the epilogue (emit).
#pragma once

#include "b2d_common.cuh"
#include "b2d_pair.cuh"
#include "b2d_sparse.cuh"
#include "b2d_tips.cuh"

namespace b2d {

// what a match adds, and what the entries are
enum IsMode : int {
    IS_WEIGHTED = 0,    // TipEntry {key, fixed-point w}: += min(w, w'), 64-bit sums
    IS_UNWEIGHTED = 1,  // bare uint32 keys + fixed-point length table: += lenfx[position], 64-bit sums
    IS_FEAT_MIN = 2,    // SpEntry {key, int32 count}: += min(v, v'), 32-bit sums (Bray-Curtis, b2d_sparse.cuh)
    IS_FEAT_COUNT = 3   // SpEntry, value ignored: += 1 (Jaccard)
};

constexpr int IS_THREADS = 1024;
constexpr int IS_CWARPS = 31;                       // consumer warps; warp 31 produces
struct IsHeader {
    int nrow;               // row entries of the batch
    int roff;               // row entry k of the batch lives in row slot k + roff
    int coff;               // column entry e (relative to its bucket) lives in column slot e + coff
    int f0;                 // first position of the batch (multiple of 8)
    int done;               // 1: no more work for this tile
    int next;               // first row entry of the batch not yet claimed by a consumer warp
    int pad[10];
};
static_assert(sizeof(IsHeader) == 64, "header is 64 bytes");

template <int MODE, int CHUNK>
struct IsGeom {
    static constexpr bool wide = MODE == IS_WEIGHTED || MODE == IS_UNWEIGHTED;        // 64-bit sums
    static constexpr int esize = MODE == IS_WEIGHTED ? 16 : MODE == IS_UNWEIGHTED ? 4 : 8;
    // feature tables: two stages of 4 092 entries beat three of 2 044 (33.6 -> 31.2 ms at config 3) and four of
    // 1 532 (41 ms): a (tile, chunk) holds few matches, so the fixed cost of a batch is what counts
    static constexpr int stages = 2;
    // entries per stage and side (column block / row block)
    static constexpr int ent_bytes = MODE == IS_WEIGHTED ? 20480 : MODE == IS_UNWEIGHTED ? 14336 : 32768;
    static constexpr int align_entries = 16 / esize;                                  // bulk copies: 16-byte units
    static constexpr int cap = ent_bytes / esize - 2 * align_entries;                 // entries per batch and side
    static constexpr int run_bytes = (CHUNK + 8) * 2 + 16;                            // run starts of a whole chunk
    static constexpr int lf_bytes = MODE == IS_UNWEIGHTED ? CHUNK * 8 : 0;            // fixed-point lengths
    static constexpr int stage_bytes = 2 * ent_bytes + run_bytes + lf_bytes + 64;
    static constexpr size_t acc_bytes = (size_t)TILE * TILE * (wide ? 8 : 4);
    static constexpr size_t smem = acc_bytes + (size_t)stages * stage_bytes + 2 * stages * sizeof(uint64_t);
    static_assert(cap >= 8 * TILE, "eight positions of a block must fit one batch");
    static_assert(stage_bytes % 16 == 0 && smem <= 232448, "shared memory per CTA");
};

// entries_v: TipEntry[] (weighted), uint32 keys[] (unweighted) or SpEntry[] (feature tables);
// runs: [nbuckets][CHUNK + 8] run starts per position / feature of every bucket.
template <int MODE, int G, int CHUNK>
__global__ void __launch_bounds__(IS_THREADS, 1)
k_pair_isect(const void* __restrict__ entries_v, const int64_t* __restrict__ bucket_ptr,
             const unsigned short* __restrict__ runs, const long long* __restrict__ lenfx, int64_t nch,
             const long long* __restrict__ Ufx, const double* __restrict__ scale,
             unsigned long long* __restrict__ n_matches, PairArgs a) {
    using Geo = IsGeom<MODE, CHUNK>;
    constexpr bool WIDE = Geo::wide;
    constexpr int ENT = Geo::ent_bytes, CAP = Geo::cap, STAGE = Geo::stage_bytes, NST = Geo::stages;
    constexpr int RUNS = 2 * ENT, LF = RUNS + Geo::run_bytes, HDR = LF + Geo::lf_bytes;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint32_t* acc_lo = reinterpret_cast<uint32_t*>(smem_raw);   // [128][128]
    uint32_t* acc_hi = acc_lo + TILE * TILE;                    // [128][128], 64-bit modes only
    unsigned char* stage0 = smem_raw + Geo::acc_bytes;
    uint64_t* full = reinterpret_cast<uint64_t*>(stage0 + (size_t)NST * STAGE);
    uint64_t* empty = full + NST;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int bi = a.tiles[2 * blockIdx.x], bj = a.tiles[2 * blockIdx.x + 1];
    for (int x = tid; x < (WIDE ? 2 : 1) * TILE * TILE; x += IS_THREADS) acc_lo[x] = 0u;
    if (tid == 0) {
        for (int s = 0; s < NST; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], IS_CWARPS);
        }
        mbar_fence_init();
    }
    __syncthreads();

    const unsigned char* ent_b = reinterpret_cast<const unsigned char*>(entries_v);
    unsigned nm = 0;  // matches this thread accumulated

    if (warp == IS_CWARPS) {
        // ------------------------------- producer -------------------------------
        int s = 0;
        uint32_t ph = 0;
        int64_t it = 0;
        for (int64_t ch = 0; ch <= nch; ++ch) {
            const bool last = ch == nch;
            int64_t r0 = 0, r1 = 0, c0 = 0, c1 = 0;
            const unsigned short* crun = nullptr;
            const unsigned short* rrun = nullptr;
            if (!last) {
                const int64_t rbk = (int64_t)bi * nch + ch, cbk = (int64_t)bj * nch + ch;
                r0 = bucket_ptr[rbk];
                r1 = bucket_ptr[rbk + 1];
                c0 = bucket_ptr[cbk];
                c1 = bucket_ptr[cbk + 1];
                if (r1 == r0 || c1 == c0) continue;
                crun = runs + (size_t)cbk * (CHUNK + 8);
                rrun = runs + (size_t)rbk * (CHUNK + 8);
            }
            int f0 = 0, ce0 = 0, re0 = 0;
            const int cn = (int)(c1 - c0), rn = (int)(r1 - r0);
            while (last || f0 < CHUNK) {
                int f1 = CHUNK, ce1 = cn, re1 = rn;
                if (!last && (cn - ce0 > CAP || rn - re0 > CAP)) {
                    // largest f1 in (f0, CHUNK], a multiple of 8, with at most CAP column entries and at
                    // most CAP row entries in [f0, f1): CHUNK / 32 position strides first, then 8-position
                    // strides (run starts are monotone)
                    constexpr int S1 = CHUNK / 32;
                    int base = f0;
                    {
                        const int cand = base + S1 * (lane + 1);
                        const bool ok = cand <= CHUNK && (int)crun[cand] - ce0 <= CAP && (int)rrun[cand] - re0 <= CAP;
                        base += S1 * __popc(__ballot_sync(0xffffffffu, ok));
                    }
                    {
                        const int cand = base + 8 * (lane + 1);
                        const bool ok = lane < S1 / 8 - 1 && cand <= CHUNK && (int)crun[cand] - ce0 <= CAP &&
                                        (int)rrun[cand] - re0 <= CAP;
                        base += 8 * __popc(__ballot_sync(0xffffffffu, ok));
                    }
                    f1 = base;   // > f0: eight positions hold at most 1024 entries <= CAP on either side
                    ce1 = (int)crun[f1];
                    re1 = (int)rrun[f1];
                }
                if (last || (ce1 > ce0 && re1 > re0)) {
                    if (it >= NST) mbar_wait_sleepy(&empty[s], ph ^ 1u);
                    if (lane == 0) {
                        unsigned char* st = stage0 + (size_t)s * STAGE;
                        IsHeader* h = reinterpret_cast<IsHeader*>(st + HDR);
                        h->done = last ? 1 : 0;
                        h->next = 0;
                        if (last) {
                            mbar_arrive(&full[s]);
                        } else {
                            // global index of slot 0 of either side, aligned down to a 16-byte unit
                            constexpr int64_t AL = Geo::align_entries;
                            const int64_t gc0 = (c0 + ce0) / AL * AL, gc1 = (c0 + ce1 + AL - 1) / AL * AL;
                            const int64_t gr0 = (r0 + re0) / AL * AL, gr1 = (r0 + re1 + AL - 1) / AL * AL;
                            const uint32_t cbytes = (uint32_t)(gc1 - gc0) * (uint32_t)Geo::esize;
                            const uint32_t wbytes = (uint32_t)(gr1 - gr0) * (uint32_t)Geo::esize;
                            const uint32_t rbytes = (uint32_t)(f1 - f0 + 8) * 2u;
                            const uint32_t lbytes = MODE == IS_UNWEIGHTED ? (uint32_t)(f1 - f0) * 8u : 0u;
                            h->nrow = re1 - re0;
                            h->roff = (int)(r0 + re0 - gr0);
                            h->coff = (int)(c0 - gc0);
                            h->f0 = f0;
                            mbar_arrive_expect_tx(&full[s], cbytes + wbytes + rbytes + lbytes);
                            bulk_g2s(st, ent_b + gc0 * Geo::esize, cbytes, &full[s]);
                            bulk_g2s(st + ENT, ent_b + gr0 * Geo::esize, wbytes, &full[s]);
                            bulk_g2s(st + RUNS, crun + f0, rbytes, &full[s]);
                            if (MODE == IS_UNWEIGHTED) bulk_g2s(st + LF, lenfx + ch * CHUNK + f0, lbytes, &full[s]);
                        }
                    }
                    __syncwarp();
                    ++it;
                    if (++s == NST) {
                        s = 0;
                        ph ^= 1u;
                    }
                }
                if (last) break;
                f0 = f1;
                ce0 = ce1;
                re0 = re1;
            }
        }
    } else {
        // ------------------------------- consumers -------------------------------
        // Everything a batch needs is in the stage: column entries, ROW entries, the column run starts
        // (and the lengths).  A warp claims SLOTS row entries at a time through the counter in the
        // header (the next claim is already in flight while it works), each group of G lanes takes one
        // of them and strides over the column run of its position.
        constexpr int SLOTS = 32 / G;
        const int slot = lane / G, g = lane % G;
        int s = 0;
        uint32_t ph = 0;
        for (;;) {
            mbar_wait_sleepy(&full[s], ph);
            const unsigned char* st = stage0 + (size_t)s * STAGE;
            const IsHeader* h = reinterpret_cast<const IsHeader*>(st + HDR);
            if (h->done) break;
            const unsigned short* fstart = reinterpret_cast<const unsigned short*>(st + RUNS);
            const int coff = h->coff, f0 = h->f0, nrow = h->nrow, roff = h->roff;
            int* next = const_cast<int*>(&h->next);
            auto grab = [&]() -> int {
                int b = 0;
                if (lane == 0) b = atomicAdd(next, SLOTS);
                return __shfl_sync(0xffffffffu, b, 0);
            };
            int b = grab();
            while (b < nrow) {
                const int bn = grab();   // the next claim travels while this one is worked on
                const int k = b + slot;
                if (k < nrow) {
                    if (MODE == IS_WEIGHTED) {
                        const TipEntry* centry = reinterpret_cast<const TipEntry*>(st);
                        const TipEntry re = reinterpret_cast<const TipEntry*>(st + ENT)[k + roff];
                        const uint32_t f = (re.key >> 7) - (uint32_t)f0, r = re.key & 127u;
                        const int e0 = (int)fstart[f] + coff, e1 = (int)fstart[f + 1] + coff;
                        const uint32_t rbase = r * TILE, rx = r & 31u;
                        const unsigned long long rv = (unsigned long long)re.val;
                        if (g == 0) nm += (unsigned)(e1 - e0);
#pragma unroll 2
                        for (int e = e0 + g; e < e1; e += G) {
                            const TipEntry ce = centry[e];
                            const unsigned long long cv = (unsigned long long)ce.val;
                            const unsigned long long add = rv < cv ? rv : cv;   // both non-negative, < 2^61
                            const uint32_t cell = rbase + ((ce.key & 127u) ^ rx);
                            const uint32_t alo = (uint32_t)add;
                            const uint32_t old = atomicAdd(&acc_lo[cell], alo);
                            atomicAdd(&acc_hi[cell], (uint32_t)(add >> 32) + (old + alo < old ? 1u : 0u));
                        }
                    } else if (MODE == IS_UNWEIGHTED) {
                        const uint32_t* ckey = reinterpret_cast<const uint32_t*>(st);
                        const uint32_t rk = reinterpret_cast<const uint32_t*>(st + ENT)[k + roff];
                        const uint32_t f = (rk >> 7) - (uint32_t)f0, r = rk & 127u;
                        const int e0 = (int)fstart[f] + coff, e1 = (int)fstart[f + 1] + coff;
                        const unsigned long long add = (unsigned long long)reinterpret_cast<const long long*>(st + LF)[f];
                        const uint32_t alo = (uint32_t)add, ahi = (uint32_t)(add >> 32);
                        const uint32_t rbase = r * TILE, rx = r & 31u;
                        if (g == 0) nm += (unsigned)(e1 - e0);
#pragma unroll 2
                        for (int e = e0 + g; e < e1; e += G) {
                            const uint32_t cell = rbase + ((ckey[e] & 127u) ^ rx);
                            const uint32_t old = atomicAdd(&acc_lo[cell], alo);
                            atomicAdd(&acc_hi[cell], ahi + (old + alo < old ? 1u : 0u));
                        }
                    } else {
                        const SpEntry* centry = reinterpret_cast<const SpEntry*>(st);
                        const SpEntry re = reinterpret_cast<const SpEntry*>(st + ENT)[k + roff];
                        const uint32_t f = (re.key >> 7) - (uint32_t)f0, r = re.key & 127u;
                        const int e0 = (int)fstart[f] + coff, e1 = (int)fstart[f + 1] + coff;
                        const uint32_t rbase = r * TILE, rx = r & 31u;
                        if (g == 0) nm += (unsigned)(e1 - e0);
                        for (int e = e0 + g; e < e1; e += G) {
                            const SpEntry ce = centry[e];
                            const uint32_t add = MODE == IS_FEAT_COUNT ? 1u : (uint32_t)(re.val < ce.val ? re.val : ce.val);
                            atomicAdd(&acc_lo[rbase + ((ce.key & 127u) ^ rx)], add);
                        }
                    }
                }
                b = bn;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
            if (++s == NST) {
                s = 0;
                ph ^= 1u;
            }
        }
    }
    __syncthreads();
    if (n_matches) {
#pragma unroll
        for (int o = 16; o; o >>= 1) nm += __shfl_xor_sync(0xffffffffu, nm, o);
        if (lane == 0 && nm) atomicAdd(n_matches, (unsigned long long)nm);
    }
    const int64_t base = a.row_base[bi];
    const int64_t i0 = (int64_t)bi * TILE;
    if (WIDE) {
        const double inv = scale[1];
        for (int x = tid; x < TILE * TILE; x += IS_THREADS) {
            const int rr = x / TILE, c = x % TILE;
            const int64_t i = i0 + rr, j = (int64_t)bj * TILE + c;
            if (i >= j || j >= a.n) continue;
            const int xs = rr * TILE + (c ^ (rr & 31));
            const long long m = (long long)(((unsigned long long)acc_hi[xs] << 32) | acc_lo[xs]);
            const long long t = Ufx[i] + Ufx[j] - 2 * m;
            emit(a, base, i0, i, j, (double)t * inv);
        }
    } else {
        for (int x = tid; x < TILE * TILE; x += IS_THREADS) {
            const int rr = x / TILE, c = x % TILE;
            emit(a, base, i0, i0 + rr, (int64_t)bj * TILE + c, (double)acc_lo[rr * TILE + (c ^ (rr & 31))]);
        }
    }
}

}  // namespace b2d
