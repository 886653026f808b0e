// Embedding kernels: CSR scatter + bottom-up tree propagation.
//
// Replaces _nodes_by_counts / _traverse_reduce
// (/root/reference/skbio/diversity/_phylogenetic.pyx:73-222): instead of a dense
// (n_nodes x n_samples) int64 matrix swept parent-by-parent, every sample (one lane) walks
// the tree once in the engine's heavy-child-first postorder with a tiny private stack
// (depth <= log2(n_nodes)+1, see b2d_tree.h), so the sweep is parallel over samples,
// sequential over nodes and independent of tree shape (a caterpillar costs the same as a
// balanced tree).  All integer sums are exact and order-independent -> bit-identical node
// counts.
#pragma once

#include "b2d_common.cuh"

namespace b2d {

constexpr int MAX_STACK = 64;

// Node flags (one byte per node position q)
constexpr uint8_t NF_HAS_CHILDREN = 1;  // value = own + pop()
constexpr uint8_t NF_FOLD = 2;          // fold value into the sibling accumulator (top += v)
constexpr uint8_t NF_PAD = 4;           // padding position, nothing to do

// ---------------------------------------------------------------------------------------
// Per-sample totals and normalisers from the CSR table (one warp per sample).
//   total[s]  = sum of counts placed on TIP nodes      (_unifrac.py:559-560: np.take(u, tip_indices).sum())
//   norm[s]   = sum_tips d_tip * (count/total)         (branch correction term of one sample,
//               _unifrac.py:594-619 splits as C_u + C_v because d is zero off the tips)
// ---------------------------------------------------------------------------------------
__global__ void k_sample_totals_tree(const int64_t* __restrict__ indptr,
                                     const int32_t* __restrict__ qidx,
                                     const int64_t* __restrict__ values,
                                     const uint8_t* __restrict__ flags,
                                     const double* __restrict__ tipdist_q,
                                     int64_t n, long long* __restrict__ total,
                                     double* __restrict__ norm) {
    int64_t s = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (s >= n) return;
    int64_t e0 = indptr[s], e1 = indptr[s + 1];
    long long t = 0;
    for (int64_t e = e0 + lane; e < e1; e += 32) {
        int32_t q = qidx[e];
        if (!(flags[q] & NF_HAS_CHILDREN)) t += values ? values[e] : 1;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    double c = 0.0;
    if (tipdist_q != nullptr) {
        double td = (double)t;
        for (int64_t e = e0 + lane; e < e1; e += 32) {
            int32_t q = qidx[e];
            if (!(flags[q] & NF_HAS_CHILDREN)) {
                double v = (double)(values ? values[e] : 1);
                double pr = (t > 0) ? v / td : v;
                c += tipdist_q[q] * pr;
            }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    }
    if (lane == 0) {
        total[s] = t;
        if (norm) norm[s] = c;
    }
}

// Feature tables: per-sample sum of values (Bray-Curtis denominator pieces), fp64.
__global__ void k_sample_sums_features(const int64_t* __restrict__ indptr,
                                       const double* __restrict__ values, int64_t n,
                                       double* __restrict__ sums) {
    int64_t s = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (s >= n) return;
    int64_t e0 = indptr[s], e1 = indptr[s + 1];
    double c = 0.0;
    for (int64_t e = e0 + lane; e < e1; e += 32) c += values ? values[e] : 1.0;
#pragma unroll
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) sums[s] = c;
}

// ---------------------------------------------------------------------------------------
// Scatter CSR entries into the (zeroed) embedding buffers.  One warp per sample.
// ---------------------------------------------------------------------------------------
// fp64/int64 layout, node range [q0, q1): slot = ((blk*QR + (q-q0))*TILE + sl)
template <typename V>
__global__ void k_scatter_dense(const int64_t* __restrict__ indptr,
                                const int32_t* __restrict__ qidx, const V* __restrict__ values,
                                int64_t n, int64_t q0, int64_t q1, int64_t QR,
                                unsigned long long* __restrict__ emb, int as_double) {
    int64_t s = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (s >= n) return;
    int64_t blk = s / TILE, sl = s % TILE;
    int64_t e0 = indptr[s], e1 = indptr[s + 1];
    for (int64_t e = e0 + lane; e < e1; e += 32) {
        int64_t q = qidx[e];
        if (q < q0 || q >= q1) continue;
        unsigned long long* slot = emb + ((blk * QR + (q - q0)) * TILE + sl);
        if (as_double) {
            double v = values ? (double)values[e] : 1.0;
            *reinterpret_cast<double*>(slot) = v;  // node ids are unique per sample
        } else {
            long long v = values ? (long long)values[e] : 1;
            *slot = (unsigned long long)v;
        }
    }
}

// bits layout: word = ((blk*NS + q/128)*TILE + sl)*4 + (q%128)/32, bit q%32
__global__ void k_scatter_bits(const int64_t* __restrict__ indptr,
                               const int32_t* __restrict__ qidx, int64_t n, int64_t NS,
                               uint32_t* __restrict__ bits) {
    int64_t s = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (s >= n) return;
    int64_t blk = s / TILE, sl = s % TILE;
    int64_t e0 = indptr[s], e1 = indptr[s + 1];
    for (int64_t e = e0 + lane; e < e1; e += 32) {
        int64_t q = qidx[e];
        int64_t w = ((blk * NS + (q >> 7)) * TILE + sl) * 4 + ((q & 127) >> 5);
        atomicOr(bits + w, 1u << (q & 31));
    }
}

// ---------------------------------------------------------------------------------------
// Weighted propagation: lane = sample, walks node positions [q0, q1) of the resident range.
// In : emb slots hold the int64 counts placed directly on each node (0 elsewhere).
// Out: emb slots hold the int64 count of the node's whole subtree (exact integer sums, so
//      bit-identical to _traverse_reduce whatever the order).
// The per-lane stack is carried across calls in `state` ([MAX_STACK][npad]); its depth is
// known to the host because it depends on the tree only.  The walk is a chain of dependent
// integer operations per lane, so everything that is not on that chain (the int64 -> fp64
// proportion conversion, one correctly rounded division per element) lives in the separate,
// bandwidth-bound kernel k_counts_to_proportions.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_propagate_weighted(unsigned long long* __restrict__ emb, const uint8_t* __restrict__ flags,
                     int64_t npad, int64_t q0, int64_t q1, int64_t QR,
                     long long* __restrict__ state, int depth0) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= npad) return;
    int64_t blk = s / TILE, sl = s % TILE;
    unsigned long long* p = emb + (blk * QR) * TILE + sl;
    long long S[MAX_STACK];
    int d = depth0;
    for (int i = 0; i < d; ++i) S[i] = state[(int64_t)i * npad + s];
    long long T = d > 0 ? S[d - 1] : 0;
    // Software pipeline: the counts of group g+PF are requested before group g is processed
    // (the buffer is updated in place, so the compiler cannot hoist these loads itself).
    // Streaming (evict-first) accesses keep L1 for the per-lane stack in local memory.
    constexpr int G = 8, PF = 3;
    long long own[PF + 1][G];
#pragma unroll
    for (int g = 0; g < PF; ++g)
#pragma unroll
        for (int u = 0; u < G; ++u)
            own[g][u] = (q0 + (int64_t)g * G < q1) ? (long long)__ldcs(p + ((int64_t)g * G + u) * TILE) : 0;
    for (int64_t qg = q0; qg < q1; qg += (PF + 1) * G) {
#pragma unroll
        for (int ph = 0; ph <= PF; ++ph) {
            const int64_t q = qg + (int64_t)ph * G;
            if (q >= q1) break;
            {   // prefetch group ph+PF (slot (ph+PF) % (PF+1))
                const int64_t qn = q + (int64_t)PF * G;
                if (qn < q1) {
#pragma unroll
                    for (int u = 0; u < G; ++u)
                        own[(ph + PF) % (PF + 1)][u] = (long long)__ldcs(p + (qn - q0 + u) * TILE);
                }
            }
            const unsigned long long f8 = *reinterpret_cast<const unsigned long long*>(flags + q);
#pragma unroll
            for (int u = 0; u < G; ++u) {
                const unsigned f = (unsigned)(f8 >> (8 * u)) & 0xffu;
                long long v = own[ph][u];
                if (f & NF_PAD) continue;
                if (f & NF_HAS_CHILDREN) {  // pop the children accumulator
                    v += T;
                    d -= 1;
                    if (d > 0) T = S[d - 1];
                    __stcs(p + (q - q0 + u) * TILE, (unsigned long long)v);  // tips keep their own count
                }
                if (f & NF_FOLD) {
                    T += v;
                } else {
                    if (d > 0) S[d - 1] = T;
                    T = v;
                    d += 1;
                }
            }
        }
    }
    if (d > 0) S[d - 1] = T;
    for (int i = 0; i < d; ++i) state[(int64_t)i * npad + s] = S[i];
}

// Subtree decomposition, phase 1: thread = (sample, span); walks one span of whole small
// subtrees with an empty stack.  At a subtree root (NF_CUT) the value is final and the stack
// is reset instead of pushed.  Plenty of independent warps -> no software prefetch needed.
constexpr uint8_t NF_CUT = 8;
__global__ void __launch_bounds__(128)
k_propagate_weighted_spans(unsigned long long* __restrict__ emb, const uint8_t* __restrict__ flags1,
                           const int64_t* __restrict__ spans, int64_t npad, int64_t QR) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= npad) return;
    const int64_t qa = spans[2 * (int64_t)blockIdx.y], qb = spans[2 * (int64_t)blockIdx.y + 1];
    const int64_t blk = s / TILE, sl = s % TILE;
    unsigned long long* p = emb + (blk * QR) * TILE + sl;
    long long S[MAX_STACK];
    int d = 0;
    long long T = 0;
    for (int64_t q = qa; q < qb; ++q) {
        const unsigned f = flags1[q];
        long long v = (long long)p[q * TILE];
        if (f & NF_HAS_CHILDREN) {
            v += T;
            d -= 1;
            if (d > 0) T = S[d - 1];
            p[q * TILE] = (unsigned long long)v;
        }
        if (f & NF_CUT) {
            d = 0;
        } else if (f & NF_FOLD) {
            T += v;
        } else {
            if (d > 0) S[d - 1] = T;
            T = v;
            d += 1;
        }
    }
}

// Phase 2: lane = sample, sequential over the (short) list of top positions; subtree roots
// computed in phase 1 enter as leaves.
__global__ void __launch_bounds__(128)
k_propagate_weighted_top(unsigned long long* __restrict__ emb, const int32_t* __restrict__ top_pos,
                         const uint8_t* __restrict__ top_flags, int64_t ntop, int64_t npad,
                         int64_t QR) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= npad) return;
    const int64_t blk = s / TILE, sl = s % TILE;
    unsigned long long* p = emb + (blk * QR) * TILE + sl;
    long long S[MAX_STACK];
    int d = 0;
    long long T = 0;
    constexpr int PF = 8;
    long long own[PF];
#pragma unroll
    for (int u = 0; u < PF; ++u) own[u] = u < ntop ? (long long)p[(int64_t)top_pos[u] * TILE] : 0;
    for (int64_t k0 = 0; k0 < ntop; k0 += PF) {
        long long cur[PF];
#pragma unroll
        for (int u = 0; u < PF; ++u) cur[u] = own[u];
#pragma unroll
        for (int u = 0; u < PF; ++u)
            own[u] = (k0 + PF + u < ntop) ? (long long)p[(int64_t)top_pos[k0 + PF + u] * TILE] : 0;
#pragma unroll
        for (int u = 0; u < PF; ++u) {
            const int64_t k = k0 + u;
            if (k >= ntop) break;
            const unsigned f = top_flags[k];
            long long v = cur[u];
            if (f & NF_HAS_CHILDREN) {
                v += T;
                d -= 1;
                if (d > 0) T = S[d - 1];
                p[(int64_t)top_pos[k] * TILE] = (unsigned long long)v;
            }
            if (f & NF_FOLD) {
                T += v;
            } else {
                if (d > 0) S[d - 1] = T;
                T = v;
                d += 1;
            }
        }
    }
}

// Generalised phylogenetic diversity (_phydiv, /root/reference/skbio/diversity/alpha/_pd.py:190-239)
// from int64 subtree counts: per sample  sum_q len_q * g(count_q)  with
//   unweighted: g = [count > 0] (and count < total when unrooted)
//   weighted  : f = count/total; unrooted: f = 2*min(f, 1-f); theta < 1: f = f^theta; g = f
// thread = sample, grid.y = chunks of 2048 node rows -> partial[chunk][sample]; summed in chunk
// order afterwards (two-level, deterministic).
__global__ void __launch_bounds__(128)
k_phydiv_partial(const unsigned long long* __restrict__ emb, const double* __restrict__ len_rel,
                 const long long* __restrict__ total, int64_t npad, int64_t rows, int64_t QR,
                 int rooted, int weighted, double theta, double* __restrict__ partial,
                 int64_t chunk0) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= npad) return;
    const int64_t blk = s / TILE, sl = s % TILE;
    const unsigned long long* p = emb + (blk * QR) * TILE + sl;
    const long long tot = total[s];
    const double totd = (double)tot;
    const int64_t qa = (int64_t)blockIdx.y * 2048, qb = (qa + 2048 < rows) ? qa + 2048 : rows;
    double acc = 0.0;
    if (tot > 0) {
        for (int64_t q = qa; q < qb; ++q) {
            const long long c = (long long)p[q * TILE];
            if (c <= 0) continue;
            if (!weighted) {
                if (rooted || c < tot) acc += len_rel[q];
            } else {
                double f = (double)c / totd;
                if (!rooted) f = 2.0 * fmin(f, 1.0 - f);
                if (theta < 1.0) f = pow(f, theta);
                acc += len_rel[q] * f;
            }
        }
    }
    partial[(chunk0 + blockIdx.y) * npad + s] = acc;
}

__global__ void k_phydiv_reduce(const double* __restrict__ partial, int64_t nchunks, int64_t npad,
                                double* __restrict__ out) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= npad) return;
    double v = 0.0;
    for (int64_t c = 0; c < nchunks; ++c) v += partial[c * npad + s];
    out[s] = v;
}

// every count of the sample, wherever it was placed (= the root's subtree count)
__global__ void k_sample_totals_all(const int64_t* __restrict__ indptr, const int64_t* __restrict__ values,
                                    int64_t n, long long* __restrict__ total) {
    int64_t s = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (s >= n) return;
    long long t = 0;
    for (int64_t e = indptr[s] + lane; e < indptr[s + 1]; e += 32) t += values ? values[e] : 1;
#pragma unroll
    for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) total[s] = t;
}

// int64 subtree counts -> fp64 proportions count/total (the count itself when total == 0,
// as _weighted_unifrac does, beta/_unifrac.py:399-410).  Elementwise, in place, two samples
// (16 bytes) per thread; HBM-bound.
__global__ void __launch_bounds__(256)
k_counts_to_proportions(unsigned long long* __restrict__ emb, const long long* __restrict__ total,
                        int64_t nb, int64_t rows, int64_t QR) {
    // grid.x covers (block, sample pair): nb * 64 threads ; grid.y strides over node rows
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nb * (TILE / 2)) return;
    int64_t blk = t / (TILE / 2), sp = t % (TILE / 2);
    const long long t0 = total[blk * TILE + 2 * sp], t1 = total[blk * TILE + 2 * sp + 1];
    const double d0 = (double)t0, d1 = (double)t1;
    ulonglong2* p = reinterpret_cast<ulonglong2*>(emb + (blk * QR) * TILE) + sp;
    for (int64_t q = blockIdx.y; q < rows; q += gridDim.y) {
        ulonglong2 v = p[q * (TILE / 2)];
        double a = (double)(long long)v.x, b = (double)(long long)v.y;
        if (t0 > 0) a = a / d0;
        if (t1 > 0) b = b / d1;
        double2 o = make_double2(a, b);
        *reinterpret_cast<double2*>(p + q * (TILE / 2)) = o;
    }
}

// ---------------------------------------------------------------------------------------
// Presence propagation on node-packed bits: lane = sample, 128 nodes (one uint4) per step.
// The private stack is a 64-bit shift register (1 bit per level).  Also accumulates
// Lsum[s] = sum_q len[q] * present(q, s) with two-level fp64 summation (per stripe, then
// across stripes) - this is Faith's PD of the sample and the per-sample part of the
// unweighted UniFrac denominator:  [u|v] = ([u] + [v] + [u^v]) / 2.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TILE)
k_propagate_bits(uint4* __restrict__ bits, const uint32_t* __restrict__ hc_mask,
                 const uint32_t* __restrict__ fold_mask, const double* __restrict__ len_q,
                 int64_t NS, double* __restrict__ Lsum) {
    __shared__ double s_len[NODE_STRIPE];
    __shared__ uint32_t s_hc[4], s_fo[4];
    int64_t blk = blockIdx.x;
    int sl = threadIdx.x;
    uint4* p = bits + (blk * NS) * TILE + sl;
    unsigned long long S = 0;
    double Ltot = 0.0;
    for (int64_t st = 0; st < NS; ++st) {
        __syncthreads();
        s_len[sl] = len_q[st * NODE_STRIPE + sl];
        if (sl < 4) {
            s_hc[sl] = hc_mask[st * 4 + sl];
            s_fo[sl] = fold_mask[st * 4 + sl];
        }
        __syncthreads();
        uint4 own = p[st * TILE];
        uint32_t w[4] = {own.x, own.y, own.z, own.w};
        double Lacc = 0.0;
#pragma unroll
        for (int wi = 0; wi < 4; ++wi) {
            uint32_t word = w[wi], res = 0;
            uint32_t hc = s_hc[wi], fo = s_fo[wi];
#pragma unroll 8
            for (int b = 0; b < 32; ++b) {
                uint32_t o = (word >> b) & 1u;
                if ((hc >> b) & 1u) {
                    o |= (uint32_t)(S & 1ull);
                    S >>= 1;
                }
                if ((fo >> b) & 1u)
                    S |= (unsigned long long)o;
                else
                    S = (S << 1) | (unsigned long long)o;
                res |= o << b;
                if (o) Lacc += s_len[wi * 32 + b];
            }
            w[wi] = res;
        }
        p[st * TILE] = make_uint4(w[0], w[1], w[2], w[3]);
        Ltot += Lacc;
    }
    Lsum[blk * TILE + sl] = Ltot;
}

}  // namespace b2d
