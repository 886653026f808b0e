// Host-side tree preparation (plain C++, no CUDA): from the reference's flat tree
// (parent[k] > k, /root/reference/skbio/tree/_tree.py:5337-5359) to the engine's own node
// order and the per-node step descriptors the propagation kernels execute.
//
// Engine order "q": postorder in which the child with the LARGEST subtree is visited first.
// While a lane sweeps nodes in that order it keeps one accumulator per ancestor that already
// has a finished child; every later child is at most half its parent's subtree, so at most
// floor(log2(m)) + 1 accumulators are live - a 64-entry stack is enough for any tree, deep
// caterpillars included.
#pragma once

#include <stdint.h>

#include <algorithm>
#include <string>
#include <vector>

namespace b2d {

struct TreeOrder {
    int64_t m = 0, mpad = 0;
    std::vector<int32_t> q_of_id;   // node id -> position q
    std::vector<uint8_t> flags;     // [mpad] NF_* per position
    std::vector<uint32_t> hc_mask;  // [mpad/32] bit = has children
    std::vector<uint32_t> fold_mask;// [mpad/32] bit = fold into sibling accumulator
    std::vector<double> len_q;      // [mpad] branch length per position (0 in the padding)
    std::vector<double> tipdist_q;  // [mpad] tip-to-root distance on tips, 0 elsewhere
    std::vector<int32_t> depth_at;  // [mpad/128 + 1] stack depth before each 128-node stripe
    int max_depth = 0;
    // per position, for the closed-form part of zero-block skipping (b2d_skip.cuh): parent
    // position (-1: root / padding), subtree size (a subtree is the position range
    // [q - size + 1, q]) and the root distance sum(len) over q .. root as an unevaluated
    // double-double sum hi + lo, so that differences along a path lose nothing to cancellation
    std::vector<int32_t> parent_q, size_q, tips_q;  // tips_q: tips in the subtree (0 in the padding)
    std::vector<double> rd_hi, rd_lo;
    // the same with the tips' own branches left out (a tip gets its parent's distance): used
    // when the tip rows are handled by the sparse intersection kernel (b2d_tips.cuh)
    std::vector<double> rd_hi_int, rd_lo_int;
    // Subtree decomposition of the weighted walk (used when the whole tree is resident):
    // phase 1 walks every maximal subtree of <= cut_size nodes independently (one span of
    // consecutive positions each, empty stack), phase 2 walks only the remaining "top" nodes
    // with those subtree roots acting as leaves.
    bool decomposed = false;
    int64_t cut_size = 0;
    std::vector<uint8_t> flags1;     // [mpad] phase-1 flags: flags | NF_CUT (8) on subtree roots
    std::vector<int64_t> spans;      // [2 * n_spans] {first position, end position}
    std::vector<int32_t> top_pos;    // phase-2 positions, ascending
    std::vector<uint8_t> top_flags;  // phase-2 flags (cut roots: no HAS_CHILDREN)
};

// Returns "" on success, else an error message.
inline std::string build_tree_order(int64_t m, const int32_t* parent, const double* length,
                                    TreeOrder& out) {
    if (m <= 0) return "tree has no nodes";
    int64_t root = -1;
    for (int64_t k = 0; k < m; ++k) {
        int32_t p = parent[k];
        if (p < 0) {
            if (root >= 0) return "tree has more than one root (parent == -1)";
            root = k;
        } else if (p <= k || p >= m) {
            return "parent[k] must satisfy k < parent[k] < n_nodes (reference id order)";
        }
    }
    if (root != m - 1) return "the root must be the last node";
    // children lists (CSR), siblings in ascending id order
    std::vector<int64_t> coff(m + 1, 0);
    for (int64_t k = 0; k < m; ++k)
        if (parent[k] >= 0) coff[parent[k] + 1]++;
    for (int64_t k = 0; k < m; ++k) coff[k + 1] += coff[k];
    std::vector<int32_t> child(m > 1 ? m - 1 : 0);
    {
        std::vector<int64_t> fill(coff.begin(), coff.end() - 1);
        for (int64_t k = 0; k < m; ++k)
            if (parent[k] >= 0) child[fill[parent[k]]++] = (int32_t)k;
    }
    // subtree sizes: children precede parents in id order
    std::vector<int64_t> size(m, 1);
    for (int64_t k = 0; k < m; ++k)
        if (parent[k] >= 0) size[parent[k]] += size[k];
    // heavy child first (stable for ties)
    for (int64_t k = 0; k < m; ++k) {
        int64_t b = coff[k], e = coff[k + 1];
        if (e - b == 2) {  // the usual case: no temporary buffer, no call
            if (size[child[b + 1]] > size[child[b]]) std::swap(child[b], child[b + 1]);
        } else if (e - b > 2 && e - b <= 16) {  // stable insertion sort, descending size
            for (int64_t i = b + 1; i < e; ++i) {
                const int32_t x = child[i];
                int64_t j = i;
                while (j > b && size[child[j - 1]] < size[x]) {
                    child[j] = child[j - 1];
                    --j;
                }
                child[j] = x;
            }
        } else if (e - b > 16) {
            std::stable_sort(child.begin() + b, child.begin() + e,
                             [&](int32_t x, int32_t y) { return size[x] > size[y]; });
        }
    }
    out.m = m;
    out.mpad = (m + 127) / 128 * 128;
    out.q_of_id.assign(m, -1);
    out.flags.assign(out.mpad, 4 /*NF_PAD*/);
    out.len_q.assign(out.mpad, 0.0);
    out.tipdist_q.assign(out.mpad, 0.0);
    out.hc_mask.assign(out.mpad / 32, 0u);
    out.fold_mask.assign(out.mpad / 32, 0xffffffffu);
    out.depth_at.assign(out.mpad / 128 + 1, 0);
    // One iterative postorder (heavy child first) fills everything that is stored per position,
    // in position order.  A frame carries what a node inherits from its parent, computed when the
    // node is pushed:
    //   * the tip-to-root distance exactly as _tip_distances accumulates it
    //     (/root/reference/skbio/diversity/_phylogenetic.pyx:52-56): d[i] = len[i] + d[parent];
    //   * the root distance sum(len) over node .. root as a double-double (TwoSum).
    out.parent_q.assign(out.mpad, -1);
    out.size_q.assign(out.mpad, 1);
    out.tips_q.assign(out.mpad, 0);
    out.rd_hi.assign(out.mpad, 0.0);
    out.rd_lo.assign(out.mpad, 0.0);
    out.rd_hi_int.assign(out.mpad, 0.0);
    out.rd_lo_int.assign(out.mpad, 0.0);
    struct Frame {
        int64_t v, next;
        double d, hi, lo;
        int32_t tips;
    };
    std::vector<Frame> st;
    st.reserve(1024);
    st.push_back({root, 0, length[root], length[root], 0.0, 0});
    int64_t q = 0;
    int depth = 0, max_depth = 0;
    while (!st.empty()) {
        Frame& fr = st.back();
        const int64_t v = fr.v;
        const int64_t b = coff[v], e = coff[v + 1];
        if (b + fr.next < e) {
            const int64_t c = child[b + fr.next];
            fr.next += 1;
            const double a = fr.hi, l = length[c];
            const double sum = a + l, bb = sum - a;
            const double err = (a - (sum - bb)) + (l - bb);  // TwoSum
            const double l2 = fr.lo + err;
            const double h = sum + l2;
            const double dch = l + fr.d;
            st.push_back({c, 0, dch, h, l2 - (h - sum), 0});  // invalidates fr
            continue;
        }
        const bool has_children = e > b;
        const int32_t tips = has_children ? fr.tips : 1;
        const double dv = fr.d, hi = fr.hi, lo = fr.lo;
        st.pop_back();
        double phi = 0.0, plo = 0.0;
        bool fold = false;
        if (!st.empty()) {
            Frame& pf = st.back();
            pf.tips += tips;
            phi = pf.hi;
            plo = pf.lo;
            fold = child[coff[pf.v]] != (int32_t)v;  // not the first visited
        }
        if (q % 128 == 0) out.depth_at[q / 128] = depth;
        out.q_of_id[v] = (int32_t)q;
        uint8_t f = 0;
        if (has_children) {
            f |= 1;
            depth -= 1;
            out.hc_mask[q >> 5] |= 1u << (q & 31);
            for (int64_t i = b; i < e; ++i) out.parent_q[out.q_of_id[child[i]]] = (int32_t)q;
        }
        if (fold) {
            f |= 2;
        } else {
            depth += 1;
            out.fold_mask[q >> 5] &= ~(1u << (q & 31));
        }
        if (depth > max_depth) max_depth = depth;
        out.flags[q] = f;
        out.len_q[q] = length[v];
        out.tipdist_q[q] = has_children ? 0.0 : dv;
        out.size_q[q] = (int32_t)size[v];
        out.tips_q[q] = tips;
        out.rd_hi[q] = hi;
        out.rd_lo[q] = lo;
        out.rd_hi_int[q] = has_children ? hi : phi;  // a tip carries its parent's distance
        out.rd_lo_int[q] = has_children ? lo : plo;
        ++q;
    }
    for (int64_t s = (m + 127) / 128; s <= out.mpad / 128; ++s) out.depth_at[s] = depth;
    if (m % 128 == 0) out.depth_at[m / 128] = depth;
    out.max_depth = max_depth;
    if (max_depth > 60) return "internal error: propagation stack deeper than 60";
    // ---- subtree decomposition ----
    {
        const int64_t B = std::min<int64_t>(4096, std::max<int64_t>(64, m / 256));
        out.cut_size = B;
        out.flags1 = out.flags;
        std::vector<int64_t> spans;
        std::vector<int32_t> top_pos;
        std::vector<uint8_t> top_flags;
        // in q order a subtree rooted at position p with s nodes occupies [p - s + 1, p]
        for (int64_t v = 0; v < m; ++v) {
            const int64_t pq = out.q_of_id[v];
            const bool big = size[v] > B;
            const bool parent_big = parent[v] < 0 || size[parent[v]] > B;
            if (big) {
                top_pos.push_back((int32_t)pq);
            } else if (parent_big) {  // root of a maximal small subtree
                top_pos.push_back((int32_t)pq);
                out.flags1[pq] |= 8;
                if (size[v] >= 2) {
                    spans.push_back(pq - size[v] + 1);
                    spans.push_back(pq + 1);
                }
            }
        }
        std::sort(top_pos.begin(), top_pos.end());
        top_flags.resize(top_pos.size());
        for (size_t k = 0; k < top_pos.size(); ++k) {
            uint8_t f = out.flags[top_pos[k]];
            if (out.flags1[top_pos[k]] & 8) f &= (uint8_t)~1u;  // already summed: acts as a leaf
            top_flags[k] = f;
        }
        // sort spans by start and merge neighbours that touch (no top node in between)
        std::vector<std::pair<int64_t, int64_t>> sp;
        for (size_t k = 0; k + 1 < spans.size(); k += 2) sp.push_back({spans[k], spans[k + 1]});
        std::sort(sp.begin(), sp.end());
        std::vector<int64_t> merged;
        for (auto& x : sp) {
            if (!merged.empty() && merged.back() == x.first &&
                x.second - merged[merged.size() - 2] <= 2 * B)
                merged.back() = x.second;
            else {
                merged.push_back(x.first);
                merged.push_back(x.second);
            }
        }
        out.spans.swap(merged);
        out.top_pos.swap(top_pos);
        out.top_flags.swap(top_flags);
        // worthwhile only if the sequential part shrinks a lot (not for caterpillars)
        out.decomposed = m >= 4096 && (int64_t)out.top_pos.size() * 8 <= m &&
                         (int64_t)out.spans.size() / 2 <= 60000;
    }
    return "";
}

// snake (boustrophedon) dealing of row stripes to shards: balanced triangle work
inline int stripe_owner(int64_t b, int n_shards) {
    int64_t r = b % (2 * (int64_t)n_shards);
    return (int)(r < n_shards ? r : 2 * n_shards - 1 - r);
}

}  // namespace b2d
