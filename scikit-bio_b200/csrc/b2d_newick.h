// Native Newick -> flat arrays (host only).  Produces the tree directly in the reference's node
// id order (TreeNode.assign_ids, /root/reference/skbio/tree/_tree.py:5337-5359) without ever
// building node objects; behaviour follows the reference reader
// (/root/reference/skbio/io/format/newick.py:270-336 and its tokenizer :380-470): nested
// [comments] are skipped, 'quoted labels' are literal with '' for a quote, underscores in
// unquoted labels become spaces, whitespace outside quotes is ignored between tokens and
// rejected inside an unquoted label, an empty label is "no name", ':' introduces a length.
#pragma once
#include <cstring>

#include <stdint.h>
#include <stdlib.h>

#include <charconv>
#include <cmath>
#include <string>
#include <string_view>
#include <unordered_map>
#include <vector>

namespace b2d {

struct FlatTree {
    std::vector<int32_t> parent;    // reference id order, root last, -1 for the root
    std::vector<double> length;     // NaN = no length given
    std::vector<int64_t> name_off;  // [n+1] offsets into name_buf
    std::vector<uint8_t> has_name;
    std::string name_buf;
};

inline std::string parse_newick(const char* s, int64_t n, FlatTree& out) {
    struct Node {
        int32_t parent;
        int32_t first_child, last_child, next_sibling;
        double length;
        int64_t name_b, name_e;  // into names
        bool has_name;
    };
    std::vector<Node> nodes;
    std::string names;
    nodes.reserve((size_t)(n / 12 + 16));
    names.reserve((size_t)(n / 4 + 16));
    auto new_node = [&](int32_t parent) {
        Node nd{parent, -1, -1, -1, std::nan(""), 0, 0, false};
        nodes.push_back(nd);
        int32_t id = (int32_t)nodes.size() - 1;
        if (parent >= 0) {
            if (nodes[parent].first_child < 0)
                nodes[parent].first_child = id;
            else
                nodes[nodes[parent].last_child].next_sibling = id;
            nodes[parent].last_child = id;
        }
        return id;
    };
    int32_t cur = new_node(-1);
    std::vector<int32_t> stack;  // open parents
    int64_t i = 0;
    bool done = false;
    bool closed_children = false;  // cur got its children by a ')' (no more '(' allowed for it)
    auto skip_ws_comments = [&]() -> bool {
        while (i < n) {
            char c = s[i];
            if (c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\f' || c == '\v') {
                ++i;
            } else if (c == '[') {
                int depth = 0;
                while (i < n) {
                    if (s[i] == '[') ++depth;
                    else if (s[i] == ']') { --depth; if (depth == 0) { ++i; break; } }
                    ++i;
                }
                if (depth != 0) return false;
            } else {
                break;
            }
        }
        return true;
    };
    // reads a label or a number token (quoted or not); returns false on malformed input
    auto read_token = [&](std::string& tok, bool& quoted) -> bool {
        tok.clear();
        quoted = false;
        if (i < n && s[i] == '\'') {
            quoted = true;
            ++i;
            while (i < n) {
                if (s[i] == '\'') {
                    if (i + 1 < n && s[i + 1] == '\'') { tok.push_back('\''); i += 2; continue; }
                    ++i;
                    return true;
                }
                tok.push_back(s[i++]);
            }
            return false;  // unbalanced quote
        }
        bool seen_ws = false;
        while (i < n) {
            char c = s[i];
            if (c == '(' || c == ')' || c == ',' || c == ':' || c == ';' || c == '\'') break;
            if (c == '[') { if (!skip_ws_comments()) return false; continue; }
            if (c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\f' || c == '\v') {
                seen_ws = true;
                ++i;
                continue;
            }
            if (seen_ws && !tok.empty()) return false;  // unescaped whitespace inside a label
            tok.push_back(c);
            ++i;
        }
        return true;
    };
    // Fast path for the overwhelmingly common token: unquoted, no whitespace, no comment.  kind[c]: 1 = a
    // structural character ends the token, 2 = whitespace / '[' / quote: the general reader takes over.
    static const struct Kinds {
        unsigned char k[256];
        Kinds() {
            memset(k, 0, sizeof(k));
            for (const char* d = "(),:;"; *d; ++d) k[(unsigned char)*d] = 1;
            for (const char* d = " \t\n\r\f\v['"; *d; ++d) k[(unsigned char)*d] = 2;
        }
    } kinds;
    auto plain_token_end = [&](int64_t from) -> int64_t {   // end of a plain token starting at `from`, or -1
        int64_t j = from;
        while (j < n && kinds.k[(unsigned char)s[j]] == 0) ++j;
        return (j >= n || kinds.k[(unsigned char)s[j]] == 1) ? j : -1;
    };
    while (!done) {
        if (i < n && kinds.k[(unsigned char)s[i]] == 2 && !skip_ws_comments()) return "unbalanced [comment] in Newick text";
        if (i >= n) break;
        char c = s[i];
        if (c == '(') {
            if (closed_children || nodes[cur].first_child >= 0 && false) return "could not parse Newick: unnested children";
            stack.push_back(cur);
            cur = new_node(cur);
            closed_children = false;
            ++i;
        } else if (c == ',') {
            if (stack.empty()) return "could not parse Newick: ',' outside parentheses";
            cur = new_node(stack.back());
            closed_children = false;
            ++i;
        } else if (c == ')') {
            if (stack.empty()) return "could not parse Newick: parentheses are unbalanced";
            cur = stack.back();
            stack.pop_back();
            closed_children = true;
            ++i;
        } else if (c == ';') {
            done = true;
            ++i;
        } else if (c == ':') {
            ++i;
            {
                const int64_t j = plain_token_end(i);
                double fv = 0.0;
                if (j > i) {
                    const auto ff = std::from_chars(s + i, s + j, fv);
                    if (ff.ec == std::errc() && ff.ptr == s + j) {
                        nodes[cur].length = fv;
                        i = j;
                        continue;
                    }
                }
            }
            if (!skip_ws_comments()) return "unbalanced [comment] in Newick text";
            std::string tok;
            bool quoted;
            if (!read_token(tok, quoted)) return "could not parse Newick: malformed length";
            // plain decimal / exponent forms go through from_chars (several times faster than
            // strtod, correctly rounded as well); anything it does not take whole (a leading '+',
            // "inf", hex) gets strtod's verdict
            double v = 0.0;
            const char* tb = tok.data();
            const char* te = tb + tok.size();
            const auto fc = std::from_chars(tb, te, v);
            if (tok.empty() || fc.ec != std::errc() || fc.ptr != te) {
                char* endp = nullptr;
                v = strtod(tok.c_str(), &endp);
                if (tok.empty() || endp == tok.c_str() || *endp != '\0')
                    return "Could not read length as numeric type: " + tok + ".";
            }
            nodes[cur].length = v;
        } else {
            {
                const int64_t j = plain_token_end(i);
                if (j > i) {   // a plain label: straight into the name blob, '_' -> ' '
                    nodes[cur].has_name = true;
                    nodes[cur].name_b = (int64_t)names.size();
                    names.append(s + i, (size_t)(j - i));
                    for (int64_t x = nodes[cur].name_b; x < (int64_t)names.size(); ++x)
                        if (names[x] == '_') names[x] = ' ';
                    nodes[cur].name_e = (int64_t)names.size();
                    i = j;
                    continue;
                }
            }
            std::string tok;
            bool quoted;
            if (!read_token(tok, quoted)) return "could not parse Newick: unbalanced quotes or whitespace in a label";
            if (!quoted)
                for (auto& ch : tok)
                    if (ch == '_') ch = ' ';
            if (!tok.empty()) {
                nodes[cur].has_name = true;
                nodes[cur].name_b = (int64_t)names.size();
                names += tok;
                nodes[cur].name_e = (int64_t)names.size();
            }
        }
    }
    if (!done || !stack.empty())
        return "Could not parse file as newick. `(Parenthesis)`, `'single-quotes'`, `[comments]` may be "
               "unbalanced, or tree may be missing its root.";
    // reference ids: walk in postorder; when a node is reached its children get consecutive ids
    const int64_t m = (int64_t)nodes.size();
    std::vector<int32_t> new_id(m, -1);
    int32_t counter = 0;
    {
        std::vector<std::pair<int32_t, int32_t>> st;  // (node, next child to descend into)
        st.push_back({0, nodes[0].first_child});
        while (!st.empty()) {
            auto& top = st.back();
            if (top.second >= 0) {
                int32_t ch = top.second;
                top.second = nodes[ch].next_sibling;
                st.push_back({ch, nodes[ch].first_child});
            } else {
                int32_t v = top.first;
                st.pop_back();
                for (int32_t ch = nodes[v].first_child; ch >= 0; ch = nodes[ch].next_sibling) new_id[ch] = counter++;
            }
        }
        new_id[0] = counter;
    }
    out.parent.assign(m, -1);
    out.length.assign(m, std::nan(""));
    out.has_name.assign(m, 0);
    std::vector<int32_t> old_of(m);
    for (int64_t v = 0; v < m; ++v) old_of[new_id[v]] = (int32_t)v;
    out.name_off.assign(m + 1, 0);
    out.name_buf.clear();
    for (int64_t k = 0; k < m; ++k) {
        const Node& nd = nodes[old_of[k]];
        out.parent[k] = nd.parent >= 0 ? new_id[nd.parent] : -1;
        out.length[k] = nd.length;
        out.has_name[k] = nd.has_name ? 1 : 0;
        out.name_off[k] = (int64_t)out.name_buf.size();
        if (nd.has_name) out.name_buf.append(names, (size_t)nd.name_b, (size_t)(nd.name_e - nd.name_b));
    }
    out.name_off[m] = (int64_t)out.name_buf.size();
    return "";
}

// Node names -> node id: open addressing over the name blob itself (no per-name allocation; a
// std::unordered_map of 200 000 string_views costs three times the probe loop below).
class NameIndex {
 public:
    NameIndex(const char* buf, const int64_t* off, size_t expected) : buf_(buf), off_(off) {
        size_t cap = 16;
        while (cap < expected * 2 + 2) cap <<= 1;
        mask_ = cap - 1;
        ids_.assign(cap, -1);
    }
    static uint64_t hash(const char* p, size_t n) {
        uint64_t h = 0xcbf29ce484222325ull;   // FNV-1a, folded
        for (size_t i = 0; i < n; ++i) {
            h ^= (unsigned char)p[i];
            h *= 0x100000001b3ull;
        }
        return h ^ (h >> 29);
    }
    // the node stored under v's name before the call (-1: none); v is stored when there was none or `overwrite`
    int32_t put(int32_t v, bool overwrite) {
        const char* p = buf_ + off_[v];
        const size_t n = (size_t)(off_[v + 1] - off_[v]);
        for (size_t s = hash(p, n) & mask_;; s = (s + 1) & mask_) {
            const int32_t u = ids_[s];
            if (u < 0) {
                ids_[s] = v;
                return -1;
            }
            if ((size_t)(off_[u + 1] - off_[u]) == n && memcmp(buf_ + off_[u], p, n) == 0) {
                if (overwrite) ids_[s] = v;
                return u;
            }
        }
    }
    int32_t find(const char* p, size_t n) const {
        for (size_t s = hash(p, n) & mask_;; s = (s + 1) & mask_) {
            const int32_t u = ids_[s];
            if (u < 0) return -1;
            if ((size_t)(off_[u + 1] - off_[u]) == n && memcmp(buf_ + off_[u], p, n) == 0) return u;
        }
    }

 private:
    const char* buf_;
    const int64_t* off_;
    size_t mask_;
    std::vector<int32_t> ids_;
};

// Name -> node id for a list of query names, with the reference's lookup semantics
// (/root/reference/skbio/diversity/_phylogenetic.pyx:195-199): every node name is scanned in id order,
// tips and internal nodes alike, and a later node with the same name overwrites an earlier one.
// out[k] = node id or -1.  Names and queries are byte blobs with offsets; query entries may be
// followed by `query_sep` separator bytes that are not part of the name.
inline void lookup_names(const char* name_buf, const int64_t* name_off, const uint8_t* has_name, int64_t m,
                         const char* query_buf, const int64_t* query_off, int64_t k, int query_sep, int32_t* out) {
    NameIndex map(name_buf, name_off, (size_t)m);
    for (int64_t v = 0; v < m; ++v)
        if (has_name[v]) map.put((int32_t)v, true);
    for (int64_t q = 0; q < k; ++q)
        out[q] = map.find(query_buf + query_off[q], (size_t)(query_off[q + 1] - query_off[q] - query_sep));
}

// The name checks of _validate_taxa_and_tree (/root/reference/skbio/tree/_utils.py:83-90) on flat
// arrays: *dup_tips = some tip name occurs on two tips; *n_missing = distinct query names that are
// not tip names.  (An unnamed tip counts as the name None, as in the reference's list of tip names.)
inline void validate_names(const char* name_buf, const int64_t* name_off, const uint8_t* has_name,
                           const int32_t* parent, int64_t m, const char* query_buf, const int64_t* query_off,
                           int64_t k, int query_sep, int* dup_tips, int64_t* n_missing) {
    std::vector<uint8_t> has_child((size_t)m, 0);
    int64_t ntips = m;
    for (int64_t v = 0; v < m; ++v)
        if (parent[v] >= 0 && !has_child[parent[v]]) {
            has_child[parent[v]] = 1;
            --ntips;
        }
    NameIndex tips(name_buf, name_off, (size_t)std::max<int64_t>(ntips, 1));
    int dup = 0;
    int64_t unnamed = 0;
    for (int64_t v = 0; v < m; ++v) {
        if (has_child[v] || parent[v] < 0) continue;   // postorder(include_self=False): never the root
        if (!has_name[v]) {
            if (++unnamed > 1) dup = 1;
            continue;
        }
        if (tips.put((int32_t)v, false) >= 0) dup = 1;
    }
    int64_t missing = 0;
    std::unordered_map<std::string_view, int32_t> seen;   // queries that are no tip name: normally none
    for (int64_t q = 0; q < k; ++q) {
        const char* p = query_buf + query_off[q];
        const size_t n = (size_t)(query_off[q + 1] - query_off[q] - query_sep);
        if (tips.find(p, n) < 0 && seen.emplace(std::string_view(p, n), 0).second) ++missing;
    }
    *dup_tips = dup;
    *n_missing = missing;
}

}  // namespace b2d
