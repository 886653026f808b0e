/*
 * skbio_b200.h - C ABI of the B200-native all-pairs beta-diversity engine (libskbb200.so).
 *
 * This is the drop-in boundary for ONE path of scikit-bio:
 *
 *     skbio.diversity.beta_diversity(metric, counts, ids, tree=..., taxa=...)
 *         /root/reference/skbio/diversity/_driver.py:231-355
 *
 * for metric in {unweighted_unifrac, weighted_unifrac (normalized or not), braycurtis,
 * jaccard}.  The Python host (scikit-bio_b200/_driver.py) keeps the reference's signature,
 * validation and error behaviour, flattens the TreeNode into parent/length arrays, turns the
 * counts table into CSR and calls the functions below through ctypes - the same plugin
 * convention the reference already uses for its optional accelerator library
 * (ctypes.CDLL("libskbb.so"), /root/reference/skbio/binaries/_util.py:28-67; entry points
 * taking raw caller-owned buffers, /root/reference/skbio/binaries/_distance.py:161-193).
 * Unlike libskbb, every call returns an int status and b2d_last_error() explains a failure.
 *
 * Conventions
 *   - plain C, no torch / CUDA types in any signature; all pointers are HOST pointers to
 *     caller-owned, C-contiguous buffers; the library never frees caller memory;
 *   - nodes are numbered as TreeNode.assign_ids numbers them
 *     (/root/reference/skbio/tree/_tree.py:5337-5359): parent[k] > k for every non-root
 *     node, exactly one node has parent[k] == -1 (the last one);
 *   - results are SciPy "condensed" order: row-major upper triangle, element (i<j) at
 *     i*n + j - (i+2)(i+1)/2  (/root/reference/skbio/stats/distance/_base.py:2453), fp64;
 *   - one b2d_problem lives on one GPU and may own only a SHARD of the triangle: the
 *     128-sample row stripes are dealt to n_shards in snake order; a shard writes only its
 *     own stripes of the condensed vector (disjoint contiguous slices), so N processes or
 *     N threads - one per GPU - fill one result with no inter-GPU exchange;
 *   - thread-compatible: one call at a time per b2d_problem; different problems may be
 *     driven from different host threads.
 */
#ifndef SKBIO_B200_H
#define SKBIO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2D_API_VERSION 1

/* status codes */
#define B2D_OK 0
#define B2D_ERR_INVALID_ARGUMENT 1
#define B2D_ERR_CUDA 2
#define B2D_ERR_OUT_OF_MEMORY 3
#define B2D_ERR_NO_DEVICE 4
#define B2D_ERR_STATE 5

/* metrics */
#define B2D_UNWEIGHTED_UNIFRAC 0          /* replaces _unweighted_unifrac, beta/_unifrac.py:340-371 */
#define B2D_WEIGHTED_UNIFRAC 1            /* replaces _weighted_unifrac, beta/_unifrac.py:374-418 */
#define B2D_WEIGHTED_UNIFRAC_NORMALIZED 2 /* replaces _weighted_unifrac_normalized, :421-471 */
#define B2D_BRAYCURTIS 3                  /* replaces scipy pdist 'braycurtis' called at _driver.py:349-354 */
#define B2D_JACCARD 4                     /* replaces scipy pdist 'jaccard'    called at _driver.py:349-354 */
#define B2D_FAITH_PD 5                    /* per-sample vector only (no pair kernels): _faith_pd, alpha/_pd.py:34-51;
                                             read it with b2d_problem_sample_vector */

/* Mirrors skbb_get_api_version (/root/reference/skbio/binaries/_util.py:92-110). */
int b2d_api_version(void);
/* Number of CUDA devices visible; 0 when there is none (then every compute call fails). */
int b2d_device_count(void);
/* Message of the last failure on the calling thread ("" if none). */
const char* b2d_last_error(void);

/* Page-locked host memory for result vectors (so device->host copies are true DMA). */
void* b2d_host_alloc(size_t bytes);
void b2d_host_free(void* ptr);

typedef struct b2d_problem b2d_problem;

/*
 * Tree-based problem (UniFrac).  Replaces, on the device,
 *   vectorize_counts_and_tree  /root/reference/skbio/diversity/_util.py:123-170
 *   _nodes_by_counts           /root/reference/skbio/diversity/_phylogenetic.pyx:148-222
 *   _traverse_reduce           /root/reference/skbio/diversity/_phylogenetic.pyx:73-143
 *   _tip_distances             /root/reference/skbio/diversity/_phylogenetic.pyx:24-68
 *
 *   parent[n_nodes], length[n_nodes]  flat tree (length: NaN/None already replaced by 0.0)
 *   indptr[n_samples+1], node_ids[nnz], values[nnz]
 *       CSR sample x node table: the non-zero counts of each sample with the NODE id each
 *       observed taxon maps to (the host applies the reference's name lookup).  values are
 *       int64 (the reference truncates to int64, _phylogenetic.pyx:186); values may be NULL
 *       for presence/absence input (unweighted UniFrac only).  A node id may appear at most
 *       once per sample.
 *   shard, n_shards   this problem owns row stripes {b : snake(b, n_shards) == shard}.
 */
int b2d_problem_create_tree(b2d_problem** out, int device,
                            int64_t n_samples, int64_t n_nodes,
                            const int32_t* parent, const double* length,
                            const int64_t* indptr, const int32_t* node_ids,
                            const int64_t* values,
                            int shard, int n_shards);

/* The same with the CSR indices given as TABLE COLUMNS plus a column -> node id table
 * (node_of_col[n_cols], -1 for a column whose taxon is not in the tree; such a column must not be
 * observed): the composition column -> node -> engine position happens on the device, so the host
 * never rewrites the nnz-long index array (20 M entries at config 2). */
int b2d_problem_create_tree_columns(b2d_problem** out, int device,
                                    int64_t n_samples, int64_t n_nodes,
                                    const int32_t* parent, const double* length,
                                    const int64_t* indptr, const int32_t* col_ids,
                                    const int64_t* values, const int32_t* node_of_col,
                                    int64_t n_cols, int shard, int n_shards);

/* Host only: taxon name -> node id with the reference's lookup semantics (every node name scanned in
 * id order, a later node with the same name wins, /root/reference/skbio/diversity/_phylogenetic.pyx:195-199);
 * names and queries are byte blobs with offsets [count + 1] (every query entry is followed by
 * query_sep separator bytes that are not part of the name); out[k] = node id or -1.
 * b2d_names_validate: the name checks of _validate_taxa_and_tree (/root/reference/skbio/tree/_utils.py:83-90):
 * *dup_tips = a tip name occurs twice, *n_missing = distinct query names that are not tip names. */
int b2d_names_lookup(const char* name_buf, const int64_t* name_off, const uint8_t* has_name, int64_t n_names,
                     const char* query_buf, const int64_t* query_off, int64_t n_queries, int query_sep,
                     int32_t* out);
int b2d_names_validate(const char* name_buf, const int64_t* name_off, const uint8_t* has_name,
                       const int32_t* parent, int64_t n_names, const char* query_buf, const int64_t* query_off,
                       int64_t n_queries, int query_sep, int* dup_tips, int64_t* n_missing);

/*
 * Feature-table problem (Bray-Curtis, Jaccard): CSR sample x feature, values fp64 (SciPy
 * promotes the table to double, scipy/spatial/distance.py pdist); values may be NULL for
 * presence/absence input (Jaccard only).
 */
int b2d_problem_create_features(b2d_problem** out, int device,
                                int64_t n_samples, int64_t n_features,
                                const int64_t* indptr, const int32_t* feature_ids,
                                const double* values,
                                int shard, int n_shards);

/* Restrict the problem to an explicit list of 128x128 pair tiles, tiles[2*t] = row stripe,
 * tiles[2*t+1] = column stripe (row <= column).  The problem then owns exactly the row
 * stripes named; entries of those stripes outside the selected tiles read 0 after compute.
 * This is the engine's form of partial_beta_diversity / block_beta_diversity
 * (/root/reference/skbio/diversity/_driver.py:364-460, _block.py:263-356). */
int b2d_problem_select_tiles(b2d_problem* p, const int32_t* tiles, int64_t n_tiles);

/* Build the per-node sample embedding and run the pair kernels for `metric` on the
 * problem's device.  Blocks until the shard's distances are complete in device memory. */
int b2d_problem_compute(b2d_problem* p, int metric);

/* Distributed build of the per-block products (one process - or host thread - per GPU).
 *
 * Without this every shard builds the embedding, the sparse-row entries, the packed dense rows
 * and the group bitmaps of ALL samples, because its tiles need every column block: replicated
 * work that grows with the number of samples while the pair work per GPU stays fixed.  With
 * collectives installed, rank r of `world` builds those products only for its contiguous range
 * of 128-sample blocks and the ranks exchange them (everything is laid out block-major, so the
 * exchange is an in-place all-gather of slices of the same buffers); the pair kernels then run
 * exactly as before.  The library does not link a communication library: the host supplies two
 * callbacks (torch.distributed / NCCL over NVLink under torchrun - skbio_b200.distributed; peer
 * copies between host threads for beta_diversity(..., devices=[...])):
 *
 *   allgather_host(user, mine, all, bytes)   all[r * bytes .. ] = rank r's `mine` (host memory,
 *                                            a few dozen bytes: votes, entry counts, max U)
 *   allgatherv_device(user, n_buffers, bases, offsets, counts, world)
 *                                            device memory, in place, for each buffer k < n_buffers:
 *                                            on return every rank holds bytes [offsets[k*world + r],
 *                                            + counts[k*world + r]) of bases[k] as rank r wrote them
 *                                            (rank r's own slice is already there).  The library
 *                                            has drained its streams before the call; the callback
 *                                            returns when the data is usable by any stream of the
 *                                            device.
 * Both return 0 on success; every rank calls them in the same order.  Requires shard == rank and
 * n_shards == world at creation (otherwise B2D_ERR_INVALID_ARGUMENT).  Used by weighted and
 * unweighted UniFrac when the whole embedding is resident; a compute that cannot use it (several
 * node passes, custom tiles, a rank that voted no) is replicated as before.  Pairs flagged by the
 * fixed-point guard need the embedding of both samples: the rank that sees one redoes its own
 * compute replicated and votes no from then on.  world <= 1 or NULL callbacks remove them. */
typedef int (*b2d_allgather_host_fn)(void* user, const void* mine, void* all, int64_t bytes);
typedef int (*b2d_allgatherv_device_fn)(void* user, int n_buffers, void* const* bases, const int64_t* offsets,
                                        const int64_t* counts, int world);
int b2d_problem_set_collectives(b2d_problem* p, int rank, int world, b2d_allgather_host_fn allgather_host,
                                b2d_allgatherv_device_fn allgatherv_device, void* user);

/* Block range [first, last) of 128-sample blocks rank `rank` of `world` builds (host arithmetic only). */
int b2d_build_range(int64_t n_samples, int rank, int world, int64_t* first_block, int64_t* last_block);

/* Host only: minimum, number of non-zero and number of positive elements of a value array in one
 * multi-threaded pass (dtype: 0 int64, 1 float64, 2 int32, 3 float32; NaN never becomes the minimum).
 * What the host needs of the table's stored values for `_validate_counts_matrix`'s negativity check
 * (/root/reference/skbio/diversity/_util.py:51-54) and for the CSR conversion. */
int b2d_host_scan(const void* data, int64_t n, int dtype, double* min_out, int64_t* nonzero_out, int64_t* positive_out);

/* Device-to-device copy (same or peer device, unified addressing) that returns when the data is in place
 * (a plain device-to-device cudaMemcpy does not wait on the host's side): what an in-process
 * allgatherv_device callback is made of. */
int b2d_device_copy(void* dst, const void* src, int64_t bytes);

/* Number of condensed elements this shard owns. */
int64_t b2d_problem_result_len(const b2d_problem* p);

/* Copy the shard's slices into the FULL condensed vector out[n(n-1)/2] (device->host). Only
 * the owned slices are written.  `out` may be pageable or page-locked memory. */
int b2d_problem_fetch(b2d_problem* p, double* out_condensed);

/* Copy the shard's result COMPACTLY (owned stripes back to back, ascending stripe order)
 * into out[b2d_problem_result_len]; stripe_offsets (may be NULL) receives, for every owned
 * stripe, {first condensed index, count} pairs - at most 2*ceil(n/128) int64. */
int b2d_problem_fetch_compact(b2d_problem* p, double* out, int64_t* stripe_offsets,
                              int64_t* n_stripes_out);

/* Per-sample vector computed on the way (length n_samples): Faith's PD for a tree problem
 * after an unweighted compute (sum of branch lengths of observed nodes; replaces
 * _faith_pd, /root/reference/skbio/diversity/alpha/_pd.py:19-51). */
int b2d_problem_sample_vector(b2d_problem* p, double* out_n);

/* Generalised phylogenetic diversity per sample for a tree problem created WITH count values
 * (replaces _phydiv, /root/reference/skbio/diversity/alpha/_pd.py:190-239): rooted 0/1,
 * weight 0 (unweighted), 1 (abundance-weighted) or theta in (0,1).  out_n: n_samples doubles. */
int b2d_problem_phydiv(b2d_problem* p, int rooted, double weight, double* out_n);

/* Device-side phase timings of the last compute, milliseconds (CUDA events on the
 * problem's stream): [0] embedding build, [1] pair kernels, [2] total compute,
 * [3] last fetch (D2H), [4] upload at create (H2D), [5] last pair-kernel launch count,
 * [6] nodes resident per pass, [7] number of passes, [8] node visits the weighted pair
 * kernel executed (per 32x32 sample patch) and [9] the visits a kernel without zero-block
 * skipping would make (both 0 when skipping was not used), [10] the part of [1] spent on
 * the skipping bitmaps and closed-form sums, [11] kernels launched by the last compute (all
 * of them: embedding, bitmaps, pair launches, closed form, finalize), [12] duration of the
 * sparse-row intersection kernel (0 when every row went through the dense kernel), [13] number of
 * pairs the fixed-point guard flagged (near-duplicate samples; normally 0) - just those pairs were
 * recomputed in floating point from the resident embedding,
 * [14] largest number of tips below a node whose row went through the sparse intersection
 * kernel (0: none, 1: tip rows only), [15] entries of those rows, [16] (sample pair, row) matches
 * that kernel accumulated, [17] which intersection kernel ran (2: pipelined, b2d_isect.cuh; 1: first
 * generation; 0: none), [18] host time of the exchanges of a distributed build (ms, 0 when the
 * build was replicated) and [19] the bytes this rank received in them, [20] CSR bytes (indptr, indices,
 * values) uploaded from the host when the problem was created.  n: slots to fill (<= 21). */
int b2d_problem_timings(const b2d_problem* p, double* ms, int n);

int b2d_problem_destroy(b2d_problem* p);

/* One-shot convenience: create + compute + fetch + destroy (host buffers in, host buffer
 * out; this is the end-to-end call the Python host makes for a single GPU). */
int b2d_beta_diversity_tree(int metric, int device,
                            int64_t n_samples, int64_t n_nodes,
                            const int32_t* parent, const double* length,
                            const int64_t* indptr, const int32_t* node_ids,
                            const int64_t* values,
                            int shard, int n_shards, double* out_condensed);
int b2d_beta_diversity_features(int metric, int device,
                                int64_t n_samples, int64_t n_features,
                                const int64_t* indptr, const int32_t* feature_ids,
                                const double* values,
                                int shard, int n_shards, double* out_condensed);

/* Process-wide tuning knobs:
 *   "chunk_nodes"       nodes per pair-kernel launch = granularity of the two-level fp64
 *                       summation (0: default 8192; the zero-skipping kernel runs twice that)
 *   "emb_budget_bytes"  cap on the embedding bytes resident per pass (0: 85 % of free memory)
 *   "weighted_skip"     1 (default): the weighted / dense Bray-Curtis pair kernel skips nodes at
 *                       which a 32-sample group is all zero and adds their closed-form
 *                       contribution; 0: every node is visited
 *   "tips_sparse"       1 (default): with weighted_skip, the tip rows of a tree problem go through a
 *                       sparse intersection kernel in exact fixed point instead of the dense
 *                       kernel; 0: dense.  Falls back by itself for very dense tables and when
 *                       near-duplicate samples are found.
 *   "sparse_rows_max_tips"  0 (default): when the whole embedding is resident, rows of nodes with
 *                       up to about 0.3 / (table density) tips below them (0.16 / density for
 *                       unweighted UniFrac) join the tips in the intersection kernel; k > 0: up to k tips (1 = tip rows only)
 *   "cache_embedding"   1 (default): device buffers are parked per device when a problem is
 *                       destroyed and reused by the next one (cudaMalloc / cudaFree of large blocks
 *                       cost 0.1 - 0.7 s), up to "cache_limit_bytes" per device; they are released
 *                       when an allocation fails; 0: release now, stop parking
 *   "cache_limit_bytes" most bytes parked per device (default 4 GiB; what does not fit goes back
 *                       to the driver when it is freed); "release_cache": release everything now
 *   "sparse_features"   -1 auto, 0 never, 1 whenever legal: intersection kernel for integer
 *                       feature tables (Bray-Curtis / Jaccard)
 *   "walk_decomposed"   0: always the sequential tree walk;  "plain_kernels" 1: cross-check
 *                       kernels without bulk copies.
 *   "lazy_proportions"  1 (default): with the whole embedding resident and sparse rows in use, the kernels that read
 *                       the embedding divide count / total themselves (same correctly rounded division, identical
 *                       bits) instead of a separate pass over the whole embedding; 0: the separate pass
 *   "partial_upload"    1: tree problems created with n_shards > 1 upload only the CSR rows of the shard's own
 *                       build range (b2d_build_range); the first compute (or phydiv) all-gathers the other
 *                       shards' rows device to device, so collectives must be installed before it.  Set it
 *                       on every rank before creating the problems.  Default 0: every shard uploads all rows
 *   "isect_kernel"      -1 (default): pipelined intersection kernel for the sparse rows (b2d_isect.cuh),
 *                       except unweighted UniFrac with more than 6 000 keys per (block, chunk) bucket;
 *                       1: always; 0: the first-generation kernels
 *   "isect_group"       lanes sharing one row entry in that kernel: 1, 2, 4 or 8; 0 (default): 4 for
 *                       weighted UniFrac, 2 for unweighted, 1 for feature tables */
int b2d_set_option(const char* name, int64_t value);

/* On-device micro-benchmarks used by bench.py for the roofline denominators:
 * kind 0: FP64 pipe, returns DFMA instructions/s;  kind 1: HBM copy, returns bytes/s;
 * kind 2: conflict-free 64-bit shared-memory gathers/s (the unweighted kernel's bound);
 * kind 3: shared-memory int32 atomicAdd/s on random cells (the sparse feature kernel's bound). */
int b2d_microbench(int device, int kind, double* result);

/* PERMANOVA partial statistic for many groupings at once ("next" consumer of the distance
 * matrix): s_w_out[g] = sum over pairs i<j with groupings[g][i] == groupings[g][j] of
 * d_ij^2 * inv_group_size[group].  Replaces the per-permutation loop over
 * permanova_f_stat_sW_condensed_cy (/root/reference/skbio/stats/distance/_cutils.pyx:418-470,
 * /root/reference/skbio/stats/distance/_base.py:2340-2364); the optional accelerator the reference
 * binds for the same job is skbb_permanova_fp64 (/root/reference/skbio/binaries/_distance.py:161-193).
 * condensed: n(n-1)/2 distances (host); groupings: [n_groupings][n] int32 group ids. */
int b2d_permanova_sw(int device, int64_t n, const double* condensed, int64_t n_groupings,
                     const int32_t* groupings, int32_t n_groups, const double* inv_group_size,
                     double* s_w_out);

/* Rows [r0, r1) of the square (redundant) distance matrix, out[(i - r0) * n + j], gathered on the
 * device from the condensed result of a problem that owns the whole triangle (n_shards == 1, no tile
 * selection).  This is what a streaming lsmat / binary_dm writer consumes row block by row block
 * (/root/reference/skbio/io/format/lsmat.py:266-279, binary_dm.py:157-184) without ever holding the
 * n x n matrix. */
int b2d_problem_fetch_square_rows(b2d_problem* p, int64_t r0, int64_t r1, double* out);

/* Gower centering of the distance matrix, the first step of pcoa: E = -d^2/2, its row means and
 * global mean, F = E - mean_row - mean_col + mean (replaces e_matrix_means_cy + f_matrix_inplace_cy,
 * /root/reference/skbio/stats/ordination/_cutils.pyx:19-123).  row_means[n], *global_mean and
 * f_matrix[n*n] are host outputs; each may be NULL.  b2d_problem_center works on the device-resident
 * result of a problem that owns the whole triangle (the 8*P bytes never travel to the host);
 * b2d_center_condensed takes a condensed host vector. */
int b2d_problem_center(b2d_problem* p, double* row_means, double* global_mean, double* f_matrix);
int b2d_center_condensed(int device, int64_t n, const double* condensed, double* row_means,
                         double* global_mean, double* f_matrix);

/* Host only: Newick text -> flat tree in the reference's node id order, without node objects
 * (stands in for TreeNode.read + TreeNode.to_array, /root/reference/skbio/io/format/newick.py:270-336,
 * /root/reference/skbio/tree/_tree.py:6235-6329).  length[k] is NaN where the text gave none;
 * names are returned as one byte buffer with offsets (UTF-8 bytes as in the text). */
typedef struct b2d_flat_tree b2d_flat_tree;
int b2d_newick_parse(const char* text, int64_t len, b2d_flat_tree** out);
int64_t b2d_flat_tree_nodes(const b2d_flat_tree* t);
int64_t b2d_flat_tree_name_bytes(const b2d_flat_tree* t);
int b2d_flat_tree_get(const b2d_flat_tree* t, int32_t* parent, double* length, int64_t* name_off,
                      uint8_t* has_name, char* name_buf);
void b2d_flat_tree_free(b2d_flat_tree* t);

/* Host only (no GPU needed): which slices of the condensed vector shard `shard` of
 * `n_shards` owns: {first condensed index, count} per owned row stripe, ascending (at most
 * 2*ceil(n/128) int64), the number of stripes and the total element count.  This is the
 * layout b2d_problem_fetch_compact returns. */
int b2d_shard_layout(int64_t n_samples, int shard, int n_shards, int64_t* stripe_offsets,
                     int64_t* n_stripes_out, int64_t* total_out);

/* Test hook (needs the GPU): the per-node sample embedding the device builds - CSR scatter + tree
 * propagation, the engine's replacement of _nodes_by_counts / _traverse_reduce
 * (/root/reference/skbio/diversity/_phylogenetic.pyx:73-222) - copied back as
 * out[n_samples][n_nodes] int64 in REFERENCE node id order, i.e. the matrix
 * vectorize_counts_and_tree returns (/root/reference/skbio/diversity/_util.py:123-170; pinned by
 * /root/reference/skbio/diversity/tests/test_util.py:27-34).  variant 0: subtree counts, sequential
 * walk in node passes (honours "emb_budget_bytes"); 1: subtree counts, subtree-decomposed walk;
 * 2: presence bits as 0 / 1. */
int b2d_debug_fetch_embedding(b2d_problem* p, int variant, int64_t* out);

/* Test hook (host only, no GPU needed): the engine's node order for a flat tree.  Arrays
 * sized n_nodes (q_of_id) or n_nodes rounded up to a multiple of 128 (the others); any may
 * be NULL.  flags: bit0 = node has children, bit1 = fold into the sibling accumulator,
 * bit2 = padding. */
int b2d_debug_tree_order(int64_t n_nodes, const int32_t* parent, const double* length,
                         int32_t* q_of_id, uint8_t* flags_mpad, double* len_q_mpad,
                         double* tipdist_q_mpad, int32_t* max_depth);

/* Test hook (host only): the per-position arrays behind the closed-form part of zero-block
 * skipping, all sized n_nodes rounded up to 128: parent position (-1 root / padding), subtree
 * size (the subtree of position q is the range [q - size + 1, q]) and the distance to the
 * root (sum of lengths over q .. root) as an unevaluated double-double sum hi + lo; the number
 * of tips in the subtree (0 in the padding); and the root distance without the tips' own
 * branches (a tip carries its parent's).  Any output may be NULL. */
int b2d_debug_tree_paths(int64_t n_nodes, const int32_t* parent, const double* length,
                         int32_t* parent_q_mpad, int32_t* size_q_mpad, double* rd_hi_mpad,
                         double* rd_lo_mpad, int32_t* tips_q_mpad, double* rd_hi_int_mpad,
                         double* rd_lo_int_mpad);

#ifdef __cplusplus
}
#endif
#endif /* SKBIO_B200_H */
