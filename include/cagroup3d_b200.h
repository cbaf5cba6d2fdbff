/* cagroup3d_b200 -- C ABI of the B200-native CAGroup3D inference hot path.
 *
 * One shared library (cagroup3d_b200/libcagroup3d_b200.so), plain pointers and sizes, no torch
 * types.  Every pointer is a DEVICE pointer unless its name starts with `h_`; `stream` is a
 * cudaStream_t passed as void*.  Every function returns 0 on success, a cudaError_t (> 0) when a
 * launch failed, or a negative value for an argument it refuses.  Nothing allocates, nothing
 * synchronises, nothing calls exit(): the caller owns all buffers (sizes stated per function).
 *
 * The reference interface each entry point replaces is cited as file:line relative to
 * /root/reference (Haiyang-W/CAGroup3D).  "ME" = MinkowskiEngine v0.5.4, the un-vendored
 * dependency whose coordinate manager and convolution the reference calls from Python; SURVEY.md
 * Appendix A (A1..A20) is the statement of its semantics these functions implement.
 *
 * Coordinates are int32 rows (b, x, y, z), 16 bytes, |x|,|y|,|z| < 32768, 0 <= b < 65536.
 * Feature matrices are row-major fp32 [rows, channels].
 * A hash table is (keys: u64[capacity], vals: i32[capacity]), capacity = cg3d_hash_capacity(n).
 * A rule map is a neighbour table nbr: i32[K][n_out] (tap-major), entry = input row or -1; its
 * rule pairs {(tap, in, out)} are exactly ME's kernel map.
 */
#ifndef CAGROUP3D_B200_H
#define CAGROUP3D_B200_H

#ifdef __cplusplus
extern "C" {
#endif

/* ---- coordinate manager (ME coordinate maps: cagroup3d.py:24, biresnet.py strided convs,
 *      cagroup_head.py:257-271, cagroup_roi_head.py:62-69) -------------------------------------- */

/* smallest power of two >= 2n (>= 1024). host-only helper. */
int cg3d_hash_capacity(int n);

/* ints of scratch cg3d_exclusive_scan_i32 needs in `block_sums`. host-only helper. */
int cg3d_scan_workspace_ints(int n);

/* out[i] = (b, floor(x/vx)*mul, floor(y/vy)*mul, floor(z/vz)*mul) from rows pts[i*ld + 0..3] = (b,x,y,z).
 * IEEE fp32 division then floor, i.e. ME's float->int coordinate conversion (A1) after the
 * reference's `coordinates[:, 1:] /= voxel_size` (cagroup3d.py:21-22).  *err_count += rows whose
 * voxel index does not fit the 16-bit key fields. */
int cg3d_quantize(const float* pts, int ld, int n, float vx, float vy, float vz, int mul, int* out_coords,
                  int* err_count, void* stream);

/* out = (b, floor(c/ts)*ts) : coordinates of the stride-`ts` map (A6). */
int cg3d_stride_coords(const int* coords, int n, int ts, int* out, void* stream);

/* exclusive prefix sum; *total = sum.  block_sums: cg3d_scan_workspace_ints(n) ints. */
int cg3d_exclusive_scan_i32(const int* in, int n, int* out, int* block_sums, int* total, void* stream);

/* Hash-unique with the first-occurrence (min row index) winner == ME CPU insert order (A2).
 *   out_coords[u]  unique rows in order of first occurrence          (capacity n rows)
 *   first_row[u]   input row that won                (may be NULL)
 *   inverse[i]     unique row of input row i         (may be NULL)
 *   *n_unique      number of unique rows (device int)
 * On return the table maps key -> unique row.  workspace: 3*n + cg3d_scan_workspace_ints(n) ints. */
int cg3d_unique_first(const int* coords, int n, unsigned long long* keys, int* vals, int capacity, int* out_coords,
                      int* first_row, int* inverse, int* n_unique, int* workspace, void* stream);

/* The same for a source whose row count still lives in DEVICE memory: rows 0 .. min(*n_rows, n_max) - 1 of `coords`
 * are strided to tensor stride `ts` (ts <= 1: taken as they are) and hash-uniqued in first-occurrence order.  Launches are
 * sized by the upper bound n_max, so a whole pyramid of strided maps (the stride 2 .. 512 maps of BiResNet and DAPPM,
 * biresnet.py:109-127,265-268: each level made from the unique rows of another) is built without reading a size back; the
 * host reads all counts once at the end.  capacity: power of two >= 2 * n_max; out_coords: n_max rows;
 * workspace: 3 * n_max + cg3d_scan_workspace_ints(n_max) ints. */
int cg3d_unique_first_dev(const int* coords, const int* n_rows, int n_max, int ts, unsigned long long* keys, int* vals,
                          int capacity, int* out_coords, int* n_unique, int* workspace, void* stream);

/* table over already-unique rows: key -> row index. */
int cg3d_hash_build(const int* coords, int n, unsigned long long* keys, int* vals, int capacity, void* stream);

/* rows[i] = row of query[i] in the table or -1. */
int cg3d_hash_lookup(const int* query, int n, const unsigned long long* keys, const int* vals, int capacity, int* rows,
                     void* stream);

/* ME kernel map of a (possibly strided) convolution or of conv(x, coordinates) (A5-A7, A12):
 * nbr[tap][o] = row of (out_coords[o] + offset(tap) * step) in the input table.  Taps are x-fastest,
 * centred for odd ksize and 0..k-1 for even ksize; step = the input's tensor stride. */
int cg3d_neighbor_table(const int* out_coords, int n_out, const unsigned long long* keys, const int* vals,
                        int capacity, int ksize, int step, int* nbr, void* stream);
/* The same table when the output rows ARE the input map's rows (same-stride convolution, odd ksize): the rule map is
 * symmetric (nbr[t][o] = i <=> nbr[K-1-t][i] = o), so only the taps below the centre are probed and every hit also
 * writes its mirror entry.  Result identical to cg3d_neighbor_table. */
int cg3d_neighbor_table_symmetric(const int* coords, int n, const unsigned long long* keys, const int* vals, int capacity,
                                  int ksize, int step, int* nbr, void* stream);

/* transposed-convolution kernel maps onto existing fine coordinates:
 *   ksize 2: MinkowskiConvolutionTranspose(k=2,s=2) (biresnet.py:309; A8), ts_coarse = input stride
 *   ksize 3: MinkowskiGenerativeConvolutionTranspose(k=3,s=3)(E, coordinates=A.C) (cagroup_head.py:274; A13)
 * nbr: i32[ksize^3][n_fine], one valid tap per row. */
int cg3d_transpose_table(const int* fine_coords, int n_fine, const unsigned long long* keys, const int* vals,
                         int capacity, int ksize, int ts_coarse, int* nbr, void* stream);

/* Tile ordering of a coordinate map.  keys[i] = b << 33 | Morton code of (x, y, z) / stride (11 bits per axis,
 * wrapped), vals[i] = i.  Sorting the pairs on key bits [6, 33 + batch bits) with cg3d_sort_pairs gives a row
 * permutation `order` in which every run of 128 rows is a compact patch of one sample, so a conv tile touches few
 * distinct taps and neighbouring input rows.  ME leaves the row order of a map unspecified; here the ORDER OF THE
 * MAP IS NOT CHANGED -- the permutation only decides which output rows share a CTA (rule maps are built for
 * cg3d_gather_coords(coords, order) and the conv writes row order[j]). */
int cg3d_morton_keys(const int* coords, int n, int stride, unsigned long long* keys, int* vals, void* stream);
int cg3d_gather_coords(const int* coords, const int* order, int n, int* out, void* stream);

/* Tile ordering by tap pattern.  keys[o] = bit mask of the taps of rule map `nbr` (i32[K][n], K = ksize^3) that
 * have a neighbour for output row o (K <= 27), or of the 27 coarse blocks of taps (K > 27); vals[o] = o.  Sorting
 * the pairs (cg3d_sort_pairs on bits [0, 27)) groups rows with the same pattern; cg3d_permute_table then gives the
 * positional table out[k][j] = nbr[k][order[j]] whose 128-row tiles touch few taps -- the convolution skips the
 * (tile, tap) pairs that are empty.  The map's row order is not changed (see out_rows of cg3d_spconv_*).
 * coords != NULL: bits 27.. of the key = batch index of the row / group_div, so that sorted rows stay grouped
 * (grouped convolutions: one weight group per class, batch index = class * B + b). */
int cg3d_table_mask_keys(const int* nbr, int K, int ksize, int n, const int* coords, int group_div,
                         unsigned long long* keys, int* vals, void* stream);
int cg3d_permute_table(const int* nbr, int K, int n, const int* order, int* out, void* stream);

/* *count = number of entries >= 0 (rule pairs P of SURVEY.md section 8d). */
int cg3d_count_rules(const int* nbr, long long total, unsigned long long* count, void* stream);

/* ---- sparse convolution (ME MinkowskiConvolution / ConvolutionTranspose forward, A4-A8, A19) ---- */

/* out[o, :] = act( (sum_k in_act(in[nbr[k][o], :]) @ W[g][k]) * scale[g] + shift[g] + residual[o, :] )
 * W: [G][K][Cin][Cout] (ME's `kernel` layout).  nbr == NULL means K == 1 on identical rows (1x1 conv,
 * nn.Linear).  scale/shift/residual may be NULL (folded eval-mode BatchNorm / bias / block residual).
 * ldi / ldo: row strides (floats) of `in` / `out` (a conv may read a column slice of a wider matrix and
 * write into a slice of a concat buffer); residual rows are dense [n_out, Cout].
 * in_act: activation applied to gathered input rows (1 = the `self.relu(x)` in front of every BiResNet
 * stage, biresnet.py:366-394); act: activation of the result.  0 none, 1 ReLU, 2 ELU.
 * Grouped mode (tile_row0 != NULL): tile t covers rows [tile_row0[t], +tile_rows[t]) (<= 64 for simt,
 * <= 128 for tc) with weight/scale/shift group tile_group[t] -- used to run all per-class convolutions
 * of cagroup_head.py:227-282 in one launch.
 * out_rows (may be NULL): the table is POSITIONAL -- nbr[k][j] belongs to output row out_rows[j] (tile order from
 * cg3d_morton_keys); the result / residual of position j live at row out_rows[j].  NULL: position == row.
 * cg3d_spconv_simt: exact fp32 FFMA path (any Cin/Cout). */
int cg3d_spconv_simt(const float* in, int ldi, int in_act, const int* nbr, const float* W, float* out, int ldo,
                     int n_out, int Cin, int Cout, int K, const float* scale, const float* shift,
                     const float* residual, int act, const int* tile_row0, const int* tile_rows,
                     const int* tile_group, int n_tiles, const int* out_rows, void* stream);

/* Tensor-core path of the same contraction (tcgen05, accumulator in TMEM, 128 rows x NT columns per CTA, two CTAs
 * per SM).  fp32 in / fp32 out; inside, both operands are split into bf16 hi + lo and three products are
 * accumulated in fp32 (error <= ~2^-16 relative per product).  Needs Cin % 32 == 0, Cout % 64 == 0, K <= 729.
 * Both operands are split ONCE, outside the kernel:
 *   cg3d_split_bf16         activations: out[r] = per 32-channel chunk [bf16 hi (32) | bf16 lo (32)] of in[r] (or of
 *                           relu(in[r]) when relu = 1: the activation the consumer applies to its input); same byte
 *                           count as the fp32 rows.  The conv gathers these rows straight into shared memory with
 *                           cp.async (no register staging, no per-tap conversion);
 *   cg3d_spconv_tc_prepare  weights: the pre-swizzled shared-memory image of W, made once per weight tensor (same
 *                           byte count as the fp32 weights), streamed by bulk async copies.
 * cg3d_spconv_tc_ntile(Cout) = NT (0: unsupported).  Grouped mode / out_rows as above with tiles of <= 128 rows.
 * n_in: rows of in_split (rule-map entries must lie in [0, n_in) or be -1 = no neighbour: a zero row).  Launches whose
 * column tile is 64 wide (Cout == 64 layers) keep the gathered operand in tensor memory (spconv_ts.cu: register gather ->
 * tcgen05.st, tcgen05.mma with A from TMEM); wider tiles gather into shared memory with cp.async.
 * splitk_ws (may be NULL): launches with few row tiles and a long K loop (the 7^3 RoI pooling contraction) are split over
 * cg3d_spconv_tc_splitk(...) CTAs per tile; the partial slabs (that many x n_out x Cout floats) are added in a fixed order
 * by a second pass, so the result stays deterministic.
 * out_split (may be NULL): the epilogue also writes the result (ReLU'd when out_split_relu = 1) in the split layout
 * [n_out][2 * Cout], i.e. the operand of the next convolution, which then needs no cg3d_split_bf16 pass. */
int cg3d_spconv_tc_ntile(int Cout);
int cg3d_spconv_tc_prepare(const float* W, int G, int K, int Cin, int Cout, unsigned char* img, void* stream);
int cg3d_split_bf16(const float* in, int ld, int n, int C, int relu, unsigned short* out, void* stream);
int cg3d_spconv_tc_splitk(int n_out, int Cin, int Cout, int K, int grouped, int n_tiles);   /* split-K factor (1 = none) */
int cg3d_spconv_tc(const unsigned short* in_split, int n_in, const int* nbr, const unsigned char* wimg, float* out, int ldo,
                   int n_out, int Cin, int Cout, int K, const float* scale, const float* shift, const float* residual,
                   int act, const int* tile_row0, const int* tile_rows, const int* tile_group, int n_tiles,
                   const int* out_rows, unsigned short* out_split, int out_split_relu, float* splitk_ws, void* stream);

/* out = act(x * scale + shift + add) on an [n, C] matrix with row strides ldx / ldo; scale, shift, add
 * ([n, C] dense) may be NULL.  Pre-activation BatchNorm+ReLU of DAPPM (biresnet.py:109-174). */
int cg3d_affine_act(const float* x, int ldx, const float* scale, const float* shift, const float* add, float* out,
                    int ldo, long long n, int C, int act, void* stream);

/* ---- pooling / interpolation / quantisation ---------------------------------------------------- */

/* SparseTensor.features_at_coordinates (biresnet.py:182-197,376,389,394; A9):
 * out[q] = base[q] + trilinear sample of the stride-`ts` tensor (feats, table) at integer query rows. */
int cg3d_interp_trilinear(const int* query, int nq, const unsigned long long* keys, const int* vals, int capacity,
                          int ts, const float* feats, int C, const float* base, float* out, void* stream);

/* MinkowskiAvgPooling (biresnet.py:109-127; A10): out[o] = mean of input rows of the same batch
 * index with |dx|,|dy|,|dz| <= half. */
int cg3d_avgpool_window(const int* out_coords, int n_out, const int* in_coords, int n_in, int half, const float* feats,
                        int C, float* out, void* stream);

/* UNWEIGHTED_AVERAGE quantisation (cagroup_head.py:257-271; A3): out[u] = mean over points p with
 * inverse[p] == u of feat(p); ref == NULL: feat(p) = srcA[p*ldA ..]; else ref[p] = (row, kind):
 * kind >= 0 -> srcA[row*ldA + kind*C ..], kind < 0 -> srcB[row*ldB ..].  counts: n_unique floats; workspace:
 * cg3d_segment_mean_workspace(n, n_unique) 64-bit words.  Histogram -> scan -> segment fill -> one warp per unique row;
 * the sums are taken in 2^-30 fixed point (integer adds), so the result does not depend on the order in which a
 * segment was filled: bit-repeatable forward, no 64-bit atomics. */
int cg3d_segment_mean_workspace(int n, int n_unique);
int cg3d_segment_mean(const float* srcA, int ldA, const float* srcB, int ldB, const int* ref, const int* inverse,
                      int n, int n_unique, int C, float* out, float* counts, long long* workspace, void* stream);

/* out[r, :] = src[rows[r]*ld + col0 .. +C] / divisor  (rows == NULL: identity). */
int cg3d_gather_rows(const float* src, int ld, int col0, const int* rows, int n, int C, float divisor, float* out,
                     void* stream);

/* ---- detection head (cagroup_head.py) ---------------------------------------------------------- */

/* minmax6 = (min x,y,z, max x,y,z) over all rows (cagroup_head.py:209-211). */
int cg3d_coord_bounds(const int* coords, int n, int* minmax6, void* stream);

/* first[b] = first row of sample b (pad_id, cagroup_head.py:207); 0x7F7F7F7F where a sample is empty. */
int cg3d_first_rows(const int* coords, int n, int B, int* first, void* stream);

/* voted[i, v, :] = clamp(coords[i]*voxel_size + offsets[i, v, :], bounds -+ tensor_stride) (:216-225). */
int cg3d_vote_points(const int* coords, const float* offsets, int n, int nv, float voxel_size, int tensor_stride,
                     const int* minmax6, float* voted, void* stream);

/* flags[c*n + i] = sigmoid(sem[i, c]) > thr (:229-230), class-major. */
int cg3d_semantic_flags(const float* sem, int n, int ncls, float thr, int* flags, void* stream);

/* sel_rows[pos[t]] = t % n for every set flag (pos = exclusive scan of flags). */
int cg3d_compact_rows(const int* flags, const int* pos, int n, int ncls, int* sel_rows, void* stream);

/* All classes at once (:231-271): fused point p of class c -> coordsA (class voxel size),
 * coordsE (expand * class voxel size, multiplied back by expand), ref = (source row, vote index or -1).
 * Batch index of the emitted rows is c*B + b. */
int cg3d_class_points(const int* coords, const float* voted, const int* sel_rows, const int* sel_off,
                      const int* fused_off, const int* pad_rows, const float* vsA, const float* vsE, int ncls, int B,
                      int nv, int expand, int total, float voxel_size, int* coordsA, int* coordsE, int* ref,
                      void* stream);

/* pred rows = [centerness | cls logits (ncls) | reg (nreg)]; scores = sigmoid(cls)*sigmoid(ctr),
 * boxes per _bbox_pred_to_bbox (:590-593, 636-649, 654-703). */
int cg3d_head_decode(const float* pred, int ld, const int* coords, int n, int ncls, int nreg, int B, const float* vsA,
                     const float* scales, float* scores, float* maxscore, float* boxes, int box_dim, void* stream);

/* ---- NMS / IoU (pcdet/ops/iou3d_nms) ----------------------------------------------------------- */

/* mode 0: boxes_overlap_bev_gpu (iou3d_nms.cpp:44-66), 1: boxes_iou_bev_gpu (:68-88),
 * 2: axis-aligned BEV IoU (iou3d_nms_kernel.cu:314-325).  out: [na, nb]. */
int cg3d_boxes_pairwise_bev(const float* boxes_a, int na, const float* boxes_b, int nb, int mode, float* out,
                            void* stream);

/* nms_gpu (rotated=1, iou3d_nms.cpp:90-136) / nms_normal_gpu (rotated=0, :139-186) for n_segments
 * independent instances in one launch.  Instance s owns rows [seg_offsets[s], seg_offsets[s+1]) of
 * sorted_boxes [n_boxes, 7], already in descending-score order; max_segment_len >= the longest instance.
 * keep[i] in {0,1}; kept_count[s] (may be NULL).
 * The reference returns keep indices through a CPU tensor after a blocking D2H of an N x N/64 bitmask and a host
 * loop (iou3d_nms.cpp:110-133); here the flags stay on the device and all (sample, class) instances run in one
 * launch: the alive bitset in shared memory, instances walked in blocks of 64 boxes (pair tile of the block, keep
 * bits of the block, then every later box against the block's kept boxes), a thread-block cluster of up to 8 CTAs
 * per instance when the launch has fewer instances than SMs.  -2: max_segment_len too large for the bitset. */
int cg3d_nms_segments(const float* sorted_boxes, int n_boxes, const int* seg_offsets, int n_segments, int max_segment_len,
                      float thr, int rotated, int* keep, int* kept_count, void* stream);

/* ---- selection primitives (torch.sort / topk / nonzero / boolean masks on the path) ---- */

/* ints of scratch cg3d_sort_pairs needs. host-only helper. */
int cg3d_sort_workspace_ints(int n);

/* stable LSD radix sort of (u64 key, i32 value) pairs on key bits [begin_bit, end_bit) (rounded up to whole 8-bit
 * digits), ascending, in place (keys_tmp / vals_tmp: n elements of scratch each; workspace: cg3d_sort_workspace_ints(n)
 * ints, cleared by the call).  One launch builds the histograms of all digits, then one launch per digit: tiles of 2048 or 4096
 * keys in ticket order, the offsets of a tile from the published counts of the tiles before it (chained look-back).  Replaces scores.sort(descending=True)
 * (iou3d_nms_utils.py:92,110) and max_scores.topk (cagroup_head.py:596) with keys built by the
 * functions below: (segment << 32 | ~ordered(score)), ties keep the lower index first. */
int cg3d_sort_pairs(unsigned long long* keys, int* vals, int n, int begin_bit, int end_bit,
                    unsigned long long* keys_tmp, int* vals_tmp, int* workspace, void* stream);

/* counts[s] = #{i : ids[i] == s}, s in [0, nseg). */
int cg3d_histogram_i32(const int* ids, int n, int nseg, int* counts, void* stream);

/* out[pos[i]] = payload ? payload[i] : i for every set flag (torch.nonzero / boolean-mask gather). */
int cg3d_compact_i32(const int* flags, const int* pos, int n, const int* payload, int* out, void* stream);

/* seg[i] = b * ncls + c for class-map voxel rows whose batch index is c * B + b. */
int cg3d_map_segments(const int* coords, int n, int B, int ncls, int* seg, void* stream);

/* top NMS_PRE per (sample, class map) (cagroup_head.py:595-599): key = (seg, ~maxscore) for segments
 * longer than nms_pre, (seg, 0) otherwise (keeps row order); vals = row. */
int cg3d_topk_keys(const int* seg, const float* maxscore, int n, const int* seg_counts, int nms_pre,
                   unsigned long long* keys, int* vals, void* stream);

/* flags[i] = rank of sorted element i inside its segment < nms_pre. */
int cg3d_rank_filter(const unsigned long long* keys, int n, const int* seg_off, int nms_pre, int* flags, void* stream);

/* flags[j * ncls + i] = scores[cand[j], i] > thr (cagroup_head.py:753). */
int cg3d_pair_flags(const float* scores, const int* cand, int nc, int ncls, float thr, int* flags, void* stream);

/* compacted (key, source row) of every flagged pair; key = ((b * ncls + i) << 32 | ~score). */
int cg3d_pair_keys(const float* scores, const int* cand, const int* seg_of_row, int nc, int ncls, const int* flags,
                   const int* pos, unsigned long long* keys, int* src_row, void* stream);

/* RoI-stage equivalents (cagroup_roi_head.py:437-446): one pair per RoI, class = roi_labels. */
int cg3d_roi_flags(const float* roi_scores, int n, float thr, int* flags, void* stream);
int cg3d_roi_keys(const float* roi_scores, const int* roi_labels, int n, int rois_per_sample, int ncls,
                  const int* flags, const int* pos, unsigned long long* keys, int* src_row, void* stream);

/* seg[i] = keys[i] >> 32. */
int cg3d_key_segments(const unsigned long long* keys, int n, int* seg, void* stream);

/* out[p] = 7-wide box of row src_row[p] (yaw 0 if box_dim == 6; negated if flip) = boxes[order]. */
int cg3d_gather_boxes(const float* boxes, int box_dim, const int* src_row, int n, int flip, float* out, void* stream);

/* pack survivors in (sample, class, descending score) order: boxes, scores, labels, sample index
 * (cagroup_head.py:773-797, cagroup_roi_head.py:455-475). */
int cg3d_emit_detections(const float* sorted_boxes, const unsigned long long* keys, const int* keep, const int* pos,
                         int n, int ncls, int with_yaw, int flip, float* out_boxes, float* out_scores, int* out_labels,
                         int* out_sample, void* stream);

/* reoder_rois_for_refining (cagroup_roi_head.py:328-362): zero-padded (B, rmax, 7) RoIs, yaw negated. */
int cg3d_pad_rois(const float* det_boxes, const float* det_scores, const int* det_labels, const int* sample_off, int B,
                  int rmax, float* rois, float* roi_scores, int* roi_labels, void* stream);

/* ---- RoI head (cagroup_roi_head.py) ------------------------------------------------------------ */

/* 7^3 grid points per RoI -> voxel rows at coord_key * clamp(floor(p / voxel_size)) (:54-68, :199-224). */
int cg3d_roi_grid_coords(const float* rois, int n_rois, int rois_per_sample, int grid, int with_yaw, float voxel_size,
                         int half_extent, int coord_key, int* out_coords, void* stream);

/* rule map of the 7^3 "pooling conv" evaluated at RoI centres (:72-90; A20) on top of the
 * unique-inverse of the grid voxels: nbr: i32[grid^3][n_rois]. */
int cg3d_roi_pool_table(const int* inverse, int n_rois, int grid, int* nbr, void* stream);

/* CAGroupResidualCoder.decode_torch + rotate + add centre (:477-510, cagroup_utils.py:147-197). */
int cg3d_roi_decode(const float* rois, const float* reg, int n, int code_size, int sincos, float* out, void* stream);

/* ---- training-loss ops (pcdet/ops/knn, pcdet/ops/rotated_iou/cuda_op) ------------------------------ */

/* knn_wrapper(b, n, m, nsample, xyz, new_xyz, idx, dist2) (knn.cpp:28-46, knn_cuda.cu:58-94): for each of
 * the m queries of every batch element the k <= 100 nearest of its n points, ascending squared distance.
 * xyz [b,n,3], query [b,m,3] fp32; idx [b,m,k] i32, dist2 [b,m,k] f32.  Candidates are scanned in index
 * order with strict `<`, so ties resolve as in the reference (first index wins at k = 1). */
int cg3d_knn(const float* xyz, int b, int n, const float* query, int m, int k, int* idx, float* dist2, void* stream);

/* The k = 1 case on a uniform grid over the points (cells of >= 0.04 m, counting sort, ring search): idx / dist2
 * ([b,m]) bit-identical to cg3d_knn / knn_cuda.cu -- the winner is the minimum of (dist2, index) with the same distance
 * expression -- at O(m * points near a query) instead of O(m n).  workspace: cg3d_knn_grid_workspace(n) ints, 16-byte
 * aligned.  Fewer than 4096 points per batch element: the exhaustive kernel. */
int cg3d_knn_grid_workspace(int n);
int cg3d_knn_grid(const float* xyz, int b, int n, const float* query, int m, int* idx, float* dist2, int* workspace, void* stream);

/* sort_vertices_forward(vertices [b,n,m,2] f32, mask [b,n,m] bool, num_valid [b,n] i32) -> idx [b,n,9] i32
 * (sort_vert.cpp:6-33, sort_vert_kernel.cu:42-134): counter-clockwise order of the valid polygon vertices,
 * first index repeated, padded with an invalid intersection index.  The caller allocates idx (the reference
 * allocates it itself); 9 <= m <= 32 (the reference passes m = 24). */
int cg3d_sort_vertices(const float* vertices, const unsigned char* mask, const int* num_valid, int b, int n, int m,
                       int* idx, void* stream);

/* ---- backward of the sparse convolution (training; MinkowskiEngine's ConvolutionBackward over the forward's kernel
 * map; called from cagroup_head.py:322-555 / tools/train.py via loss.backward()) -------- */

/* nbrT: i32[K][n_in], nbrT[k][i] = output row o with nbr[k][o] == i, -1 where none (a convolution's rule map holds an
 * input row at most once per tap).  nbr: i32[K][n_cols]; out_rows (may be NULL): the table is positional, column j
 * belongs to output row out_rows[j].  dX = cg3d_spconv_*(dY, nbrT, transposed weights). */
int cg3d_table_transpose(const int* nbr, int K, int n_cols, const int* out_rows, int n_in, int* nbrT, void* stream);

/* Wt[m][co][ci] = W[m][ci][co] for n_mats = G * K matrices (the weights dX is convolved with). */
int cg3d_transpose_weights(const float* W, int n_mats, int Cin, int Cout, float* Wt, void* stream);

/* dW[k] = sum over columns j in [col0, col1) with nbr[k][j] >= 0 of  in_act(x[nbr[k][j]])^T (x) dy[row(j)],
 * row(j) = out_rows ? out_rows[j] : j;  dW: f32[K][Cin][Cout].  nbr == NULL with K == 1: identity rows (1x1 conv /
 * Linear).  [col0, col1) selects the contiguous columns of one weight group of a grouped convolution (one call per
 * group).  Deterministic: chunks of columns write partial slabs that are added in order; `slabs` holds
 * cg3d_spconv_wgrad_slabs(col1 - col0, Cin, Cout, K) * K * Cin * Cout floats (may be NULL when that count is 1).
 * in_act: 0 none, 1 ReLU (the forward's in_act).  fp32 FFMA. */
int cg3d_spconv_wgrad_slabs(int n_cols, int Cin, int Cout, int K);
int cg3d_spconv_wgrad(const float* x, int ldx, int in_act, const int* nbr, const float* dy, int ldy, int n_cols, int col0,
                      int col1, int Cin, int Cout, int K, const int* out_rows, float* slabs, float* dW, void* stream);

/* The same sums on the tensor cores (tcgen05, pair index = contraction dimension, both operands MN-major) for
 * Cin % 64 == 0 and Cout % 64 == 0.  x_split / dy_split: the cg3d_split_bf16 images of in_act(x) (n_in rows, 2 Cin bf16) and
 * of dy (2 Cout bf16) -- the forward's operand and the dX launch's operand, so no extra pass; every bf16 product
 * (hi*hi, hi*lo, lo*hi, lo*lo) is accumulated in fp32.  slabs / dW / columns / determinism as cg3d_spconv_wgrad. */
int cg3d_spconv_wgrad_tc(const unsigned short* x_split, const int* nbr, const unsigned short* dy_split, int n_cols, int col0,
                         int col1, int Cin, int Cout, int K, const int* out_rows, float* slabs, float* dW, void* stream);

/* ---- training-mode BatchNorm and the backward of the row-gather ops (training; MinkowskiBatchNorm = BatchNorm1d over
 * the rows of a sparse tensor, biresnet.py:8-103; features_at_coordinates backward, biresnet.py:182-197,376-394;
 * UNWEIGHTED_AVERAGE quantisation backward, cagroup_head.py:257-271).  No float atomics: bit-repeatable. ---------- */

/* floats of `workspace` the two BatchNorm calls below need for an [n, C] matrix */
int cg3d_bn_train_workspace(long long n, int C);

/* Batch statistics of x: f32[n][C] (row stride ldx): mean[c], rstd[c] = 1 / sqrt(biased var + eps), and the folded
 * scale[c] = gamma[c] * rstd[c], shift[c] = beta[c] - mean[c] * scale[c] that cg3d_affine_act / the conv epilogues apply
 * (gamma / beta / scale / shift may be NULL).  running_mean / running_var (may be NULL) are updated as
 * torch.nn.BatchNorm1d does: r = (1 - momentum) * r + momentum * (mean | unbiased var).  n == 0 is an error, as in torch. */
int cg3d_bn_train_stats(const float* x, int ldx, long long n, int C, float eps, float momentum, const float* gamma,
                        const float* beta, float* workspace, float* mean, float* rstd, float* scale, float* shift,
                        float* running_mean, float* running_var, void* stream);

/* Backward of y = act(gamma * (x - mean) * rstd + beta (+ residual)):  dgamma[c], dbeta[c], dx: f32[n][C] (row stride lddx).
 * y_mask (may be NULL; row stride ldm): the forward's OUTPUT when act is ReLU -- dy counts only where y_mask > 0.
 * dres (may be NULL; row stride lddr): the gradient of the residual (= the masked dy). */
int cg3d_bn_train_backward(const float* x, int ldx, const float* dy, int ldy, const float* y_mask, int ldm, long long n, int C,
                           const float* mean, const float* rstd, const float* gamma, float* workspace, float* dx, int lddx,
                           float* dres, int lddr, float* dgamma, float* dbeta, void* stream);

/* Backward of cg3d_interp_trilinear when the query rows are ALL rows of a coordinate map of stride tq (tq divides ts), in
 * that map's row order (true for every call site of the backbone): dF[r] = sum over query voxels q at c_r + d,
 * d in (-ts, ts)^3, of prod_axis(1 - |d| / ts) * dOut[q].  src_coords: i32[n_src][4] rows of the stride-ts source map;
 * (qkeys, qvals, qcapacity): the query map's hash table; dOut: f32[nq][C]; dF: f32[n_src][C] (written, not added to). */
int cg3d_interp_trilinear_backward(const int* src_coords, int n_src, int ts, const unsigned long long* qkeys, const int* qvals,
                                   int qcapacity, int tq, const float* dOut, int C, float* dF, void* stream);

/* Backward of cg3d_segment_mean with ref == NULL: dIn[p] = dOut[inverse[p]] / counts[inverse[p]] (row stride ldi). */
int cg3d_segment_mean_backward(const float* dOut, const int* inverse, const float* counts, long long n, int C, float* dIn,
                               int ldi, void* stream);

/* Backward of cg3d_avgpool_window: dIn[i] = sum over output rows o with |c_i - c_o| <= half (same batch) of
 * dOut[o] / count[o], count[o] = the inputs in o's window (recomputed into `counts`, n_out floats).  dIn: f32[n_in][C]. */
int cg3d_avgpool_window_backward(const int* out_coords, int n_out, const int* in_coords, int n_in, int half, const float* dOut,
                                 int C, float* counts, float* dIn, void* stream);

/* dx = dy * act'(.) evaluated through the forward's OUTPUT y (MinkowskiReLU / MinkowskiELU backward): ReLU: y > 0;
 * ELU (alpha 1): y > 0 ? 1 : y + 1; act 0: copy.  Row strides ldy / ldyy / lddx. */
int cg3d_act_backward(const float* dy, int ldy, const float* y, int ldyy, long long n, int C, int act, float* dx, int lddx,
                      void* stream);

/* ---- first-stage training targets and the focal loss (cagroup_head.py:400-555 _loss_single / get_targets) ----------- */

/* CAGroup3DAssigner.assign (cagroup3d_assigner.py:62-133) for ONE sample.  locs: f32[n][3], the locations of all class
 * maps concatenated in class order; cls_offsets: i32[n_cls + 1], class c owns rows [cls_offsets[c], cls_offsets[c+1]);
 * gt_boxes: f32[m][7] (x, y, z, dx, dy, dz, yaw); gt_labels: i32[m].  A location is positive for the smallest-volume box of
 * its class that contains it and for which its centerness is above the box's (topk+1)-th best (first box index on ties).
 * Outputs: kth f32[m] (scratch: the per-box thresholds), centerness f32[n], box_targets f32[n][7], labels i64[n] (-1 =
 * negative), box_index i32[n] (may be NULL; index into gt_boxes of the chosen box, -1 when the class has no box). */
int cg3d_assign(const float* locs, int n, const int* cls_offsets, int n_cls, const float* gt_boxes, const int* gt_labels, int m,
                int topk, float* kth, float* centerness, float* box_targets, long long* labels, int* box_index, void* stream);

/* CAGroup3DAssigner.assign_semantic (cagroup3d_assigner.py:135-158): label of the smallest box containing the point (-1
 * outside all boxes) and instance label (box index + 1, 0 outside). */
int cg3d_assign_semantic(const float* points, int n, const float* gt_boxes, const int* gt_labels, int m, long long* labels,
                         long long* ins_labels, void* stream);

/* FocalLoss(use_sigmoid=True) (loss_utils.py:917-961,1012-1032): loss[0] = sum over [n][C] of the sigmoid focal loss /
 * avg_factor with labels i64[n] in [0, C) or negative / >= C for background; grad (may be NULL): f32[n][C] = d loss / d pred.
 * workspace: cg3d_focal_loss_workspace(n, C) floats (chunk partial sums, added in order). */
int cg3d_focal_loss_workspace(long long n, int C);
int cg3d_focal_loss(const float* pred, const long long* labels, long long n, int C, float gamma, float alpha, float avg_factor,
                    float* workspace, float* loss, float* grad, void* stream);

/* The other first-stage loss terms (cagroup_head.py:505-554), each with its gradient in the same pass; `workspace`:
 * cg3d_loss_workspace(rows) floats; loss: 1 float; grad may be NULL.
 *   bce:       CrossEntropy(use_sigmoid) (loss_utils.py:813-846) over n logits / targets in [0, 1] (target < 0: ignored),
 *              sum / (avg_factor + fp32 eps);
 *   iou_aa:    IoU3DLoss(with_yaw=False) (iou3d_loss.py:31-76): sum of weight * (1 - IoU) / avg_factor of axis-aligned
 *              (x, y, z, dx, dy, dz) boxes, rows strided by ldp / ldt; grad: first 6 columns of f32[n][ldg].  The caller
 *              skips the call when no weight is positive (the reference returns 0 then);
 *   smooth_l1: SmoothL1Loss(beta, reduction='sum') (loss_utils.py:1042-1074) over [n][C] with an [n][C] weight. */
int cg3d_loss_workspace(long long n);
int cg3d_bce_loss(const float* pred, const float* target, int n, float avg_factor, float* workspace, float* loss, float* grad,
                  void* stream);
int cg3d_iou_loss_aa(const float* pred, int ldp, const float* target, int ldt, const float* weight, int n, float avg_factor,
                     float* workspace, float* loss, float* grad, int ldg, void* stream);
int cg3d_smooth_l1_loss(const float* pred, const float* target, const float* weight, long long n, int C, float beta,
                        float* workspace, float* loss, float* grad, void* stream);

/* Vote targets from the per-point masks (cagroup_head.py:454-496, ScanNet branch).  scene_points: f32[n][ld] (xyz first);
 * sem_mask / ins_mask: i64[n]; instance ids in [0, n_inst); an instance whose FIRST point has a semantic label < n_classes
 * votes for the centre of the gt box (f32[m][7]) nearest to the centre of its own axis-aligned bounding box.  voxel_points:
 * f32[nv][3]; nearest: i32[nv], the k = 1 neighbour of every voxel among the scene points (cg3d_knn).  workspace: i32[n_inst * 8];
 * centers: f32[n_inst][3] (output: matched centre, -10000 for background, 0 for unused ids); targets: f32[nv][3] =
 * centre - voxel with components < -100 set to 0; mask: f32[nv] = 1 where no component was below -100. */
int cg3d_vote_targets(const float* scene_points, int ld, const long long* sem_mask, const long long* ins_mask, int n,
                      int n_inst, int n_classes, const float* gt_boxes, int m, const float* voxel_points, const int* nearest,
                      int nv, int* workspace, float* centers, float* targets, float* mask, void* stream);

/* out[c] = sum over the n rows of x[r][c] (row stride ldx): the gradient of a conv bias (semantic_conv / cls_conv,
 * cagroup_head.py:167,176).  workspace: cg3d_bn_train_workspace(n, C) floats; chunk partial sums added in order. */
int cg3d_column_sum(const float* x, int ldx, long long n, int C, float* workspace, float* out, void* stream);

/* out[u] = src[order[j]] summed over j in [seg_off[u], seg_off[u+1]) IN THAT ORDER (f32[n_seg][C]): the scatter-add of a
 * gather whose index map is not injective -- the dX of the RoI pooling contraction (cagroup_roi_head.py:72-90: two RoIs
 * can reach the same unique grid voxel at the same tap) -- with `order` a stable sort of the point ids by target row
 * (cg3d_sort_pairs) and seg_off the exclusive scan of cg3d_histogram_i32 of the targets.  No atomics. */
int cg3d_segment_sum_sorted(const float* src, const int* order, const int* seg_off, int n_seg, int C, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif
