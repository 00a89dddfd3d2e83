/* instancerefer_b200 — C ABI of the B200-native (sm_100a) InstanceRefer hot path.
 *
 * This is the operator boundary SURVEY.md §8(b) defines: what a maintainer of
 * CurryYuan/InstanceRefer binds (ctypes) to replace the native code the
 * reference reaches through torchsparse / torch_cluster / torch_scatter / cuDNN on
 * InstanceRefer.forward (models/instancerefer.py:37-70).
 *
 * Conventions: every function returns IR_OK (0) or a negative error code and never throws;
 * `ir_last_error()` holds the message.  Nothing allocates: inputs, outputs and workspaces are
 * caller-owned DEVICE buffers (raw pointers + explicit sizes).  Everything is asynchronous on
 * `stream` (a cudaStream_t), with no host synchronisation and device-side counts.  sm_100a only;
 * there is no CPU fallback.
 */
#ifndef INSTANCEREFER_B200_H
#define INSTANCEREFER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IR_OK 0
#define IR_ERR_ARG (-1)
#define IR_ERR_CUDA (-2)
#define IR_ERR_UNSUPPORTED (-3)

typedef void* ir_stream_t; /* cudaStream_t */

int ir_version(void);
const char* ir_last_error(void);
/* IR_OK iff `device` is compute capability 10.x (B200). */
int ir_check_device(int device);

/* Number of kernels this library has launched in this process (bench.py's gpu_launches). */
int64_t ir_launch_count(void);
/* Measurement hooks: while enabled every sparse-conv layer brackets its pair-GEMM and its reduce
 * launch with CUDA events on the launching stream; ir_profile_read synchronises, returns per-layer
 * per-launch milliseconds and meta = (cin, cout, K, used_tcgen05) x n, and resets the record.
 * on = N > 1 launches each timed kernel N times back to back (idempotent) inside its event span so the
 * host launch gap of an eager launch is amortised; the reported time is the per-launch average. */
int ir_profile_enable(int on);
int ir_profile_read(float* gemm_ms, float* reduce_ms, int32_t* meta, int32_t cap, int32_t* n_out);

/* Live per-launch spans of the conv kernels (bench.py roofline): while buf != NULL every layer call made through
 * ir_encoder_features / ir_spconv_layer takes the next 4-u64 slot {pair-GEMM begin, end, reduce-or-stem begin, end} (GPU
 * timer ns; begin = first CTA past its dependency wait via atomicMin, end = last CTA done via atomicMax — the caller
 * presets begins to ~0 and ends to 0 before every run).  Slots are assigned at launch (or capture) time in call order, so a
 * CUDA graph captured while the buffer is set keeps writing its slots on every replay.  buf must hold 4 * 256 u64.
 * ir_conv_stamps_meta returns (cin, cout, K, used_tcgen05) per slot.  ir_conv_stamps_set(NULL) switches it off. */
int ir_conv_stamps_set(uint64_t* buf);
int ir_conv_stamps_meta(int32_t* meta, int32_t cap, int32_t* n_out);

/* Gather of the forward pair-GEMM: 0 = 16-byte global loads issued by the 12 producer warps (k_pairgemm_tc); 1 = TMA
 * (cp.async.bulk.tensor tile::gather4, one warp issuing 32 gathers of 4 rulebook rows x 128 B per stage) into a raw
 * fp32 shared-memory stage that the converter warps split into the fp16 hi/lo operand stages (k_pairgemm_tma). */
int ir_gather_mode_set(int mode);

/* Tuning knobs: CTAs per pair-GEMM launch (default 2 per SM = 296) and per reduce / stem launch (default
 * 5 per SM, swept inside the step); values <= 0 leave a knob unchanged.  Smaller grids let the two encoders' chains co-reside. */
int ir_tune_set(int pairgemm_ctas, int reduce_ctas);

/* Feature-pass mode of ir_encoder_features[_pair]: 0 (default) = one pair-GEMM + one reduce launch per layer, chained
 * by programmatic dependent launch; 1 = all 13 layers in ONE persistent launch (ticketed pair-GEMM / reduce items,
 * device-side tile counts, range-scaled split-fp16; not taken while ir_profile_enable is on or for the SIMT path). */
int ir_encoder_mode_set(int mode);

/* Profiling aid of the persistent encoder kernel: when buf != NULL every item (ticket t) of the following launches
 * writes GPU-timer stamps into buf[8*t ..]: 0 decoded, 1 weights staged, 2 dependency satisfied, 3 work done (thread 0),
 * 4 whole CTA done, 5 completion published, 7 = phase << 32 | CTA.  buf must hold 8 * 8192 u64.  NULL switches it off. */
int ir_encoder_persist_debug(uint64_t* buf);
/* Resident CTAs per SM of the persistent encoder kernel as computed by the CUDA runtime (2 expected), -1 on error. */
int ir_encoder_persist_occupancy(void);

/* Timeline aid: a one-thread kernel writes the GPU nanosecond timer into buf[idx] on `stream` (works inside stream
 * capture, so a replayed CUDA graph leaves a branch-level timeline behind; tools/timeline.py). */
int ir_debug_stamp(uint64_t* buf, int32_t idx, ir_stream_t stream);

/* Tuning aid for the tcgen05 pair-GEMM (tools/bench_spconv.py): bit0 skip gather loads, bit1 skip T
 * stores, bit2 skip MMA issue.  0 = normal operation. */
int ir_debug_set(int flags);

/* ------------------------------------------------------------------ sparse-voxel encoder
 * SparseConvEncoder / BEVEncoder (models/basic_blocks.py:59-95,136-171): 13 sparse convs
 * (stem k3; 4 x [k2 s2 down, residual k3 k3]) with eval-mode BatchNorm, ReLU and the identity
 * skip fused into each conv's reduce epilogue.  Layer order everywhere below:
 *   0 stem | 1 s1.down 2 s1.res_a 3 s1.res_b | 4..6 stage2 | 7..9 stage3 | 10..12 stage4      */
#define IR_ENC_LAYERS 13
#define IR_ENC_LEVELS 5

typedef struct {
    int32_t cin;                           /* input channels (7 = xyz+rgb+height)               */
    int32_t use_tc;                        /* 0: SIMT fp32 pair-GEMM, 1: tcgen05 split-fp16      */
    const float* weight[IR_ENC_LAYERS];    /* (K,Cin,Cout) fp32, reference `kernel` layout        */
    const float* wprep[IR_ENC_LAYERS];     /* ir_spconv_prepare_weights images (use_tc=1) or NULL */
    const float* bn_scale[IR_ENC_LAYERS];  /* gamma / sqrt(var+eps)                               */
    const float* bn_shift[IR_ENC_LAYERS];  /* beta - mean * scale                                 */
} ir_encoder_params;

/* Byte offsets inside an encoder workspace (for tests / tracing; all counts are device ints). */
typedef struct {
    int64_t n_max, cap, total_bytes;
    int64_t off_nlvl;                      /* int32[8]  : rows per level (stride 1,2,4,8,16)      */
    int64_t off_kcount;                    /* int32[9][32]: pairs per offset; maps 0-4 = k3 at
                                              level l, maps 5-8 = k2s2 from level l to l+1        */
    int64_t off_scan, scan_stride;         /* u64[5][scan_stride] compaction state                */
    int64_t off_sync;                      /* 512 B: ticket / phase counters / range maxima / time stamps of
                                              the persistent encoder kernel (inside the zeroed prefix)       */
    int64_t zero_bytes;                    /* bytes cleared from off_nlvl on reset                 */
    int64_t off_keys, off_vals;            /* 5 tables: keys u64[cap]; vals {minrow,row} i32[cap]  */
    int64_t off_coords[IR_ENC_LEVELS];     /* int32 (n_max,4) [x,y,z,b] per level                 */
    int64_t off_pslot;
    int64_t off_k3_in[IR_ENC_LEVELS], off_k3_slot[IR_ENC_LEVELS]; /* in_idx[27][n_max], slot[27][n_max] */
    int64_t off_k2_in[4], off_k2_slot[4];                         /* in_idx[8][n_max],  slot[8][n_max]  */
    int64_t off_feat0;                     /* fp32 (n_max, 8)  voxelised level-0 features          */
    int64_t off_feat[3];                   /* fp32 (n_max,128) rotating activations               */
    int64_t off_T;                         /* fp32 (27*n_max,128) pair products                   */
} ir_encoder_layout_t;

int ir_encoder_layout(int64_t n_max, ir_encoder_layout_t* out);
size_t ir_encoder_workspace_bytes(int64_t n_max);

/* Clears counters and hash tables of a workspace (3 memsets).  Call before ir_voxelize. */
int ir_encoder_reset(void* ws, int64_t n_max, ir_stream_t stream);

/* First-point-wins voxelisation of candidate instances (sparse_quantize + sparse_collate_tensors,
 * models/attribute_module.py:65-71,101): pts (n_inst, ppi, fdim) fp32; candidate m reads
 * instance cand[m]; coords = floor((double)xyz / voxel), batch index = m.  Output (level-0 coords,
 * features, count, hash table) stays inside `ws`; rows are in first-occurrence order.
 * Requires n_cand*ppi <= n_max. */
int ir_voxelize(const float* pts, const int32_t* cand, int32_t n_cand, int32_t ppi, int32_t fdim,
                double voxel, void* ws, int64_t n_max, ir_stream_t stream);

/* Loader-side voxelisation (sparse_quantize of whole scenes + sparse_collate batch index,
 * lib/dataset.py:255-261,456-469): clouds (n, ppi, fdim) fp32, cloud[m] = which cloud feeds batch index
 * m; first point wins, rows in first-occurrence order.  coords_out (n_cloud*ppi, 4) int32 [x,y,z,b],
 * feats_out (n_cloud*ppi, fdim), count_out device int32.  scratch: ir_voxelize_points_scratch_bytes. */
size_t ir_voxelize_points_scratch_bytes(int64_t n_pts);
int ir_voxelize_points(const float* pts, const int32_t* cloud, int32_t n_cloud, int32_t ppi, int32_t fdim,
                       double voxel, void* scratch, int32_t* coords_out, float* feats_out,
                       int32_t* count_out, ir_stream_t stream);

/* Builds levels 1-4 and all 9 kernel maps.  coords0==NULL: level 0 comes from ir_voxelize in
 * `ws`; otherwise coords0 (n0,4) int32 is hashed here (ir_encoder_reset is implied).  n0 is an
 * upper bound on the row count when n0_dev (device int, e.g. inside a CUDA graph) is given. */
int ir_encoder_build_maps(const int32_t* coords0, int32_t n0, const int32_t* n0_dev, void* ws,
                          int64_t n_max, ir_stream_t stream);

/* Feature pass over prebuilt maps.  feats0==NULL: use the voxelised features in `ws`.
 * feats_out (n_max,128) fp32; row count = nlvl[4] inside ws, coords = off_coords[4]. */
int ir_encoder_features(const ir_encoder_params* p, const float* feats0, void* ws, int64_t n_max,
                        float* feats_out, ir_stream_t stream);

/* The same feature pass for TWO encoders with identical topology (InstanceRefer: the candidate
 * encoder and the scene encoder) sharing every launch: each pair-GEMM / reduce kernel serves both
 * problems, halving the dependent-launch chain. */
int ir_encoder_features_pair(const ir_encoder_params* pa, const float* feats0a, void* wsa,
                             int64_t n_max_a, float* out_a, const ir_encoder_params* pb,
                             const float* feats0b, void* wsb, int64_t n_max_b, float* out_b,
                             ir_stream_t stream);

/* One sparse conv layer on explicit buffers (unit tests / tracing): rulebook = in_idx (K,seg_cap)
 * [input row of pair pos], slot (K,seg_cap) [pair pos of output row o, or -1], count (K). */
int ir_spconv_layer(const float* feat_in, int32_t cin, int32_t cout, int32_t K,
                    const int32_t* in_idx, int64_t seg_cap, const int32_t* slot,
                    const int32_t* count, const int32_t* n_out_dev, int64_t n_max,
                    const float* weight, const float* wprep, int32_t use_tc, const float* scale,
                    const float* shift, const float* resid, int32_t relu, float* T, float* out,
                    ir_stream_t stream);

/* ir_spconv_layer without BN epilogue whose tcgen05 pair-GEMM scales the gathered rows by the power of
 * two that brings *in_absmax (device scalar, max |feat_in|) to 2^13 before the fp16 hi/lo split and
 * scales the result back exactly: inputs of any magnitude (gradients: this is the dgrad of the training
 * step, run on the transposed rulebook with W^T) keep ~22 significant bits relative to their maximum. */
int ir_spconv_layer_scaled(const float* feat_in, const float* in_absmax, int32_t cin, int32_t cout, int32_t K,
                           const int32_t* in_idx, int64_t seg_cap, const int32_t* slot, const int32_t* count,
                           const int32_t* n_out_dev, int64_t n_max, const float* weight, int32_t use_tc,
                           const float* resid, float* T, float* out, ir_stream_t stream);

/* Weight buffer the tcgen05 pair-GEMM reads: a 16-byte aligned copy of the reference layout
 * (K,Cin,Cout); W[k] is staged by one TMA bulk copy per CTA and split into fp16 hi / lo parts on its
 * way into TMEM.  out holds ir_spconv_wprep_floats(K,cin,cout) floats. */
int64_t ir_spconv_wprep_floats(int32_t K, int32_t cin, int32_t cout);
int ir_spconv_prepare_weights(const float* weight, int32_t K, int32_t cin, int32_t cout,
                              float* out, ir_stream_t stream);

/* spnn.GlobalMaxPooling (models/attribute_module.py:105): out[m,:] = max over rows with b==m.
 * enc_scratch: uint32 (n_seg, C) scratch. */
int ir_segmax(const float* feats, const int32_t* coords, const int32_t* n_dev, int64_t n_max,
              int32_t C, int32_t n_seg, uint32_t* enc_scratch, float* out, ir_stream_t stream);

/* ------------------------------------------------------------------ scene head
 * SparseCrop + ToDenseBEVConvolution + BatchNorm2d + ReLU (models/basic_blocks.py:174-243,
 * models/scene_module.py:25-30): keep 0<=xyz<(240,400,80); f' = f @ kernel[z/stride];
 * dense[b, x/stride, y/stride, :] = relu(bn(sum f')), NHWC (B,15,25,128).  tmp: (n_max,128) fp32,
 * cell: int32 (n_max).  out_absmax (or NULL): device float receiving max|out| (atomicMax on the bit pattern; the caller
 * zeroes it) — the range scale of the tcgen05 Conv2d behind it. */
int ir_bev(const float* feats, const int32_t* coords, const int32_t* n_dev, int64_t n_max,
           int32_t stride, const float* kernel, const float* bn_scale, const float* bn_shift,
           int32_t B, float* tmp, int32_t* cell, float* out, float* out_absmax, ir_stream_t stream);

/* Conv2d 3x3, no padding, NHWC activations, weight repacked to [ky][kx][Cin][Cout];
 * y = act(scale*(conv + bias) + shift).  (models/scene_module.py:33-38) */
int ir_conv2d_3x3(const float* in, int32_t B, int32_t H, int32_t W, int32_t C, const float* wpack,
                  const float* bias, const float* scale, const float* shift, int32_t relu,
                  float* out, ir_stream_t stream);

/* The same Conv2d on the tcgen05 rule GEMM: a 3x3 valid convolution over a dense grid is a sparse convolution with a
 * closed-form rulebook (K = 9; in_idx[k][o] = input pixel of tap k = ky*3+kx of output pixel o, slot[k][o] = o,
 * count[k] = n_out, all device int32 built once per grid shape by the caller), weight repacked to (9, Cin, Cout) =
 * [ky][kx][Cin][Cout], 16-byte aligned.  y = act(scale * conv + shift) (the conv bias folded into shift by the caller);
 * in_absmax / out_absmax as in the sparse layers (range-scaled split-fp16; NULL = unscaled / not recorded).
 * T: fp32 (9 * n_out, 128) scratch.  C = 128 only.  (models/scene_module.py:33-38) */
int ir_conv2d_3x3_tc(const float* in, const int32_t* in_idx, const int32_t* slot, const int32_t* count,
                     const int32_t* n_out_dev, int64_t n_out, const float* wprep, const float* scale,
                     const float* shift, int32_t relu, const float* in_absmax, float* out_absmax, float* T,
                     float* out, ir_stream_t stream);

/* Language-guided attention over BEV cells (models/scene_module.py:73-83):
 * atten = softmax_cells(feats . q / sqrt(C)); scene_feat = sum atten * feats. */
int ir_scene_attention(const float* feats, const float* q, int32_t B, int32_t ncell, int32_t C,
                       float* atten, float* scene_feat, ir_stream_t stream);

/* ------------------------------------------------------------------ language
 * y = act(x W^T + b) (nn.Linear; models/lang_module.py:33-37 and the hoisted GRU input GEMMs). */
int ir_linear(const float* x, int32_t M, int32_t K, const float* W, const float* b, int32_t N,
              int32_t relu, float* y, ir_stream_t stream);

/* One bidirectional GRU layer over packed sequences (models/lang_module.py:53-57): xproj
 * (B,L,2,3H) = x W_ih^T + b_ih per direction; whh (2,3H,H); bhh (2,3H); out (B,L,2H), zeros at
 * t >= len[b]; the reverse direction starts at each sample's own last token.  H = 128. */
int ir_gru_layer(const float* xproj, const float* whh, const float* bhh, const int64_t* lengths,
                 int32_t B, int32_t L, int32_t H, float* out, ir_stream_t stream);

/* Four masked attention poolings over tokens (models/lang_module.py:60-83): logits = fc(feats);
 * softmax over L, * mask, renormalise; pooled = atten @ embed.  fcw (4,D), fcb (4);
 * atten (4,B,L), pooled (4,B,E). */
int ir_token_attention(const float* feats, const float* embed, int64_t embed_stride,
                       const int64_t* lengths, const float* fcw, const float* fcb, int32_t B,
                       int32_t L, int32_t D, int32_t E, float* atten, float* pooled,
                       ir_stream_t stream);

/* ------------------------------------------------------------------ matching heads
 * Fused 2-layer head: h = relu(norm(x W1 + b1)); y = h W2 + b2 with W1 (K,N1), W2 (N1,N2) in
 * (in,out) layout (transposed nn.Linear weights, coalesced reads), then
 *   mode 0: write y;  1: write y/max(|y|,1e-12);
 *   2: score[r] = <y/max(|y|,1e-12), partner[seg[r]]>;  3: score[r] = cos(y, partner[seg[r]]), eps 1e-8.
 * norm 0: none, 1: per-channel affine (eval BatchNorm1d), 2: LayerNorm(eps 1e-5).
 * (models/attribute_module.py:88-90,108-126; relation_module.py:82,101-103;
 *  scene_module.py:44-57,84-104).  Dims <= 256. */
int ir_mlp_head(const float* x, int32_t M, int32_t K, const float* W1, const float* b1, int32_t N1,
                int32_t norm, const float* g, const float* beta, const float* W2, const float* b2,
                int32_t N2, int32_t mode, const float* partner, const int32_t* seg, float* y,
                float* score, ir_stream_t stream);

/* Per-scene softmax / argmax over candidates of the summed score (extra fused output; the
 * reference sums and arg-maxes on the host, lib/eval_helper.py:61-67).  seg_ofs (n_seg+1). */
int ir_candidate_softmax(const float* s_attr, const float* s_rel, const float* s_scene,
                         const int32_t* seg_ofs, int32_t n_seg, float* prob, int32_t* argmax,
                         ir_stream_t stream);

/* ------------------------------------------------------------------ relation
 * Per-instance column means of (n_inst, ppi, fdim) point blocks (models/relation_module.py:67). */
int ir_instance_mean(const float* pts, int32_t n_inst, int32_t ppi, int32_t fdim, float* mean,
                     ir_stream_t stream);

/* Brute-force kNN inside scene segments (torch_cluster.knn semantics, models/basic_blocks.py:120):
 * for query q (row qidx[q] of support) the k nearest support rows of the same segment, ascending
 * squared distance, ties -> lower index; nbr (nq,k) int32, -1 padded.  seg_ofs (n_seg+1) row
 * offsets of each scene in `xyz`; qseg (nq) scene of each query. */
int ir_knn(const float* xyz, const int32_t* seg_ofs, const int32_t* qidx, const int32_t* qseg,
           int32_t nq, int32_t k, int32_t* nbr, ir_stream_t stream);

/* DynamicEdgeConv message + max aggregation (models/basic_blocks.py:125-133): x (S,F) with the
 * class one-hot in the last `ncls` columns; per edge j->i: w = W2w relu(W1w [p_j-p_i, oh_i, oh_j]),
 * msg = W2m relu(W1m [x_i, w, x_j]); out[i] = max_j msg (0 if no edge).  F<=32, hidden 64/128.
 * All four weight matrices are passed in (in,out) layout, i.e. transposed nn.Linear weights. */
int ir_edgeconv(const float* x, const float* xyz, const int32_t* qidx, const int32_t* nbr,
                int32_t nq, int32_t k, int32_t F, int32_t ncls, const float* Ww1, const float* bw1,
                const float* Ww2, const float* bw2, const float* Wm1, const float* bm1,
                const float* Wm2, const float* bm2, int32_t Fout, float* out, ir_stream_t stream);

/* ================================================================== training step
 * Backward of the sparse encoders, train-mode BatchNorm, loss and optimiser.  What torchsparse's
 * sparseconv_backward, torch autograd and torch.optim.Adam do for lib/solver.py:196-205.          */

/* Transposed rulebook of one kernel map: out_idx[k][pos] = output row of pair pos,
 * slot_in[k][i] = pair of input row i (or -1).  dgrad of a sparse conv is ir_spconv_layer run on
 * (out_idx, slot_in) with the weight transposed to (K,Cout,Cin): dX = sum_k dY[o] @ W[k]^T.       */
int ir_rulebook_transpose(const int32_t* in_idx, const int32_t* slot, int32_t K, int64_t seg_cap,
                          const int32_t* n_out_dev, int64_t n_max, int32_t* out_idx, int32_t* slot_in,
                          ir_stream_t stream);

/* dW[k] = sum over pairs of offset k of x[in_idx]^T dy[out_idx]   (K,Cin,Cout), overwritten.      */
int ir_spconv_wgrad(const float* x, int32_t cin, const float* dy, int32_t cout, int32_t K,
                    const int32_t* in_idx, const int32_t* out_idx, const int32_t* count,
                    int64_t seg_cap, float* dW, ir_stream_t stream);

/* The same on tcgen05 for Cin, Cout in {64,128}: both gathered operands are MN-major UMMA operands (the
 * pair index is the contraction), split-fp16 hi/lo with dy range-scaled by *dy_absmax (max |dy|, device
 * scalar); other shapes, use_tc = 0 or dy_absmax = NULL fall through to ir_spconv_wgrad.            */
int ir_spconv_wgrad_scaled(const float* x, int32_t cin, const float* dy, const float* dy_absmax, int32_t cout,
                           int32_t K, const int32_t* in_idx, const int32_t* out_idx, const int32_t* count,
                           int64_t seg_cap, int32_t use_tc, float* dW, ir_stream_t stream);

/* Train-mode BatchNorm over the rows of a (n, C) matrix (spnn.BatchNorm over voxels, BatchNorm1d,
 * BatchNorm2d on NHWC cells): batch mean / biased variance, y = act((x-mean)*rstd*gamma + beta
 * (+resid)); running statistics updated with `momentum` (unbiased variance) when given.
 * n_dev (optional) = device row count bounded by n.  scratch: float[ir_bn_scratch_floats(C)]
 * (per-CTA partial sums; deterministic, no atomics).  C divides 256, C >= 4.                     */
int64_t ir_bn_scratch_floats(int32_t C);
int ir_bn_train_fwd(const float* x, const int32_t* n_dev, int32_t n, int32_t C, const float* gamma,
                    const float* beta, const float* resid, int32_t relu, float eps, float momentum,
                    float* running_mean, float* running_var, float* scratch, float* mean,
                    float* rstd, float* y, ir_stream_t stream);
/* g = dy*[y>0] (relu); dbeta = sum g; dgamma = sum g*xhat; dx; dresid = g (optional);
 * absmax_out (optional device float) = max |dx|, the range hint of ir_spconv_layer_scaled.        */
int ir_bn_train_bwd(const float* dy, const float* y, const float* x, const int32_t* n_dev, int32_t n,
                    int32_t C, const float* mean, const float* rstd, const float* gamma, int32_t relu,
                    float* scratch, float* dx, float* dresid, float* dgamma, float* dbeta,
                    float* absmax_out, ir_stream_t stream);

/* Backward of ir_segmax: the gradient of out[b,c] goes to the first row attaining the maximum.
 * arg_scratch: int32 (n_seg, C). */
int ir_segmax_bwd(const float* feats, const int32_t* coords, const int32_t* n_dev, int64_t n_max,
                  int32_t C, int32_t n_seg, const float* pooled, const float* dpooled,
                  int32_t* arg_scratch, float* dfeats, ir_stream_t stream);

/* Mean cross-entropy over B rows and its gradient (lib/loss_helper.py:155,189-193). B <= 1024.    */
int ir_cross_entropy(const float* logits, const int64_t* labels, int32_t B, int32_t N, float* loss,
                     float* dlogits, ir_stream_t stream);

/* 9-way region label of compute_scene_mask_loss (lib/loss_helper.py:131-153); (B,3) fp64 inputs,
 * inputs_were_f32 keeps fp32 rounding of the thirds when every operand was fp32 in the caller.    */
int ir_region_label(const double* ref_center, const double* point_min, const double* point_max,
                    int32_t B, int32_t inputs_were_f32, int64_t* label, ir_stream_t stream);

/* Reference loss of get_loss (lib/loss_helper.py:225-260): per scene b, IoU of its candidate boxes
 * pred_obb[obb_ofs[b]:obb_ofs[b+1]] (7 doubles each) with gt_obb[b] (utils/box_util.py:154-198,
 * 310-333), label = one-hot of the first maximal IoU; when the scene has >= 2 candidates
 * (score_ofs[b] >= 0 = its offset in the score vectors) and max IoU >= iou_thresh:
 * loss_scene[b] = ContrastiveLoss(margin, gamma)(s_attr+s_rel+s_scene, label) (:93-107) and
 * dscore = d loss_scene / d score.  ref_loss = sum(loss_scene) / B.                               */
int ir_ref_loss(const double* pred_obb, const int32_t* obb_ofs, const double* gt_obb,
                const int32_t* score_ofs, int32_t B, const float* s_attr, const float* s_rel,
                const float* s_scene, float margin, float gamma, double iou_thresh, float* label,
                float* loss_scene, float* dscore, float* iou_max, ir_stream_t stream);

/* get_eval (lib/eval_helper.py:11-114): per scene the candidate with the highest summed score, IoU of
 * its box with gt_obb[b] (utils/box_util.py:95-133), ref_acc (arg-max == IoU label for >= 2 candidates,
 * IoU > 0.25 otherwise; a scene without candidates scores a zero box) and the corner boxes of
 * construct_bbox_corners (utils/util.py:21-32), (B,8,3) fp64.  label = ir_ref_loss's one-hot labels. */
int ir_ref_eval(const double* pred_obb, const int32_t* obb_ofs, const double* gt_obb,
                const int32_t* score_ofs, int32_t B, const float* s_attr, const float* s_rel,
                const float* s_scene, const float* label, int32_t* pred_idx, float* ref_acc,
                double* iou, double* pred_corners, double* gt_corners, ir_stream_t stream);

/* torch.optim.Adam (amsgrad off) on one flat fp32 buffer; gradient = grad_scale*g + weight_decay*p
 * (grad_scale = 1/world_size after a sum all-reduce).  `step` counts from 1.  block_skip (or NULL): one byte per
 * 64 floats, 1 = the block belongs to a parameter that received no gradient this step and is left untouched
 * (parameter and both moments), as torch.optim.Adam does for grad None.                                    */
int ir_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                 float lr, float beta1, float beta2, float eps, float weight_decay, int32_t step,
                 float grad_scale, const uint8_t* block_skip, ir_stream_t stream);
/* The same update for a launch replayed from a CUDA graph (instancerefer_b200/train_graph.py): the three scalars
 * that change per step — lr (lib/solver.py:119-125 steps a scheduler on it), 1-beta1^step, sqrt(1-beta2^step) — are
 * computed on the host by ir_adam_hyper exactly as ir_adam_step does, copied to hyper_dev (float[3]) by the caller
 * before the replay, and read on the device.                                                                  */
int ir_adam_hyper(float lr, float beta1, float beta2, int32_t step, float* hyper3);
int ir_adam_step_dev(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                     const float* hyper_dev, float beta1, float beta2, float eps, float weight_decay,
                     float grad_scale, const uint8_t* block_skip, ir_stream_t stream);

/* ------------------------------------------------------------------ one-call encoder training passes
 * All 13 [conv -> train-mode BN (-> + skip) -> ReLU] layers forward, and their backward, as one chain
 * of launches each.  Activations (pre-BN y, post-activation out, batch mean / rstd per layer), the
 * transposed rulebooks and gradient scratch live in a caller-owned arena.  n_lvl = HOST row counts of
 * the five levels (read back once from the workspace after ir_encoder_build_maps), or NULL: every buffer
 * is laid out for the capacity n_max and every kernel reads its row count on the device, so the launch
 * sequence does not depend on the input and can be captured ONCE per capacity into a CUDA graph.   */
typedef struct {
    int32_t cin, use_tc;                      /* bit 0: forward pair-GEMM on tcgen05 (weights 16-B aligned);
                                                 bit 1: dgrad on tcgen05 (range-scaled, ir_spconv_layer_scaled);
                                                 bit 2: wgrad on tcgen05 (ir_spconv_wgrad_scaled)             */
    const float* weight[IR_ENC_LAYERS];       /* (K,Cin,Cout)                                          */
    const float* gamma[IR_ENC_LAYERS];
    const float* beta[IR_ENC_LAYERS];
    float* running_mean[IR_ENC_LAYERS];       /* updated in place by the forward (or NULL)             */
    float* running_var[IR_ENC_LAYERS];
    float momentum[IR_ENC_LAYERS];
    float eps;
} ir_encoder_train_params;

typedef struct {                              /* outputs of the backward, caller-owned                  */
    float* dweight[IR_ENC_LAYERS];
    float* dgamma[IR_ENC_LAYERS];
    float* dbeta[IR_ENC_LAYERS];
} ir_encoder_train_grads;

typedef struct {
    int64_t total_bytes;
    int64_t off_y[IR_ENC_LAYERS], off_out[IR_ENC_LAYERS];      /* fp32 (n_lvl[level], Cout)              */
    int64_t off_mean[IR_ENC_LAYERS], off_rstd[IR_ENC_LAYERS];
    int64_t off_bn_scratch;
    int64_t off_tr_out[9], off_tr_slot[9];                     /* transposed rulebooks, int32 [K][n_max]  */
    int64_t off_grad[6];                                       /* fp32 (max rows, 128) gradient buffers: 0, 2, 3 carry the
                                                                * layer-to-layer gradients, 1, 4, 5 are a ring of dY buffers
                                                                * (the wgrad of a layer may lag the chain by two layers)  */
    int64_t off_wt;                                            /* fp32 transposed weights (K,Cout,Cin), layers 1..12 */
    int64_t off_absmax;                                        /* fp32 scalar: max |dY| of the current layer */
} ir_encoder_train_layout_t;

int ir_encoder_train_layout(int64_t n_max, const int32_t* n_lvl, int32_t cin, ir_encoder_train_layout_t* out);
/* feats0 == NULL: level-0 features voxelised into `ws`.  The encoder output is arena + off_out[12]. */
int ir_encoder_train_forward(const ir_encoder_train_params* p, const float* feats0, void* ws, int64_t n_max,
                             const int32_t* n_lvl, void* arena, ir_stream_t stream);
/* dout (n_lvl[4],128) = gradient w.r.t. the encoder output; same arena as the forward call, which also left the
 * transposed rulebooks and W^T of this step in it (built on a helper stream beside the 13 layers).   */
int ir_encoder_train_backward(const ir_encoder_train_params* p, const float* feats0, void* ws, int64_t n_max,
                              const int32_t* n_lvl, void* arena, const float* dout,
                              const ir_encoder_train_grads* g, ir_stream_t stream);
/* Stages stage_hi..stage_lo of the same pass (4..1 = residual stages from the output down, 0 = stem; (4,0) is the whole
 * backward): the deep stages hold 2/3 of an encoder's parameters and finish long before the large shallow levels, so a
 * data-parallel caller runs (4,3), starts the all-reduce of those gradients (lib/solver.py:200-205 under DDP), then (2,0).
 * dout is read only when stage_hi = 4; the gradient between two calls stays in the arena.                          */
int ir_encoder_train_backward_range(const ir_encoder_train_params* p, const float* feats0, void* ws, int64_t n_max,
                                    const int32_t* n_lvl, void* arena, const float* dout,
                                    const ir_encoder_train_grads* g, int32_t stage_hi, int32_t stage_lo,
                                    ir_stream_t stream);

/* ------------------------------------------------------------------ dense training-step operators
 * (nn.Linear / LayerNorm / Dropout / F.normalize / cosine heads, Conv2d as im2col + GEMM, and the
 * backward passes of ir_bev, ir_scene_attention, ir_token_attention, ir_gru_layer, ir_edgeconv.)   */

/* C (M,N) = op(A) op(B) [+ bias(N)] [+ C] [relu];  A(m,k) = trans_a ? A[k*lda+m] : A[m*lda+k],
 * B(k,n) = trans_b ? B[n*ldb+k] : B[k*ldb+n].  fp32 SIMT (latency-bound sizes).                  */
int ir_gemm(int32_t M, int32_t N, int32_t K, const float* A, int32_t lda, int32_t trans_a,
            const float* B, int32_t ldb, int32_t trans_b, float* C, int32_t ldc, const float* bias,
            int32_t relu, int32_t accumulate, ir_stream_t stream);
/* out[c] = sum_r x[r,c]  (bias gradients; fixed summation order). */
int ir_colsum(const float* x, int32_t M, int32_t N, float* out, ir_stream_t stream);
/* dx = dy * [y > 0]. */
int ir_relu_bwd(const float* dy, const float* y, int64_t n, float* dx, ir_stream_t stream);
/* nn.Dropout(p) in train mode with a counter-based generator: mask[i] = u(seed,i) >= p,
 * y = x*mask/(1-p); backward applies the stored mask. */
int ir_dropout_fwd(const float* x, int64_t n, float p, uint64_t seed, float* y, uint8_t* mask,
                   ir_stream_t stream);
int ir_dropout_bwd(const float* dy, const uint8_t* mask, int64_t n, float p, float* dx, ir_stream_t stream);
/* Process-wide: every later dropout launch (ir_dropout_fwd and the fused heads / language / scene-tail passes) also
 * folds *step_dev (device uint64, or NULL to switch off) into its seed on the device — a step replayed from a CUDA
 * graph then draws a fresh mask although its host seed was fixed at capture.                                  */
int ir_dropout_seed_step(const uint64_t* step_dev);
/* nn.LayerNorm(N) (+ReLU) over the rows of (M,N) and its backward (dgamma/dbeta overwritten). */
int ir_layernorm_fwd(const float* x, int32_t M, int32_t N, const float* gamma, const float* beta, float eps,
                     int32_t relu, float* y, float* mean, float* rstd, ir_stream_t stream);
int ir_layernorm_bwd(const float* dy, const float* y, const float* x, int32_t M, int32_t N, const float* gamma,
                     const float* mean, const float* rstd, int32_t relu, float* dx, float* dgamma,
                     float* dbeta, ir_stream_t stream);
/* F.normalize(x, p=2, dim=1) and its backward. */
int ir_l2norm_fwd(const float* x, int32_t M, int32_t N, float* y, ir_stream_t stream);
int ir_l2norm_bwd(const float* dy, const float* x, int32_t M, int32_t N, float* dx, ir_stream_t stream);
/* score[r] vs partner[seg[r]]: mode 0 <a/max(|a|,1e-12), p> (models/attribute_module.py:113,126),
 * mode 1 cosine_similarity eps 1e-8 (relation_module.py:103, scene_module.py:104).  Backward: da and
 * dpartner (n_partner,N); row_ofs (n_partner+1) = contiguous row range of every partner row. */
int ir_match_fwd(const float* a, const float* partner, const int32_t* seg, int32_t M, int32_t N, int32_t mode,
                 float* score, ir_stream_t stream);
int ir_match_bwd(const float* dscore, const float* a, const float* partner, const int32_t* seg,
                 const int32_t* row_ofs, int32_t M, int32_t N, int32_t n_partner, int32_t mode, float* da,
                 float* dpartner, ir_stream_t stream);
/* Conv2d 3x3 (valid, NHWC) as GEMM: col (B*(H-2)*(W-2), 9*C) with column (ky,kx,c); col2im is the
 * adjoint (gradient w.r.t. the input image). */
int ir_im2col_3x3(const float* in, int32_t B, int32_t H, int32_t W, int32_t C, float* col, ir_stream_t stream);
int ir_col2im_3x3(const float* dcol, int32_t B, int32_t H, int32_t W, int32_t C, float* din, ir_stream_t stream);
/* Backward of ir_bev in raw mode (bn_scale = bn_shift = NULL: plain sums, no BN/ReLU): ddense
 * (B*375,128) -> dfeats (n,128), dkernel (n_z,128,128).  cell = the buffer ir_bev filled. */
int ir_bev_bwd(const float* ddense, const float* feats, const int32_t* coords, const int32_t* cell,
               const int32_t* n_dev, int64_t n_max, int32_t stride, const float* kernel, int32_t n_z,
               float* dfeats, float* dkernel, ir_stream_t stream);
/* Backward of ir_scene_attention: dscene (B,C) [+ datten (B,ncell) or NULL] -> dfeats, dq. */
int ir_scene_attention_bwd(const float* feats, const float* q, const float* atten, const float* dscene,
                           const float* datten, int32_t B, int32_t ncell, int32_t C, float* dfeats, float* dq,
                           ir_stream_t stream);
/* Backward of ir_token_attention w.r.t. feats, embed and the four fc layers; dfcw_part (B,4,D) and
 * dfcb_part (B,4) are per-sample partials (sum over B with ir_colsum). */
int ir_token_attention_bwd(const float* feats, const float* embed, int64_t embed_stride, const int64_t* lengths,
                           const float* fcw, const float* fcb, const float* atten, const float* dpooled,
                           int32_t B, int32_t L, int32_t D, int32_t E, float* dfeats, float* dembed,
                           float* dfcw_part, float* dfcb_part, ir_stream_t stream);
/* Backward through time of ir_gru_layer: dout (B,L,2H) -> dxproj (B,L,2,3H), plus per step
 * dhp (B,L,2,3H) = gradient w.r.t. W_hh h + b_hh and hprev (B,L,2,H) = the step's input state, so that
 * dW_hh[d] = dhp[:,:,d,:]^T @ hprev[:,:,d,:] (ir_gemm) and db_hh = column sums of dhp (ir_colsum). */
int ir_gru_layer_bwd(const float* xproj, const float* whh, const float* bhh, const int64_t* lengths,
                     const float* out, const float* dout, int32_t B, int32_t L, int32_t H, float* dxproj,
                     float* dhp, float* hprev, ir_stream_t stream);
/* Train-mode EdgeConv pieces (models/basic_blocks.py:125-133): w == NULL builds w_in (E, 3+2*ncls) =
 * [xyz_j - xyz_i, onehot_i, onehot_j]; otherwise e_in (E, 3F) = [x_i, w, x_j]; E = nq*k, rows of
 * missing neighbours (nbr < 0) are zero.  ir_edge_max_*: max over the valid edges of a query. */
int ir_edge_inputs(const float* x, const float* xyz, const int32_t* qidx, const int32_t* nbr, int32_t nq,
                   int32_t k, int32_t F, int32_t ncls, const float* w, float* w_in, float* e_in,
                   ir_stream_t stream);
int ir_edge_max_fwd(const float* msg, const int32_t* nbr, int32_t nq, int32_t k, int32_t C, float* out,
                    int32_t* arg, ir_stream_t stream);
int ir_edge_max_bwd(const float* dout, const int32_t* arg, int32_t nq, int32_t k, int32_t C, float* dmsg,
                    ir_stream_t stream);

/* Fused two-layer head of the matching modules in train mode: Linear(K->N1) -> {norm 1: BatchNorm1d with batch
 * statistics | norm 2: LayerNorm} -> ReLU [-> Dropout(drop_p, seed)] -> Linear(N1->N2), ONE call per direction
 * (models/attribute_module.py:22-32, relation_module.py:13-25, scene_module.py:40-58).  Weights in nn.Linear
 * layout: w1 (N1,K), w2 (N2,N1).  The arena keeps the intermediates for the backward. */
typedef struct {
    int32_t M, K, N1, N2, norm;
    float eps, momentum, drop_p;
    uint64_t seed;
    const float *w1, *b1, *gamma, *beta, *w2, *b2;
    float *running_mean, *running_var;          /* BatchNorm only (updated in place) */
} ir_mlp_head_t;
int64_t ir_mlp_head_arena_bytes(int32_t M, int32_t N1);
int ir_mlp_head_train_fwd(const ir_mlp_head_t* h, const float* x, void* arena, float* y, ir_stream_t stream);
int ir_mlp_head_train_bwd(const ir_mlp_head_t* h, const float* x, void* arena, const float* dy, float* dx,
                          float* dw1, float* db1, float* dgamma, float* dbeta, float* dw2, float* db2,
                          ir_stream_t stream);

/* Language encoder in train mode (models/lang_module.py:51-108) as ONE call per direction: word MLP
 * (Linear-ReLU-Dropout-Linear-ReLU on the B*L live tokens) -> 2-layer packed biGRU (H = 128, D = 2H = 256) -> four
 * masked attention pools over the projected embeddings -> classifier on pooled[1].  All weights in the reference's
 * nn layouts; [layer][direction] for the GRU tensors.  Outputs: pooled (4,B,D) [attr, cls, rel, scene], scores
 * (B,n_cls); the GRU output (lang_feat) and the attention maps stay in the arena (ir_lang_train_view).  */
typedef struct {
    int32_t B, L, E_in, D, H, n_cls;
    float drop_p;
    uint64_t seed;
    const float *w0, *b0, *w3, *b3;
    const float *wih[2][2], *bih[2][2], *whh[2][2], *bhh[2][2];
    const float *fcw[4], *fcb[4];
    const float *wc, *bc;
} ir_lang_t;
typedef struct {
    float *dw0, *db0, *dw3, *db3;
    float *dwih[2][2], *dwhh[2][2];
    float *dbih[2], *dbhh[2];                 /* (2,3H) contiguous per layer: [direction][gate row] */
    float *dfcw, *dfcb;                       /* (4,D), (4) contiguous: fc_a, fc_cls, fc_rel, fc_scene */
    float *dwc, *dbc;
} ir_lang_grads_t;
int64_t ir_lang_train_arena_bytes(const ir_lang_t* p);
int ir_lang_train_fwd(const ir_lang_t* p, const float* x, const int64_t* lengths, void* arena, float* pooled,
                      float* scores, ir_stream_t stream);
int ir_lang_train_bwd(const ir_lang_t* p, const float* x, const int64_t* lengths, void* arena, const float* pooled,
                      const float* dpooled, const float* dscores, const ir_lang_grads_t* g, ir_stream_t stream);
/* Byte offsets inside the arena of the GRU output (B,L,2H) and the attention maps (4,B,L). */
int ir_lang_train_view(const ir_lang_t* p, int64_t* off_feats, int64_t* off_atten);

/* DynamicEdgeConv in train mode given the kNN lists (models/basic_blocks.py:98-133), ONE call per direction:
 * w = W2w relu(W1w [p_j-p_i, onehot_i, onehot_j]); msg = W2m relu(W1m [x_i, w, x_j]); out_i = max over the valid
 * edges.  Weights in nn.Linear layout: ww1 (H1, 3+2*ncls), ww2 (F, H1), wm1 (Fout, 3F), wm2 (Fout, Fout).  x / xyz
 * carry no gradient; the backward produces the eight parameter gradients. */
typedef struct {
    int32_t nq, k, F, ncls, H1, Fout;
    const float *ww1, *bw1, *ww2, *bw2, *wm1, *bm1, *wm2, *bm2;
} ir_edgeconv_t;
typedef struct { float *dww1, *dbw1, *dww2, *dbw2, *dwm1, *dbm1, *dwm2, *dbm2; } ir_edgeconv_grads_t;
int64_t ir_edgeconv_train_arena_bytes(const ir_edgeconv_t* p);
int ir_edgeconv_train_fwd(const ir_edgeconv_t* p, const float* x, const float* xyz, const int32_t* qidx,
                          const int32_t* nbr, void* arena, float* out, ir_stream_t stream);
int ir_edgeconv_train_bwd(const ir_edgeconv_t* p, void* arena, const float* dout, const ir_edgeconv_grads_t* g,
                          ir_stream_t stream);

/* Scene tail in train mode (models/scene_module.py:25-38,70-71), ONE call per direction: SparseCrop +
 * ToDenseBEVConvolution -> BatchNorm2d (batch statistics) -> ReLU -> Conv2d 3x3 -> BatchNorm2d -> ReLU -> Dropout ->
 * Conv2d 3x3.  f4 (n_rows,128) / coords / n_dev = the stride-16 output of the scene encoder; out (B,11,21,128) NHWC.
 * Conv weights in the reference layout (Cout,Cin,3,3); kernel (5,128,128). */
typedef struct {
    int32_t B;
    int64_t n_rows;
    float eps, mom0, mom1, drop_p;
    uint64_t seed;
    const float *kernel, *g0, *be0, *w1, *b1, *g1, *be1, *w2, *b2;
    float *rm0, *rv0, *rm1, *rv1;
} ir_scene_tail_t;
typedef struct { float *dkernel, *dg0, *dbe0, *dw1, *db1, *dg1, *dbe1, *dw2, *db2; } ir_scene_tail_grads_t;
int64_t ir_scene_tail_arena_bytes(int64_t n_rows, int32_t B);
int ir_scene_tail_train_fwd(const ir_scene_tail_t* p, const float* f4, const int32_t* coords, const int32_t* n_dev,
                            void* arena, float* out, ir_stream_t stream);
int ir_scene_tail_train_bwd(const ir_scene_tail_t* p, const float* f4, const int32_t* coords, const int32_t* n_dev,
                            void* arena, const float* dout, float* df4, const ir_scene_tail_grads_t* g,
                            ir_stream_t stream);

/* ------------------------------------------------------------------ scan preprocessing
 * The array work of data/scannet/prepare_data.py:30-216; the host side (PLY/JSON/TSV parsing, .npy output) is
 * instancerefer_b200/prepare_data.py.  Vertex rows are (n,9) fp32: xyz, rgb (0-255), normal.  All pointers are
 * device pointers unless named *_host.  scratch: ir_prepare_scratch_bytes(n_verts, n_faces, n_objects) bytes. */
int64_t ir_prepare_scratch_bytes(int64_t n_verts, int64_t n_faces, int32_t n_objects);
/* scannet_utils.py:18-44,111-115 (compute_normal, incl. its last-write-wins accumulation): fills columns 6..8. */
int ir_mesh_normals(float* verts9, int64_t n_verts, const int32_t* faces, int64_t n_faces, void* scratch,
                    ir_stream_t stream);
/* prepare_data.py:60-66: xyz <- fp32([x y z 1] . M^T) computed in fp64; the other six columns are copied. */
int ir_align_vertices(const float* verts9, int64_t n_verts, const double* matrix16_host, float* aligned9,
                      ir_stream_t stream);
/* prepare_data.py:73-90: label_ids[v] = seg_label[seg[v]], instance_ids[v] = seg_object[seg[v]] (0 = unannotated);
 * the dense per-segment tables are built on the host from the aggregation JSON (later groups overwrite). */
int ir_vertex_labels(const int32_t* seg_of_vert, int64_t n_verts, const int32_t* seg_label,
                     const int32_t* seg_object, int32_t n_seg_table, uint32_t* label_ids, uint32_t* instance_ids,
                     ir_stream_t stream);
/* prepare_data.py:92-131: boxes (n_objects,8) fp64 = (cx,cy,cz,dx,dy,dz,label,obj_id-1) of the vertices with
 * instance id o+1, fp32 arithmetic as numpy's; an object without vertices keeps a zero row. */
int ir_instance_boxes(const float* verts9, const uint32_t* instance_ids, int64_t n_verts, int32_t n_objects,
                      const int32_t* obj_label, void* scratch, double* boxes, ir_stream_t stream);
/* prepare_data.py:141-148: masks (n_inst, n_verts) u8 of the PointGroup proposals in list order, cls (n_inst);
 * the last proposal covering a vertex wins; ids are 1-based, 0 = none. */
int ir_pointgroup_labels(const uint8_t* masks, const int32_t* cls, int32_t n_inst, int64_t n_verts,
                         uint32_t* label_ids_pg, uint32_t* instance_ids_pg, ir_stream_t stream);
/* prepare_data.py:185 (np.logical_not(np.in1d(labels, DONOTCARE_CLASS_IDS))): ascending indices of the kept
 * vertices -> idx, their number -> *count_dev. */
int ir_keep_index(const uint32_t* sem_labels, int64_t n_verts, const int32_t* donotcare, int32_t n_donotcare,
                  void* scratch, int64_t* idx, int64_t* count_dev, ir_stream_t stream);
/* prepare_data.py:186-189,206-212 (boolean / fancy row indexing): dst[i] = src[idx[i]] for i < min(*count_dev,
 * m_max) (count_dev may be NULL); row_bytes a multiple of 4. */
int ir_gather_rows(const void* src, int64_t row_bytes, const int64_t* idx, const int64_t* count_dev, int64_t m_max,
                   void* dst, ir_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif
