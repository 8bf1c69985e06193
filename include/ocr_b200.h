/* ocr_b200.h -- C ABI of libocr_b200.so: hand-written sm_100a CUDA kernels for the hot path of
 * tgialoimtr/cnn_lstm_ctc_ocr (the weinman CNN -> BiLSTM/BiGRU -> CTC line recognizer).
 *
 * The reference has no FFI for this path: its boundary is Python functions that add nodes to a
 * TensorFlow graph.  Each entry point below therefore replaces one TensorFlow op call site in
 * the reference (cited per function); the Python host layer in cnn_lstm_ctc_ocr_b200/ keeps the
 * reference's function names on top of these.  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller unless it says "host";
 *   - the library never allocates or frees device memory and never synchronises the stream:
 *     all work is enqueued on `stream` (a CUstream / cudaStream_t handle passed as void*);
 *   - scratch memory comes from the caller: ask ocr_*_workspace_bytes first;
 *   - return value: OCR_OK (0) or a negative OCR_E* code; ocr_last_error() gives the text of the
 *     last failure on the calling thread;
 *   - tensors are dense row-major; logits are time-major [T,B,C] float32 as the reference's
 *     rnn_logits (src/weinman/model.py:212-221); the CTC blank is class C-1.
 */
#ifndef OCR_B200_H
#define OCR_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OCR_OK 0
#define OCR_EINVAL (-1)     /* bad argument (shape, null pointer, unsupported size) */
#define OCR_EWORKSPACE (-2) /* workspace too small */
#define OCR_ECUDA (-3)      /* CUDA launch/runtime error; see ocr_last_error() */
#define OCR_ENODEVICE (-4)  /* no sm_100 device */

typedef void* ocr_stream_t;

const char* ocr_last_error(void);
/* "major.minor;sm_100a;<build id>" */
const char* ocr_version(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches). */
uint64_t ocr_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * CTC loss + gradient.  Replaces tf.nn.ctc_loss at src/weinman/model.py:226-227 (defaults:
 * preprocess_collapse_repeated=False, ctc_merge_repeated=True, time_major=True) and the gradient
 * TensorFlow registers for it (d loss_b / d logits = softmax - posterior occupancy).
 *   logits        [T,B,C] f32
 *   labels        flat int32 label ids, example b owns labels[label_offsets[b] .. label_offsets[b+1])
 *   label_offsets [B+1] int32
 *   seq_len       [B] int32, 0 <= seq_len[b] <= T
 *   loss          [B] f32 out: -log p(labels_b | logits_b); +inf when no alignment exists
 *   grad          [T,B,C] f32 out (may be NULL): grad_scale * d loss_b / d logits[t,b,:], zero rows
 *                 for t >= seq_len[b]
 *   status        [B] int32 out (may be NULL): 0 ok; 1 no valid path (loss=+inf, grad=softmax, as
 *                 TF's "No valid path found"); 2 labels do not fit in seq_len (TF raises "Not enough
 *                 time for target transition sequence"; loss=0, grad=0 here and the host layer raises)
 *   grad_scale    1/B folds the reference's tf.reduce_mean (model.py:228) into the kernel
 *   max_label_len max over b of the label length (host-known; sizes shared memory)
 */
int ocr_ctc_loss_workspace_bytes(int T, int B, int C, int max_label_len, size_t* bytes);
int ocr_ctc_loss(const float* logits, int T, int B, int C, const int32_t* labels,
                 const int32_t* label_offsets, const int32_t* seq_len, int max_label_len, float* loss,
                 float* grad, int32_t* status, float grad_scale, void* workspace, size_t workspace_bytes,
                 ocr_stream_t stream);
/* Kernel-path override for tests and profiling: 0 = automatic (default), 1 = general kernel only
 * (one CTA per sequence, log-domain lattice), 2 = fast kernel with LSU loads/stores instead of TMA
 * bulk copies, 3 = fast kernel only (sequences it flags for the exact kernel keep status 100; diagnostics), 4 = fast kernel
 * with one TMA bulk copy per frame instead of one tensor-map request per 16 frames.
 * The automatic choice uses the fast kernel whenever its staging block fits in
 * shared memory and falls back to the general kernel for very long sequences. */
int ocr_ctc_loss_set_path(int path);
/* Tuning aid: per-warp clock64() stamps at the fast kernel's phase boundaries (12 int64 per warp, CTA-major);
 * pass NULL to switch it off.  Not part of the product path. */
int ocr_debug_ctc_timeline(long long* device_buffer);
/* Tuning aid: force the number of sequences per CTA of the fast kernel (1, 2, 4, 8; 0 = automatic). */
int ocr_debug_ctc_group(int G);
/* Tuning aid: L2 prefetch distance of the fast kernel in CTAs (-1 = automatic: half the resident CTAs of the grid, 0 = off). */
int ocr_debug_ctc_prefetch(int stride);
/* Tuning aid: programmatic dependent launch of the CTC kernels on (1, default) / off (0). */
int ocr_debug_ctc_pdl(int on);
/* Tuning aid: sequences the fast kernel flags are redone in its own tail (1, default) or by a second launch (0). */
int ocr_debug_ctc_inline_redo(int on);
/* Tuning aid: the fast kernel requests every TMA box of a group before its sequence lengths are known (1, default) or only
 * the boxes up to the group's longest sequence, after reading the lengths (0). */
int ocr_debug_ctc_speculate(int on);
/* Tuning aid: ring depth (16-frame boxes in flight per CTA) of the streaming CTC kernel (0 = automatic: 2, or the whole
 * sequence block when the grid is a single wave).  ocr_ctc_loss_set_path(7) keeps the streaming kernel out altogether. */
int ocr_debug_ctc_stream_nbuf(int n);

/* ---------------------------------------------------------------------------------------------
 * CTC greedy decoder.  Replaces tf.nn.ctc_greedy_decoder(merge_repeated=True) at
 * src/weinman/validate.py:86-90.  Per frame first-maximum arg-max, collapse repeats, drop blank.
 *   decoded        [B,T] int64 out, -1 padded (sparse_tensor_to_dense(default_value=-1), validate.py:91)
 *   decoded_len    [B] int32 out
 *   neg_sum_logits [B] f32 out: -sum_t max_k logits[t,b,k]  (the op's second output)
 */
int ocr_ctc_greedy_decode(const float* logits, int T, int B, int C, const int32_t* seq_len,
                          int merge_repeated, int64_t* decoded, int32_t* decoded_len,
                          float* neg_sum_logits, ocr_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * CTC beam search decoder.  Replaces tf.nn.ctc_beam_search_decoder at src/weinman/test.py:84-88
 * (beam_width=128, top_paths=1, merge_repeated=True) and src/weinman/client.py:227-231
 * (merge_repeated=False).
 *   normalize   1: per-frame log-softmax (newer TF); 0: per-frame max subtraction (older TF 1.x)
 *   decoded     [B,top_paths,T] int64 out, -1 padded
 *   decoded_len [B,top_paths] int32 out
 *   log_prob    [B,top_paths] f32 out
 * Limits: beam_width in {1..128}, top_paths <= beam_width, C <= 512.
 */
int ocr_ctc_beam_search_workspace_bytes(int T, int B, int C, int beam_width, size_t* bytes);
/* Tuning / cross-check aid: 0 = one CTA per sequence (default), 1 = one warp per sequence replaying TensorFlow's list updates
 * one insertion at a time.  Same bits either way (tests/test_beam_gpu.py). */
int ocr_debug_beam_path(int path);
/* Tuning aid: cycles per phase of CTA 0 of the CTA-per-sequence kernel, summed over its frames (host array of 16). */
int ocr_debug_beam_profile(long long* host16, int reset);
int ocr_ctc_beam_search(const float* logits, int T, int B, int C, const int32_t* seq_len, int beam_width,
                        int top_paths, int merge_repeated, int normalize, int64_t* decoded,
                        int32_t* decoded_len, float* log_prob, void* workspace, size_t workspace_bytes,
                        ocr_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Edit distance.  Replaces tf.edit_distance(hyp, truth, normalize=False) at src/weinman/test.py:90.
 *   hyp [B,hyp_stride] int64 (-1 padded or with hyp_len), truth flat int32 with offsets [B+1]
 *   dist [B] f32 out
 */
int ocr_edit_distance(const int64_t* hyp, int hyp_stride, const int32_t* hyp_len, const int32_t* truth,
                      const int32_t* truth_offsets, int B, int max_truth_len, float* dist,
                      ocr_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Dense contraction D[M,N] = act(A[M,K] * W[N,K]^T + bias[N]) on the tcgen05 tensor cores (TF32 products,
 * fp32 accumulation in TMEM, TMA-fed).  The shape of tf.layers.conv2d on im2col patches (model.py:97-104),
 * of the x-part of the tf.contrib.rnn cell kernels (model.py:173-192) and of tf.layers.dense (model.py:216-220).
 *   A [M,K] fp32 row-major, row pitch lda;  W [N,K] fp32 row-major (K-major weights), row pitch ldw;
 *   bias [N] or NULL;  D [M,N] fp32, row pitch ldd;  relu != 0 applies max(x,0).
 * A and W rows must be 16-byte aligned (pointer % 16 == 0, lda % 4 == 0, ldw % 4 == 0). */
/* Tuning aid, a bit mask.  Bit 0 (default set): the epilogue of ocr_gemm_tf32 stages 32x32 output blocks in shared memory and
 * stores them with cp.async.bulk.tensor (needs 16-byte aligned output rows); clear = one 16-byte STG per lane and row; same
 * bits either way.  Bit 1 (default clear): ocr_gemm_tf32_wgrad runs one tile per tap view; clear = the views of operands
 * with at most 64 rows are stacked in one 128-row tile (several taps per MMA). */
int ocr_debug_gemm_tma_store(int on);
int ocr_gemm_tf32(const float* A, int lda, const float* W, int ldw, const float* bias, float* D, int ldd, int M,
                  int N, int K, int relu, ocr_stream_t stream);
/* ocr_gemm_f16: ocr_gemm_tf32 with BINARY16 operands A [M, K], W [N, K] (row pitches in elements, multiples of 8; 16-byte
 * aligned): tcgen05.mma.kind::f16 with float32 sums, bias, output and epilogue as ocr_gemm_tf32.  binary16 keeps the 10 mantissa
 * bits TF32 keeps: for operands inside its range the products are the TF32 products at twice the tensor rate and half the
 * operand bytes.  Used by the training-side input projections of the recurrent layers (tf.nn.bidirectional_dynamic_rnn's
 * x * W_x, model.py:167-199); ocr_debug_proj_f16(0) selects TF32 there.  ocr_float_to_half: n (a multiple of 4) floats ->
 * binary16, round to nearest even. */
int ocr_float_to_half(const float* in, void* out, long long n, ocr_stream_t stream);
int ocr_gemm_f16(const void* A, int lda, const void* W, int ldw, const float* bias, float* D, int ldd, int M, int N, int K,
                 int relu, ocr_stream_t stream);
int ocr_debug_proj_f16(int on);

/* ---------------------------------------------------------------------------------------------
 * Recognizer layers around the GEMM (INFER mode; batch-norm folded into filters/biases by the caller).
 *
 * ocr_conv1_3x3_valid: conv1 of convnet_layers (model.py:47,134; 3x3 'valid', ONE input channel, ReLU) with
 *   validate._preprocess_image (validate.py:56-68: convert_image_dtype = u8 * float32(1/255), then - 0.5) fused in when
 *   in_is_u8 != 0.
 *   in [B,H,W] u8 or f32;  w [3,3,1,Cout] (HWIO);  out [B,H-2,W-2,Cout] f32 NHWC.  Cout % 4 == 0.
 * ocr_im2col3x3_same: patches of a 3x3 'same' convolution (model.py:97-104) taken from max_pool(in)
 *   (model.py:111-116; window pool_h x pool_w, strides stride_h x stride_w, 'valid'; 1,1,1,1 = no pooling):
 *   in [B,H,W,C] f32 NHWC -> out [B*Hp*Wp, 9*C], column order (kh, kw, c).  C % 4 == 0.
 * ocr_rows_max_to_seq: pool8 + squeeze + time-major transpose (model.py:145-147,212): in [B,H,W,C] -> out [W,B,C].
 * ocr_birnn_layer: one tf.nn.bidirectional_dynamic_rnn layer, time-major, per-example sequence_length
 *   (model.py:167-199 GRUCell; model_bu.py:167-199 LSTMCell): zero outputs past the length, the backward
 *   direction starts at frame len-1.  x [T,B,I] -> out [T,B,2H] (fw | bw).
 *   cell 0 = LSTM: wx [8H,I] (rows fw i,j,f,o then bw), wh [8H,H], bias [8H]; forget_bias 1.0; wh2 = NULL or the
 *            [8H,H] output of ocr_lstm_prepare_wh(wh) (weights pre-arranged for the persistent kernel, do it once)
 *   cell 1 = GRU : wx [6H,I] (rows per direction r,u,candidate), wh [4H,H] (r,u per direction), wh2 [2H,H], bias [6H] */
/* ocr_preprocess_train: the TRAINING-side preprocessing, mjsynth._preprocess_image (mjsynth.py:185-194) plus the zero
 *   padding the bucketing batcher applies to the PREPROCESSED tensor (mjsynth.py:56,69): in [B,Hin,W] u8 (right-padded
 *   with anything), widths [B] int32 -> out [B,Hin+1,W] f32 = u8 * float32(1/255) - 0.5 with the first row duplicated on
 *   top (31 -> 32 rows) and 0.0 at columns >= widths[b]. */
int ocr_preprocess_train(const unsigned char* in, int B, int Hin, int W, const int32_t* widths, float* out, ocr_stream_t stream);
int ocr_conv1_3x3_valid(const void* in, int in_is_u8, int B, int H, int W, const float* w, const float* bias, int Cout,
                        float* out, ocr_stream_t stream);
int ocr_im2col3x3_same(const float* in, int B, int H, int W, int C, int pool_h, int pool_w, int stride_h, int stride_w,
                       float* out, ocr_stream_t stream);
int ocr_rows_max_to_seq(const float* in, int B, int H, int W, int C, float* out, ocr_stream_t stream);
/* ocr_conv3x3_same: conv2..conv8 of convnet_layers (model.py:84-109; tf.layers.conv2d 3x3 'same' + bias, batch-norm
 *   folded, optional ReLU) as an implicit GEMM on tcgen05: the patches are gathered by cp.async straight into the
 *   swizzled operand tiles, no im2col matrix in memory.  in [B,H,W,C] f32 NHWC (C % 32 == 0), w [Cout, 9*C] K-major
 *   (column order kh, kw, c), out [B,H,W,Cout].
 * ocr_maxpool: tf.layers.max_pooling2d 'valid' (model.py:111-116): in [B,H,W,C] -> out [B,Hp,Wp,C]. */
int ocr_conv3x3_same(const float* in, int B, int H, int W, int C, const float* w, const float* bias, int Cout, int relu,
                     float* out, ocr_stream_t stream);
int ocr_maxpool(const float* in, int B, int H, int W, int C, int pool_h, int pool_w, int stride_h, int stride_w, float* out,
                ocr_stream_t stream);
/* ocr_conv3x3_same_pool: ocr_conv3x3_same followed by the max-pool that follows the layer in convnet_layers (window 2x2,
 * stride (2, stride_w), stride_w = 2: pool2, stride_w = 1: pool4 / pool6; 'valid'; model.py:105-116) as ONE launch: the pool
 * is taken in the epilogue of the halo-tile kernel and the unpooled activation never reaches memory.
 * out [B, (H-2)/2+1, stride_w == 2 ? (W-2)/2+1 : W-1, Cout].  Only shapes the halo-tile kernel takes (C = 32 / 64, H >= 12,
 * W >= 8): ocr_conv3x3_pool_fused returns 1 for them, otherwise run ocr_conv3x3_same + ocr_maxpool. */
int ocr_conv3x3_pool_fused(int B, int H, int W, int C, int Cout, int stride_w);
int ocr_conv3x3_same_pool(const float* in, int B, int H, int W, int C, const float* w, const float* bias, int Cout, int relu,
                          int stride_w, float* out, ocr_stream_t stream);
/* Tuning aid: TMA-store epilogues of the convolution kernels (32-pixel x 32-channel blocks staged in shared memory, one
 * cp.async.bulk.tensor per block) on (1, default) / off (0); same bits either way. */
int ocr_debug_conv_tma_store(int on);
/* Kernel-path override of ocr_conv3x3_same for tests: 0 = automatic (the wide shallow layers, C = 32 / 64 and H >= 12, run a
 * kernel that TMA-loads the input halo of a 16x8 pixel patch once and feeds all nine taps from it as shifted views; other
 * shapes gather each tap's patch rows with cp.async), 1 = gather kernel only, 2 = halo kernel only.  Same sums, bit for bit. */
int ocr_conv_set_path(int path);
int ocr_birnn_workspace_bytes(int cell, int T, int B, int H, size_t* bytes);
int ocr_lstm_prepare_wh(const float* wh, int H, float* wh_perm, ocr_stream_t stream);
/* Kernel-path override for tests: 0 = automatic (LSTM forward frames with B <= 256, H <= 512 run as ONE persistent
 * tcgen05 kernel for all frames; back-propagation through time likewise for B <= 64), 1 = one recurrent GEMM + one cell
 * kernel per frame everywhere, 2 = persistent forward only, 3 = persistent forward and persistent BPTT whenever the shape fits. */
int ocr_birnn_set_path(int path);
/* Tuning aid: per-frame clock64() stamps of CTA 0 of the persistent LSTM kernel (8 int64 per frame: producer past the grid
 * barrier, last h tile requested, first tile landed, last MMA issued, accumulator complete, TMEM read, cell update + stores
 * done, slice published); NULL switches it off.  Not part of the product path. */
int ocr_debug_lstm_timeline(long long* device_buffer);
/* Operand precision of the persistent recurrence kernels (LSTM and GRU): 1 (default) = IEEE binary16 h and W_h where
 * H % 64 == 0 (tcgen05.mma.kind::f16, K = 16 per instruction; the same 10 explicit mantissa bits a TF32 operand keeps),
 * 0 = TF32 operands.  Applies to weights prepared (ocr_lstm_prepare_wh) and layers run after the call. */
int ocr_debug_lstm_operands(int f16);
int ocr_birnn_layer(int cell, const float* x, int T, int B, int I, int H, const int32_t* seq_len, const float* wx,
                    const float* wh, const float* wh2, const float* bias, float* out, void* workspace,
                    size_t workspace_bytes, ocr_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Training step (src/weinman/train.py:101-141 _get_training; model.py with mode=TRAIN).  TensorFlow derives the
 * backward graph from the registered gradients of the ops model.py uses; each entry point below is one of those
 * gradients (or a forward op in its training form), so that the Python host layer (cnn_lstm_ctc_ocr_b200/train.py)
 * can replay train.py's step.  `scratch` arguments are caller-owned device memory of at least
 * 16 * max(channels, 9 * Cout) bytes unless stated otherwise.
 *
 * ocr_transpose            out[c, r] = in[r + src_shift, c] (zero outside [0, rows)); in [rows, cols] with row pitch ld_in,
 *                          out row pitch ld_out (elements).  src_shift = -+B gives the previous-frame hidden state of a
 *                          time-major [T*B, H] output.
 * ocr_nhwc_to_planar_pad   in [B,H,W,C] NHWC -> out [C, B*(H+2)*Wp] (row pitch ld_out), Wp = ocr_planar_pad_pitch(W) =
 *                          W+2 rounded up to a multiple of 4: channel-planar copy with a zero ring around every image,
 *                          the operand layout of the filter-gradient contraction.  ncopies = 3 writes three copies,
 *                          copy_stride elements apart, shifted by -1, 0, +1 pixels (copy k holds pixel r + k - 1 at r).
 * ocr_gemm_tf32_wgrad      D[batch] (M x N, row pitch ldd, batch_stride apart) =
 *                              sum_r At[a_row[batch] + m, r + a_shift[batch]] * Wt[n, r]:
 *                          the weight gradients d kernel = activations^T * d outputs of tf.layers.conv2d (model.py:97),
 *                          the RNN cells (model_bu.py:173-180) and tf.layers.dense (model.py:216) with the long pixel /
 *                          frame dimension R contiguous in both operands; split over R across the SMs (tcgen05, TF32),
 *                          partial tiles reduced in a fixed order.  a_shift / a_row are HOST arrays of nbatch (<= 9) ints
 *                          (NULL = zeros), At has a_rows rows in total (0 = M): the 3x3 taps are row-shifted views
 *                          (a_shift = +-Wp, a multiple of 4 as TMA requires) of the three pixel-shifted planar copies
 *                          (a_row = copy * C).  Reads outside [0, R) give zero.
 */
int ocr_transpose(const float* in, long long rows, int cols, int ld_in, float* out, long long ld_out, long long src_shift,
                  ocr_stream_t stream);
int ocr_planar_pad_pitch(int W);
int ocr_nhwc_to_planar_pad(const float* in, int B, int H, int W, int C, float* out, long long ld_out, int ncopies,
                           long long copy_stride, ocr_stream_t stream);
int ocr_gemm_wgrad_scratch_bytes(int M, int N, long long R, int nbatch, size_t* bytes);
int ocr_gemm_tf32_wgrad(const float* At, long long lda, const float* Wt, long long ldw, float* D, int ldd, long long batch_stride,
                        int M, int N, long long R, int nbatch, const int32_t* a_shift_host, const int32_t* a_row_host,
                        long long a_rows, void* scratch, size_t scratch_bytes, ocr_stream_t stream);
/* K-blocked operands for the filter gradients of the convolutions (the 128 bytes x rows a k-step needs are one contiguous run
 * instead of one run per channel plane, planes lying megabytes apart):
 *   ocr_planar_pad_pitch32          padded row pitch: W + 1 rounded up to 32 (one zero column between image rows is the right
 *                                   ring of one row and the left ring of the next; +- one pitch = whole blocks of 32 pixels)
 *   ocr_nhwc_to_planar_blocked      x [B,H,W,C] -> out [R/32][rows_total][32], R = B*(H+2)*pitch32, zero-ringed; copy k (pixel
 *                                   shift k-1, ncopies 1 or 3) in rows row0 + k*C ..
 *   ocr_gemm_tf32_wgrad_blocked     ocr_gemm_tf32_wgrad over such operands (At with a_rows rows, Wt with w_rows rows; R and
 *                                   every a_shift multiples of 32; bases 128-byte aligned) */
int ocr_planar_pad_pitch32(int W);
int ocr_nhwc_to_planar_blocked(const float* in, int B, int H, int W, int C, float* out, int rows_total, int row0, int ncopies,
                               ocr_stream_t stream);
int ocr_gemm_tf32_wgrad_blocked(const float* At, long long a_rows, const float* Wt, long long w_rows, float* D, int ldd,
                                long long batch_stride, int M, int N, long long R, int nbatch, const int32_t* a_shift_host,
                                const int32_t* a_row_host, void* scratch, size_t scratch_bytes, ocr_stream_t stream);
/* tf.layers.batch_normalization(training=True) (model.py:118-123) over y [rows, C] (rows = B*H*W of this replica):
 *   ocr_bn_batch_sums      sums [2][C] DOUBLES: per-channel sum and sum of squares (a data-parallel job may all-reduce them)
 *   ocr_bn_finalize        n = rows behind `sums` -> batch mean and 1/sqrt(biased variance + eps); moving_mean / moving_var
 *                          (both or neither NULL) updated in place with `momentum` (the UPDATE_OPS train.py:116-118 runs;
 *                          the moving variance takes the unbiased batch variance, as TensorFlow's fused kernel does)
 *   ocr_bn_relu_apply      out = relu(gamma * (y - mean) * inv_std + beta)       (model.py:108 norm_layer + relu)
 *   ocr_bn_relu_bwd_sums   dout = gradient w.r.t. that output -> sums [2][C] doubles (sum dz, sum dz*xhat), dgamma, dbeta [C]
 *   ocr_bn_relu_bwd_apply  dy [rows, C] = gamma*inv_std*(dz - sums[0]/n - xhat*sums[1]/n)
 *   ocr_copy_2d            strided device copy (weight-layout plumbing) */
int ocr_bn_batch_sums(const float* y, long long rows, int C, void* sums, ocr_stream_t stream);
int ocr_bn_finalize(const void* sums, long long n, int C, float eps, float momentum, float* mean, float* inv_std,
                    float* moving_mean, float* moving_var, ocr_stream_t stream);
int ocr_bn_relu_apply(const float* y, long long rows, int C, const float* mean, const float* inv_std, const float* gamma,
                      const float* beta, float* out, ocr_stream_t stream);
/* ocr_bn_relu_apply_pool: ocr_bn_relu_apply followed by ocr_maxpool(2, 2, 2, stride_w) (the pool_layer behind conv2 / conv4 /
 * conv6, model.py:105-116) in one pass over y: out [B,H,W,C] = relu(bn(y)) (kept: the pool's gradient needs it), pooled
 * [B, (H-2)/2+1, (W-2)/stride_w+1, C]. */
int ocr_bn_relu_apply_pool(const float* y, int B, int H, int W, int C, const float* mean, const float* inv_std, const float* gamma,
                           const float* beta, float* out, int stride_w, float* pooled, ocr_stream_t stream);
int ocr_bn_relu_bwd_sums(const float* y, const float* dout, long long rows, int C, const float* mean, const float* inv_std,
                         const float* gamma, const float* beta, void* sums, float* dgamma, float* dbeta, ocr_stream_t stream);
int ocr_bn_relu_bwd_apply(const float* y, const float* dout, long long rows, long long n, int C, const float* mean,
                          const float* inv_std, const float* gamma, const float* beta, const void* sums, float* dy,
                          ocr_stream_t stream);
/* ocr_bn_relu_bwd_apply_bias: ocr_bn_relu_bwd_apply that also returns dbias [C] = the per-channel sums of dy (the bias gradient of
 * the convolution in front of the batch-norm, what ocr_colsum(dy) gives) from the same pass; scratch: C doubles; C / 4 must
 * divide 256. */
int ocr_bn_relu_bwd_apply_bias(const float* y, const float* dout, long long rows, long long n, int C, const float* mean,
                               const float* inv_std, const float* gamma, const float* beta, const void* sums, float* dy,
                               float* dbias, void* scratch, ocr_stream_t stream);
/* The pooled batch-norm layers (conv2 / conv4 / conv6, model.py:105-116) without the full-size activation and without a tensor for
 * the pool's gradient.  C / 4 must divide 256; pool window 2x2, stride (2, stride_w), 'valid'.
 *   ocr_bn_relu_apply_pool_arg       pooled [B,Hp,Wp,C] = maxpool(relu(bn(y))) and arg [B,Hp,Wp,C/4] BYTES: per pooled element
 *                                    which window pixel (row-major 0..3) holds the first maximum -- TensorFlow's MaxPoolGrad
 *                                    rule -- two bits per channel, the four channels of a float4 in one byte
 *   ocr_bn_relu_bwd_sums_pool        ocr_bn_relu_bwd_sums with dout = MaxPoolGrad(dpooled) formed on the fly from arg
 *   ocr_bn_relu_bwd_apply_bias_pool  ocr_bn_relu_bwd_apply_bias, likewise (dy [B,H,W,C]; scratch: C doubles) */
int ocr_bn_relu_apply_pool_arg(const float* y, int B, int H, int W, int C, const float* mean, const float* inv_std, const float* gamma,
                               const float* beta, int stride_w, float* pooled, void* arg, ocr_stream_t stream);
int ocr_bn_relu_bwd_sums_pool(const float* y, const float* dpooled, const void* arg, int B, int H, int W, int C, int stride_w,
                              const float* mean, const float* inv_std, const float* gamma, const float* beta, void* sums,
                              float* dgamma, float* dbeta, ocr_stream_t stream);
int ocr_bn_relu_bwd_apply_bias_pool(const float* y, const float* dpooled, const void* arg, int B, int H, int W, int C, int stride_w,
                                    long long n, const float* mean, const float* inv_std, const float* gamma, const float* beta,
                                    const void* sums, float* dy, float* dbias, void* scratch, ocr_stream_t stream);
int ocr_copy_2d(const float* src, long long ld_src, float* dst, long long ld_dst, long long rows, long long cols, ocr_stream_t stream);
/* ReLU + bias-add gradients: dy = dout * (out > 0) (dy may alias dout), dbias[c] = sum_rows dy;  ocr_colsum: plain column
 * sums of x [rows, C] (row pitch ldx);  ocr_relu_bwd: dz = g * (z > 0) elementwise. */
int ocr_relu_bwd_bias(const float* out, const float* dout, long long rows, int C, float* dy, float* dbias, void* scratch,
                      ocr_stream_t stream);
int ocr_colsum(const float* x, long long rows, int C, int ldx, float* out, void* scratch, ocr_stream_t stream);
int ocr_relu_bwd(const float* z, const float* g, long long n, float* dz, ocr_stream_t stream);
/* max-pool gradients (model.py:111-116,145-147): the window's gradient goes to its first maximum. */
int ocr_maxpool_bwd(const float* in, const float* dout, int B, int H, int W, int C, int pool_h, int pool_w, int stride_h, int stride_w,
                    float* din, ocr_stream_t stream);
int ocr_rows_max_to_seq_bwd(const float* in, const float* dseq, int B, int H, int W, int C, float* din, ocr_stream_t stream);
/* conv1 filter gradient (one input channel: a reduction over pixels): dw [3,3,1,Cout]. */
int ocr_conv1_wgrad(const void* in, int in_is_u8, int B, int H, int W, const float* dy, int Cout, float* dw, void* scratch,
                    ocr_stream_t stream);
/* kernel-side filter layouts from the master HWIO tensor w [3,3,C,Cout]: w_fwd [Cout, 9*C] for ocr_conv3x3_same, w_dgrad
 * [C, 9*Cout] = the 180-degree rotated filter, so that d input = ocr_conv3x3_same(dy, w_dgrad).  Either may be NULL. */
int ocr_conv_filter_layouts(const float* w_hwio, int C, int Cout, float* w_fwd, float* w_dgrad, ocr_stream_t stream);
/* tf.train.AdamOptimizer.apply_gradients over one flat buffer (train.py:128-137); lr_t = lr*sqrt(1-b2^t)/(1-b1^t) from
 * the host, or read from lr_t_device (one float, device) when that is not NULL (a captured CUDA graph replays the
 * launch with a fresh step size); grads are multiplied by grad_scale first (1/world_size after a sum all-reduce). */
int ocr_adam_step(float* params, const float* grads, float* m, float* v, long long n, float lr_t, const float* lr_t_device,
                  float beta1, float beta2, float eps, float grad_scale, ocr_stream_t stream);
/* Bidirectional LSTM layer in training form (model_bu.py:167-199): like ocr_birnn_layer(cell 0) but keeps, per frame,
 * the gate activations (gates [T*B, 8H]: fw i, tanh j, f, o | bw ...) and cell states (cstate [T,B,2H]).
 * ocr_birnn_lstm_bwd: back-propagation through time.  dout [T,B,2H]; gates is overwritten with the gradient of the
 * gate pre-activations (zero past each example's length), from which d kernel / d bias / d input follow as dense
 * contractions.  wh_rows [2H, 4H]: rows I.. of the forward cell's TensorFlow kernel, then the backward cell's. */
/* Tuning aid for the frame-by-frame BPTT chain (B > 64), a bit mask; default 1.  Bit 0: programmatic dependent launch (same bits
 * either way).  Bit 1: float32 / TF32 operands in the recurrent product dG x W_h (clear, the default: bfloat16 operands with
 * float32 sums; gate gradients within 4e-4 of their max-norm of the TF32 product).  Bits 4..: tile width override (64/128/256). */
int ocr_debug_bptt_pdl(int on);
/* Tuning aid for the persistent BPTT kernel (B <= 64), a bit mask; default 1.  Bit 0: a batch of <= 32 (<= 64) rows is held four
 * (two) times in the 128-row operand tile so that all 128 epilogue threads share the partial-sum reads, the cell backward and
 * the scatter (same bits either way).  Bit 1: the partial sums cross L2 as float32 (clear, the default: bfloat16). */
int ocr_debug_bptt_copies(int on);
int ocr_birnn_lstm_train_workspace_bytes(int T, int B, int H, size_t* bytes);
int ocr_birnn_lstm_train_fwd(const float* x, int T, int B, int I, int H, const int32_t* seq_len, const float* wx, const float* wh,
                             const float* bias, float* out, float* gates, float* cstate, void* workspace, size_t workspace_bytes,
                             ocr_stream_t stream);
int ocr_birnn_lstm_bwd(const float* dout, int T, int B, int H, const int32_t* seq_len, float* gates, const float* cstate,
                       const float* wh_rows, void* workspace, size_t workspace_bytes, ocr_stream_t stream);

/* Bidirectional GRU layer in training form (model.py:167-199, tf.contrib.rnn.GRUCell: the reset gate multiplies h BEFORE the
 * candidate product) and its back-propagation through time.
 *   wx [6H, I] (rows per direction r, u, candidate), whg [4H, H] (r, u per direction), whc [2H, H], bias [6H] as ocr_birnn_layer;
 *   act [T*B, 6H] out: activations r | u | c per direction;  rh_all [T,B,2H] out: r * h_prev (operand of d candidate kernel).
 * ocr_birnn_gru_bwd: dout [T,B,2H]; act is overwritten with the gradients of the pre-activations (zero past each length);
 *   out = the layer's forward output; wg_rows [2H, 2H] / wc_rows [2H, H]: rows I.. of the gates / candidate kernels
 *   (TensorFlow layout), forward direction's H rows then the backward direction's. */
int ocr_birnn_gru_train_workspace_bytes(int T, int B, int H, size_t* bytes);
int ocr_birnn_gru_train_fwd(const float* x, int T, int B, int I, int H, const int32_t* seq_len, const float* wx, const float* whg,
                            const float* whc, const float* bias, float* out, float* act, float* rh_all, void* workspace,
                            size_t workspace_bytes, ocr_stream_t stream);
int ocr_birnn_gru_bwd(const float* dout, int T, int B, int H, const int32_t* seq_len, float* act, const float* out,
                      const float* wg_rows, const float* wc_rows, void* workspace, size_t workspace_bytes, ocr_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* OCR_B200_H */
