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
 * bulk copies, 3 = fast kernel only (sequences it flags for the exact kernel keep status 100; diagnostics).
 * The automatic choice uses the fast kernel whenever its staging block fits in
 * shared memory and falls back to the general kernel for very long sequences. */
int ocr_ctc_loss_set_path(int path);
/* Tuning aid: per-warp clock64() stamps at the fast kernel's phase boundaries (12 int64 per warp, CTA-major);
 * pass NULL to switch it off.  Not part of the product path. */
int ocr_debug_ctc_timeline(long long* device_buffer);

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
int ocr_gemm_tf32(const float* A, int lda, const float* W, int ldw, const float* bias, float* D, int ldd, int M,
                  int N, int K, int relu, ocr_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Recognizer layers around the GEMM (INFER mode; batch-norm folded into filters/biases by the caller).
 *
 * ocr_conv1_3x3_valid: conv1 of convnet_layers (model.py:47,134; 3x3 'valid', ONE input channel, ReLU) with
 *   validate._preprocess_image (validate.py:56-68: u8/255 - 0.5) fused in when in_is_u8 != 0.
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
int ocr_birnn_workspace_bytes(int cell, int T, int B, int H, size_t* bytes);
int ocr_lstm_prepare_wh(const float* wh, int H, float* wh_perm, ocr_stream_t stream);
/* Kernel-path override for tests: 0 = automatic (LSTM layers with B <= 128, H <= 512 run as ONE persistent
 * tcgen05 kernel for all frames), 1 = one recurrent GEMM + one cell kernel per frame. */
int ocr_birnn_set_path(int path);
int ocr_birnn_layer(int cell, const float* x, int T, int B, int I, int H, const int32_t* seq_len, const float* wx,
                    const float* wh, const float* wh2, const float* bias, float* out, void* workspace,
                    size_t workspace_bytes, ocr_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* OCR_B200_H */
