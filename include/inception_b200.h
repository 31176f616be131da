/* inception_b200.h — C ABI of libinception_b200.so
 *
 * B200-native (sm_100a) replacement for the arithmetic the reference runs inside one
 * `tf.Session.run` call:  /root/reference/PyClassification.py:208-209
 *     probs, labels = self.sess.run([self.values, self.indices], {self.image_buffer: image})
 * i.e. preprocessing (PyClassification.py:151-156,180-203) -> Inception-v3 tower
 * (inception/slim/inception_model.py:86-329, built by inception/inception_model.py:48-93)
 * -> softmax -> tf.nn.top_k(predictions, 5) (PyClassification.py:163-164).
 *
 * The reference has no FFI of its own (pure Python over the tensorflow wheel); these entry points
 * are what a maintainer binds with ctypes from `ImagenetClassifier` — see INTEGRATION.md.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success or a negative
 * ICP_E_* code and never throws; the message for the last failure on an engine (or, for
 * icp_create / the stateless calls, on the calling thread) is available from icp_last_error().
 * An engine is thread-compatible (one call at a time), distinct engines are independent.
 * All images are NHWC.  There is no CPU fallback: without a CUDA device every compute call fails
 * with ICP_E_CUDA.
 */
#ifndef INCEPTION_B200_H_
#define INCEPTION_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ICP_OK 0
#define ICP_E_INVALID (-1) /* bad argument */
#define ICP_E_IO (-2)      /* weight blob unreadable / malformed / tensor missing */
#define ICP_E_CUDA (-3)    /* CUDA runtime or driver error (includes "no device") */
#define ICP_E_NOMEM (-4)

#define ICP_PRECISION_FP16 0 /* fp16 storage, fp32 accumulate: meets the 1e-3 score parity bar */
#define ICP_PRECISION_BF16 1 /* bf16 storage, fp32 accumulate: throughput mode of BASELINE configs[2] */

typedef struct icp_engine icp_engine;

typedef struct icp_config {
  const char* weights_path; /* ICPW blob keyed by the reference's variable names (weights.py) */
  int num_classes;          /* reference: NUM_CLASSES + 1 (PyClassification.py:20,160) */
  int device;               /* CUDA ordinal */
  int max_batch;            /* largest n accepted by the forward calls */
  int precision;            /* ICP_PRECISION_* */
  int image_size;           /* network input side; 0 = 299 (PyClassification.py:153) */
  int enable_aux;           /* also run the aux_logits head (inception_model.py:249-261) */
  int use_cuda_graph;       /* capture the tower per batch size and replay it */
} icp_config;

/* Replaces ImagenetClassifier.__init__'s graph build + Saver.restore (PyClassification.py:145-172). */
int icp_create(const icp_config* cfg, icp_engine** out);
void icp_destroy(icp_engine* e);
const char* icp_last_error(const icp_engine* e); /* e may be NULL: thread-local message */

/* Tower on host float32 images already preprocessed to [-1,1] (the tensor `images` of
 * PyClassification.py:156): x[n,S,S,3] -> probs_out[n,C] (softmax, inception_model.py:329).
 * Host<->device copies are part of the call. */
int icp_forward(icp_engine* e, const float* x_host, int n, float* probs_out_host);

/* Same, plus top-k of the probabilities (PyClassification.py:163-164): vals[n,k] descending,
 * idx[n,k] int32; k <= min(num_classes, 16).  probs_out_host may be NULL. */
int icp_forward_topk(icp_engine* e, const float* x_host, int n, int k, float* probs_out_host,
                     float* topk_vals_host, int32_t* topk_idx_host);

/* Tower on host uint8 images already at network size (x[n,S,S,3], 0..255): the reference's
 * convert_image_dtype and (x-0.5)*2 (PyClassification.py:152,154-155) are fused into the first
 * conv; a quarter of icp_forward's host->device bytes.  Outputs as icp_forward_topk (k may be 0). */
int icp_forward_u8(icp_engine* e, const uint8_t* x_host, int n, int k, float* probs_out_host,
                   float* topk_vals_host, int32_t* topk_idx_host);

/* Asynchronous two-slot variant for serving loops: icp_submit_u8 enqueues host->device copy (on a copy
 * stream), tower and result copies for `slot` (0 or 1) and returns at once; icp_collect blocks until that slot's
 * outputs are in the host buffers given at submit.  With both slots in use the PCIe copy of batch i+1 overlaps
 * the tower of batch i.  x_host and the output buffers should be pinned and must stay valid until collect. */
int icp_submit_u8(icp_engine* e, const uint8_t* x_host, int n, int k, float* probs_out_host,
                  float* topk_vals_host, int32_t* topk_idx_host, int slot);
int icp_collect(icp_engine* e, int slot);
/* Multi-GPU variant: the probabilities stay on the device (probs_dev float32 [n,C], caller-owned, on the engine's
 * device) so that a collective can consume them without a host round trip; work is ordered on icp_stream(e). */
int icp_submit_u8_device_out(icp_engine* e, const uint8_t* x_host, int n, float* probs_dev, int slot);

/* Per-launch profile of the tower: op i of icp_num_ops() is one kernel launch.  icp_op_info gives its
 * reference scope name, kind (ICP_OP_*), algorithmic FLOPs and algorithmic HBM bytes per image;
 * icp_profile_forward runs the tower `reps` times on the engine's resident input with a CUDA event
 * between consecutive launches and returns the mean device milliseconds of each. */
#define ICP_OP_CONV0 0
#define ICP_OP_CONV_IGEMM 1
#define ICP_OP_MAXPOOL 2
#define ICP_OP_AVGPOOL 3
#define ICP_OP_HEAD 4
int icp_num_ops(icp_engine* e, int* num_ops);
int icp_op_info(icp_engine* e, int i, char* name, int name_cap, int* kind, double* flops_per_image,
                double* bytes_per_image);
int icp_profile_forward(icp_engine* e, int n, int reps, float* ms_per_op);

/* Device-resident variant for batching / multi-GPU callers: x_dev float32 [n,S,S,3] and probs_dev
 * float32 [n,C] are device pointers on the engine's device; work is enqueued on `stream`
 * (a cudaStream_t, NULL = the engine's own stream) and NOT synchronised. */
int icp_forward_device(icp_engine* e, const float* x_dev, int n, float* probs_dev, void* stream);
/* Same with the uint8 entry: x_dev uint8 [n,S,S,3] (0..255, already at network size) on the engine's device — the
 * batch a GPU decoder or a previous device stage left in HBM; the (x/255 - 0.5)*2 of PyClassification.py:152-155 is
 * folded into conv0 as in icp_forward_u8. */
int icp_forward_u8_device(icp_engine* e, const uint8_t* x_dev, int n, float* probs_dev, void* stream);

/* Page-lock (and later release) a host buffer the caller will pass to the batch / submit calls repeatedly — e.g. the
 * shared-memory segments through which the serving workers hand decoded pixels to the GPU-owner process — so that
 * the per-image host->device copies are true asynchronous DMA instead of staged pageable copies.  Registered on the
 * engine's device (portable to every context).  ICP_E_CUDA if the range cannot be registered; the buffer stays usable. */
int icp_host_register(icp_engine* e, void* host_ptr, size_t bytes);
int icp_host_unregister(icp_engine* e, void* host_ptr);

/* One request as the reference serves it (PyClassification.py:205-215) minus JPEG entropy
 * decoding: rgb is the decoded uint8 [h,w,3] image; crop, legacy-bilinear resize and
 * normalisation run on the GPU, then the tower, softmax and top-k. */
int icp_classify_rgb8(icp_engine* e, const uint8_t* rgb, int h, int w, int k, float* topk_vals,
                      int32_t* topk_idx);

/* One request exactly as the reference serves it — the JPEG string in, top-k out
 * (PyClassification.py:150-151 tf.image.decode_jpeg(channels=3) ... :163-164 top_k; :205-215): nvJPEG decodes on
 * the GPU into the buffer the crop/resize kernel reads, then tower, softmax, top-k.  Baseline and progressive
 * JPEGs; grayscale is expanded to 3 channels.  ICP_E_INVALID for anything that is not a decodable JPEG. */
int icp_classify_jpeg(icp_engine* e, const uint8_t* jpg, size_t len, int k, float* topk_vals, int32_t* topk_idx);

/* Decode alone (parity tap against libjpeg): rgb_out_host (may be NULL to query the size) receives uint8 [h,w,3]. */
int icp_decode_jpeg(icp_engine* e, const uint8_t* jpg, size_t len, uint8_t* rgb_out_host, size_t cap, int* h_out,
                    int* w_out);
/* nvjpegBackend_t value of the decoder in use (-1 before the first JPEG call). */
int icp_jpeg_backend(icp_engine* e, int* backend);

/* Micro-batched requests (serving front-end): n images of different sizes -> ONE tower launch at batch n.
 * rgb8: decoded uint8 [hs[i], ws[i], 3] host buffers; jpeg: JPEG byte strings decoded on the GPU.
 * topk_vals / topk_idx are [n, k].  n <= max_batch. */
int icp_classify_batch_rgb8(icp_engine* e, const uint8_t* const* rgbs, const int* hs, const int* ws, int n, int k,
                            float* topk_vals, int32_t* topk_idx);
int icp_classify_batch_jpeg(icp_engine* e, const uint8_t* const* jpgs, const size_t* lens, int n, int k,
                            float* topk_vals, int32_t* topk_idx);

/* GPU preprocessing alone (parity tap): rgb uint8 [h,w,3] -> out float32 [S,S,3]. */
int icp_preprocess_rgb8(icp_engine* e, const uint8_t* rgb, int h, int w, float* out_host);

/* Debug taps = the reference's end_points dict (inception_model.py:78-329) for the batch of the
 * most recent forward call.  Keys: conv0..conv4, pool1, pool2, mixed_*, logits, predictions,
 * aux_logits.  icp_endpoint_shape fills dims[4] = {n,h,w,c} (h=w=1 for vectors). */
int icp_endpoint_shape(icp_engine* e, const char* name, int dims[4]);
int icp_endpoint(icp_engine* e, const char* name, float* out_host);

/* Device time of the most recent forward call (CUDA events around the tower), milliseconds. */
int icp_last_forward_ms(icp_engine* e, float* ms);
/* Number of this library's kernels enqueued by one forward pass at batch n. */
int icp_launches_per_forward(icp_engine* e, int* launches);
/* The engine's stream (cudaStream_t) and its device-side staging buffers, for callers that time
 * or feed the tower themselves (bench.py). */
void* icp_stream(icp_engine* e);

/* Stateless single layer: ops.conv2d = conv (no bias) + folded BN + optional ReLU
 * (inception/slim/ops.py:229-245) through the same tcgen05 kernel the tower uses.
 * x[n,h,w,cin] f32, w[kh,kw,cin,cout] f32 (HWIO), scale/shift[cout] f32 -> y[n,oh,ow,cout] f32
 * (rounded through the 16-bit storage type).  `same` != 0 requires stride 1.  If reps > 0 and
 * ms_out != NULL the layer is re-run `reps` times and the mean device time is returned. */
int icp_conv2d_bn_relu(int device, int precision, const float* x, int n, int h, int w, int cin,
                       const float* wgt, int kh, int kw, int cout, int stride, int same,
                       const float* scale, const float* shift, int relu, float* y, int reps,
                       float* ms_out);

/* Stateless pools through the tower's kernels (ops.py:367-370, :397-400).
 * kind 0: max 3x3/2 VALID; kind 1: avg kxk stride s, same != 0 => SAME (exclude-pad divisor). */
int icp_pool2d(int device, int precision, int kind, const float* x, int n, int h, int w, int c, int k,
               int stride, int same, float* y);

/* CRC-32C of a host buffer (needs no GPU): block checksums of TF V1 checkpoints (tf_checkpoint.py, the restore of
 * PyClassification.py:166-172). */
uint32_t icp_crc32c(const void* data, size_t n);

const char* icp_version(void);

#ifdef __cplusplus
}
#endif
#endif /* INCEPTION_B200_H_ */
