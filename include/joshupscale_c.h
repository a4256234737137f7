/* C-ABI of libJoshUpscale (B200 build): the drop-in boundary.
 *
 * Plain C, plain pointers and sizes, no C++/torch types.  Two groups:
 *
 *  (1) runtime entry points - what an FFI for the reference's public API
 *      would bind (core/public/JoshUpscale/core.h:64-94):
 *        ju_create         <- createRuntime(deviceId, modelPath)        core.h:91-92
 *        ju_destroy        <- Runtime::~Runtime                          core.h:65-66
 *        ju_process        <- Runtime::processImage(in, out)             core.h:68-69
 *        ju_get_info       <- getInputWidth/Height, getOutputWidth/Height core.h:71-82
 *        ju_last_error     <- getExceptionString()                       core.h:94
 *        ju_set_log_sink   <- setLogSink(LogSink*)                       core.h:28
 *      plus what the reference does NOT have but the north star needs:
 *        ju_process_batch  (N independent streams advanced in lockstep; with n < batch images
 *                           the remaining streams still advance, on stale staging data, and
 *                           their output is dropped: keep feeding the same leading streams)
 *        ju_reset_state / ju_read_tensor / ju_write_state (recurrent-state
 *        access for oracle comparison; the reference keeps the state private
 *        in TensorRTBackend::m_InterBuffers, tensorrt_backend.cc:213-218)
 *        ju_profile_ops    (per-kernel CUDA-event timing for roofline reports)
 *
 *  (2) kernel entry points ju_launch_* - each hand-written sm_100a kernel on
 *      raw device pointers, so the parity tests can drive one kernel against
 *      the oracle.  They replace what runs inside the reference's opaque
 *      TensorRT engine (IExecutionContext::enqueueV3, tensorrt_backend.cc:258)
 *      and its castKernel (core/src/cuda_convert.cc.cu:95-108).
 *
 * All functions return 0 on success, non-zero on failure; the message is
 * available from ju_last_error() on the calling thread.  There is no CPU
 * fallback: every compute entry fails if no CUDA device is present.
 */
#ifndef JOSHUPSCALE_C_H_
#define JOSHUPSCALE_C_H_

#include <stddef.h>
#include <stdint.h>

#if defined(_WIN32)
#  define JU_API __declspec(dllexport)
#else
#  define JU_API __attribute__((visibility("default")))
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ju_runtime ju_runtime;

/* DataLocation, core.h:30 */
enum { JU_LOC_CPU = 0, JU_LOC_CUDA = 1, JU_LOC_GRAPHICS_RESOURCE = 2 };

/* Image, core.h:32-38: BGRX u8, stride in bytes (may be negative/padded). */
typedef struct ju_image {
	void *ptr;
	int32_t location;
	int64_t stride;
	uint64_t width;
	uint64_t height;
} ju_image;

typedef struct ju_info {
	uint32_t input_width, input_height, output_width, output_height;
	uint32_t padded_width, padded_height;
	uint32_t batch;            /* streams advanced per call */
	uint32_t flow_num_inputs;  /* K */
	uint32_t gen_filters, gen_blocks;
	uint32_t flow_arch;        /* 0 autoencoder, 1 resnet */
	uint32_t conv_impl;        /* 0 = SIMT reference kernels, 1 = tcgen05 */
	uint32_t kernels_per_frame;
	uint32_t device;
	double gflop_per_frame;    /* algorithmic, true channel counts */
} ju_info;

typedef struct ju_tensor_desc {
	uint32_t dtype;   /* 0 f32, 1 f16, 2 u8 */
	uint32_t ndim;
	uint64_t dims[4];
	uint64_t bytes;
} ju_tensor_desc;

typedef struct ju_op_time {
	char name[64];
	double usec;       /* mean device time per launch */
	double flops;      /* algorithmic FLOPs per launch (0 if bandwidth-bound) */
	double bytes;      /* algorithmic bytes per launch */
	int32_t tensor_bound;
	int32_t reserved;  /* network layers covered (1 per kernel; the persistent trunk and "group:" entries cover many) */
} ju_op_time;

typedef void (*ju_log_fn)(const char *tag, int level, const char *message, void *user);

/* ---- (1) runtime ---------------------------------------------------- */
JU_API int ju_create(const char *model_path, int device, int batch, ju_runtime **out);
JU_API void ju_destroy(ju_runtime *rt);
JU_API int ju_process(ju_runtime *rt, const ju_image *input, const ju_image *output);
JU_API int ju_process_batch(ju_runtime *rt, int n, const ju_image *inputs, const ju_image *outputs);
JU_API int ju_get_info(const ju_runtime *rt, ju_info *info);
JU_API const char *ju_last_error(void);
JU_API void ju_set_log_sink(ju_log_fn fn, void *user);
JU_API int ju_reset_state(ju_runtime *rt);
/* Test hook for the failure path: the next frame's tcgen05 kernel `kernel_id` (1 = persistent
 * trunk) stalls on purpose.  ju_process then fails with "frame aborted: ... wait N expired" once
 * the wait time-out (JU_WAIT_TIMEOUT_MS, default 4000) has passed, the recurrent state is NOT
 * advanced and the runtime - and every other runtime of the process - stays usable, like a
 * failed enqueue in the reference (tensorrt_backend.cc:266, 276-277). */
JU_API int ju_debug_inject_stall(ju_runtime *rt, int kernel_id);
JU_API int ju_read_tensor(ju_runtime *rt, const char *name, void *dst, uint64_t capacity, ju_tensor_desc *desc);
JU_API int ju_write_state(ju_runtime *rt, const char *name, const void *src, uint64_t bytes);
JU_API int ju_profile_ops(ju_runtime *rt, int iters, ju_op_time *ops, int capacity, int *count);
JU_API int ju_device_count(void);
/* current device of the calling thread, used by the ju_dev_*, ju_host_*, ju_timer_*
 * and ju_launch_* helpers (a ju_runtime always runs on its own device) */
JU_API int ju_set_device(int device);
JU_API const char *ju_version(void);

/* ---- (2) kernels (device pointers, stream may be NULL) -------------- */

/* u8 BGRX frame -> fp16 flow-net input with zero padding and the K-frame
 * history shift (PreprocessLayer keras_layers.py:208; ZeroPadding2D
 * models.py:780-789; last_frames models.py:823).
 * frames: [batch] tightly packed H*W*4 u8.  flow_prev/flow_next:
 * [batch, PH, PW, cstride] fp16, channels [0,3K): frame t, t-1, ... */
JU_API int ju_launch_preprocess(const uint8_t *frames, const void *flow_prev, void *flow_next,
    int batch, int h, int w, int ph, int pw, int k, int cstride, void *stream);

/* Generic KxK (K = 1 or 3) SAME convolution, NHWC fp16 in, fp32 accumulate,
 * fused bias / residual / ReLU / LeakyReLU, fp16 or fp32 out, optional
 * 2x2 pixel-shuffle store (ConvTranspose k2s2; shuffle2 = 1) or fused
 * MaxPool2D(2) (shuffle2 = 2, tcgen05 only; out is [batch,h/2,w/2,cout_stride]).
 * impl: 0 SIMT, 1 tcgen05.
 * weights: packed by ju_pack_conv_weights for that impl. */
JU_API int ju_launch_conv(int impl, const void *in, const void *weights, const float *bias,
    const void *residual, void *out, int batch, int h, int w, int cin_stride, int cin,
    int cout, int cout_stride, int ksize, int act, float slope, int out_f32, int shuffle2,
    void *stream);

/* Pack Keras-layout fp32 weights (kh,kw,Cin,Cout) scaled per output channel
 * into the kernel-native fp16 layout.  Returns bytes needed when dst == NULL.
 * `dst` is HOST memory. */
JU_API int64_t ju_pack_conv_weights(int impl, const float *kernel, const float *scale,
    int ksize, int cin, int cin_padded, int cout, void *dst);

/* Global integer options: "tc_variant" (tcgen05 conv halo layout, see
 * conv_tc.cu: 0 = 18x10-pixel halo box, 1 = 18x16-pixel cross-check layout),
 * "tc_tma_epilogue" (0/1: shared-memory epilogue with TMA residual load and
 * TMA store), "tc_pdl" (0/1: programmatic dependent launch), "tc_dual" (0/1: two alternating
 * producer / issuer pipelines in conv_tc where the layer allows it). */
JU_API int ju_set_option(const char *key, int value);

/* Stand-alone timing of one convolution shape: allocates its own buffers,
 * launches `iters` times back to back and returns the mean device time. */
JU_API int ju_bench_conv(int impl, int batch, int h, int w, int cin, int cout, int ksize,
    int with_residual, int iters, double *usec);

JU_API int ju_launch_maxpool2(const void *in, void *out, int batch, int h, int w, int c, void *stream);
JU_API int ju_launch_upscale2(const void *in, void *out, int batch, int h, int w, int c, void *stream);

/* dense_image_warp (tfa/dense_image_warp.py:87-245) of the previous HR output
 * fused with SpaceToDepth(4) + Concatenate (models.py:523-530): writes the
 * generator input [batch,H,W,cstride] fp16: ch 0-2 current LR frame,
 * 3+(i*4+j)*3+c warped HR, rest zero.
 * pre_gen: [batch,4H,4W,4] fp16.  flow_head: [batch,PH,PW,32] fp32 BEFORE
 * depth_to_space (channel (i*4+j)*2+c, c=0 dy, c=1 dx).  frames: u8 BGRX.
 * taps (optional, may be NULL): [batch,4H,4W,4] fp32 (fy, fx, ay, ax) for the
 * "warp/indexing exact" check. */
JU_API int ju_launch_warp_s2d(const void *pre_gen, const float *flow_head, const uint8_t *frames,
    void *gen_in, float *taps, int batch, int h, int w, int ph, int pw, int cstride, void *stream);

/* conv_trans_2 + tanh + legacy-bilinear x4 of the input + add + clip +
 * uint8 pack + fp16 state write (models.py:573-593; keras_layers.py:227-230;
 * cuda_convert.cc.cu:39-45).  mid: [batch,2H,2W,32] fp16.  w2: [4][3][32]
 * fp32 (i*2+j, o, c), bias[3].  out_bgrx: [batch,4H,4W,4] u8 tightly packed.
 * pre_gen_next: [batch,4H,4W,4] fp16.  out_raw (optional): fp32 [batch,4H,4W,3]. */
JU_API int ju_launch_final(const void *mid, const float *w2, const float *bias2, const uint8_t *frames,
    uint8_t *out_bgrx, void *pre_gen_next, float *out_raw, int batch, int h, int w, void *stream);

/* Fused generator tail (tcgen05): conv_trans_1 (+folded BN, activation) ->
 * conv_trans_2 + bias -> tanh -> + legacy-bilinear x4 of the input -> clip ->
 * BGRX u8 + fp16 state, without materialising the [2H,2W,32] intermediate
 * (models.py:559-593, 808-823).  trunk: [batch,H,W,64] fp16.  w1: conv_trans_1
 * expressed as a 1x1 conv (1,1,64,128) with channel (i*2+j)*32+o, packed by
 * ju_pack_conv_weights(impl=1); bias1[128].  Other arguments as ju_launch_final. */
JU_API int ju_launch_tail(const void *trunk, const void *w1, const float *bias1, const float *w2,
    const float *bias2, const uint8_t *frames, uint8_t *out_bgrx, void *pre_gen_next, float *out_raw,
    int batch, int h, int w, int act, float slope, void *stream);

/* raw device memory helpers for the ctypes-side tests (no torch needed) */
JU_API int ju_dev_alloc(void **ptr, uint64_t bytes);
JU_API int ju_dev_free(void *ptr);
JU_API int ju_dev_upload(void *dst, const void *src, uint64_t bytes);
JU_API int ju_dev_download(void *dst, const void *src, uint64_t bytes);
JU_API int ju_dev_memset(void *dst, int value, uint64_t bytes);
JU_API int ju_dev_sync(void);
/* pinned host memory (for end-to-end timing with asynchronous copies) */
JU_API int ju_host_alloc(void **ptr, uint64_t bytes);
JU_API int ju_host_free(void *ptr);
/* evict L2 by overwriting a scratch buffer larger than the 126 MB L2 */
JU_API int ju_l2_flush(void);
/* CUDA-event stopwatch on the default stream: begin records, end records,
 * synchronises and returns the elapsed device time in microseconds. */
JU_API int ju_timer_begin(void);
JU_API int ju_timer_end(double *usec);
/* Self-check of the kernels' u8 -> float conversion (PreprocessLayer, keras_layers.py:195-208):
 * fills two HOST arrays of 256 floats with x/255 - 0.5 as the kernels compute it (division-free)
 * and as an IEEE fp32 division computes it; they must be bit-identical. */
JU_API int ju_u8_conversion_table(float *fast256, float *ieee256);

#ifdef __cplusplus
}
#endif
#endif /* JOSHUPSCALE_C_H_ */
