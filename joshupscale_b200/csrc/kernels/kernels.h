// Internal launch interface of the sm_100a kernels.  Host code (engine.cc)
// and the C-ABI (api.cc) call these; each returns the CUDA launch status.
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <memory>

namespace ju {

// Per-stream frame addresses, read by the pixel-I/O kernels from DEVICE memory
// so that a captured CUDA graph can be replayed on new images: the host
// rewrites this small table before each launch instead of re-instantiating.
// Strides are signed byte strides (bottom-up images have negative strides,
// avisynth_plugin/src/main.cc:125-142).
struct FrameIO {
	const uint8_t *in;
	long long in_stride;
	uint8_t *out;
	long long out_stride;
};

enum Act : int { ACT_NONE = 0, ACT_RELU = 1, ACT_LRELU = 2 };

// Status block of the tcgen05 kernels (device memory, one per engine; see tc_common.cuh).  A
// pipeline wait that exceeds `timeout_ms` records (kernel id << 8) | wait code here and in the
// host-mapped word, the kernel drains without side effects and the engine throws after the
// frame's stream synchronize - a stall is a recoverable exception, not a trap.
struct TcStatus {
	int code;         // 0 = ok
	int timeout_ms;   // 0 = default (4 s)
	int inject;       // fault injection (tests): id of the kernel whose pipeline stalls on purpose
	int where;        // (layer << 16) | CTA of the wait that expired (persistent kernels)
	int *host_code;   // mapped pinned host copy of {code, where} (may be null)
	unsigned int pending;  // bit c set: some thread was blocked in wait c when the frame was aborted
	int pad;
};
enum TcKernelId : int { TC_KERNEL_TRUNK = 1, TC_KERNEL_CONV = 2, TC_KERNEL_TAIL = 3, TC_KERNEL_FLOW = 4 };
const char *tc_kernel_name(int id);

struct ConvArgs {
	const __half *in;        // [batch, h, w, cin_stride]
	const void *weights;     // packed for the chosen impl
	const float *bias;       // [cout] or nullptr
	const __half *residual;  // [batch, h, w, cout_stride] or nullptr
	void *out;               // fp16 or fp32; [batch, h, w, cout_stride] (shuffle2: [batch, 2h, 2w, cout_stride])
	int batch, h, w;
	int cin_stride;          // channel stride of `in` (elements)
	int cin;                 // channels reduced over (multiple of 16, <= cin_stride)
	int cin_live;            // 0, or the number of leading channels that can be non-zero (weights and
	                         // activations beyond are zero): the tcgen05 kernel skips all-zero K steps
	int cout;                // output channels computed (shuffle2: 4 * per-pixel channels)
	int cout_stride;         // channel stride of `out` / `residual`
	int ksize;               // 1 or 3
	int act;
	float slope;
	int out_f32;
	int shuffle2;            // ConvTranspose k2s2: channel q*(cout/4)+o -> pixel (2y+q/2, 2x+q%2), channel o
	int pool;                // fuse MaxPool2D(2) into the epilogue: out is [batch, h/2, w/2, cout_stride] (tcgen05 only)
};

// brightness (optional, [batch] fp32 or nullptr): subtracted from the flow-net input
cudaError_t launch_preprocess(const FrameIO *io, const __half *flow_prev, __half *flow_next,
    const float *brightness, int batch, int h, int w, int ph, int pw, int k, int cstride, cudaStream_t s);

// normalize_brightness (scripts/training/models.py:772-779; utils.py:151):
// out[b] = mean over (h, w, c) of cur * BGR_LUMA * 3, deterministic summation order
cudaError_t launch_brightness(const FrameIO *io, float *out, int batch, int h, int w, cudaStream_t s);

// u8 -> float conversion table: the division-free kernel formulation and the IEEE-division
// reference for all 256 byte values (device pointers to 256 floats each)
cudaError_t launch_u8_table(float *fast256, float *ieee256, cudaStream_t s);

cudaError_t launch_conv_simt(const ConvArgs &a, cudaStream_t s);
// bytes of the SIMT weight layout [tap][cin_padded][cout] fp16
size_t conv_simt_weight_bytes(int ksize, int cin_padded, int cout);
void conv_simt_pack_weights(const float *kernel, const float *scale, int ksize, int cin,
    int cin_padded, int cout, __half *dst);

// ---- tcgen05 implicit-GEMM convolution (conv_tc.cu) ----------------------
// A prepared launch: two TMA tensor maps (activations, weights) + parameters.
// Built once per layer at plan time; replayed by the CUDA graph.
struct ConvTcLaunch {
	alignas(64) unsigned char map_a[128];
	alignas(64) unsigned char map_b[128];
	alignas(64) unsigned char map_c[128];  // output (TMA store)
	alignas(64) unsigned char map_r[128];  // residual (TMA load)
	alignas(8) unsigned char params[192];
	int grid;
	unsigned int smem_bytes;
	int pdl;
};
bool conv_tc_supported(const ConvArgs &a);
// variant: 0 = 18x10-pixel halo box (default), 1 = 18x16-pixel halo box (cross-check layout);
// tma_epilogue: shared-memory epilogue with TMA residual load + TMA store; pdl: programmatic
// dependent launch; dual: two alternating producer / issuer pipelines where the layer allows.
struct ConvTcOptions {
	int variant, tma_epilogue, pdl, dual;
	unsigned int smem_limit;  // shared-memory budget of one layer; 0 = everything a CTA may have (227 KB)
};
// process-wide defaults (changed by ju_set_option); engines copy them at creation
ConvTcOptions &conv_tc_default_options();
// a.weights must point at DEVICE memory packed by conv_tc_pack_weights with cin_padded == a.cin.
cudaError_t conv_tc_prepare(const ConvArgs &a, const ConvTcOptions &opt, ConvTcLaunch *out);
cudaError_t conv_tc_launch(const ConvTcLaunch &l, TcStatus *status, cudaStream_t s);
size_t conv_tc_weight_bytes(int ksize, int cin_padded, int cout);
void conv_tc_pack_weights(const float *kernel, const float *scale, int ksize, int cin,
    int cin_padded, int cout, __half *dst);

// ---- fused generator tail (tail_tc.cu): conv_trans_1 GEMM + conv_trans_2 +
// tanh + bilinear x4 + add + clip + u8 pack + state write in one kernel ------
struct TailArgs {
	const __half *in;       // trunk output [batch, h, w, cin_stride] (64 live channels)
	int cin_stride;
	const void *weights1;   // conv_trans_1 as 1x1 conv to 128 ch, packed by conv_tc_pack_weights (device)
	// HOST copies (they become kernel parameters): conv_trans_1's folded BN bias [>= 32], conv_trans_2
	// [4][3][32] and its bias [3]
	const float *bias1_host;
	const float *w2_host;
	const float *bias2_host;
	const FrameIO *io;
	__half *pre_gen_next;   // [batch, 4h, 4w, 4]
	float *out_raw;         // optional
	const float *brightness;  // optional [batch]: subtracted from the recurrent state
	int batch, h, w;
	int act;
	float slope;
	int pdl;
	int tile_row_begin, tile_row_end;  // optional band of 16-row tile rows (batch 1 only); 0, 0 = everything
};
struct TailTcLaunch {
	alignas(64) unsigned char map_a[128];
	alignas(64) unsigned char map_b[128];
	alignas(8) unsigned char params[1856];
	int grid;
	unsigned int smem_bytes;
	int pdl;
};
cudaError_t tail_tc_prepare(const TailArgs &a, TailTcLaunch *out);
cudaError_t tail_tc_launch(const TailTcLaunch &l, TcStatus *status, cudaStream_t s);

// ---- persistent ResBlock trunk (trunk_df_tc.cu): all 3x3 64->64 layers in one launch
struct TrunkArgs {
	void *buffers[3];        // T0 (input of the first block), T1, T2: [batch, h, w, cstride] fp16
	int cstride;
	const void *weights;     // n_layers x conv_tc_pack_weights(3, 64, 64, 64), concatenated
	const float *bias;       // [n_layers][64]
	unsigned int *sync_counter;  // two zero-initialised device words per engine (counter, epoch)
	unsigned int *flags;         // dataflow version: zero-initialised [n_layers][waves] counters (>= n_layers * tiles words)
	int batch, h, w;
	int n_layers;            // 2 x ResBlocks (+ 1 if lead_in is set)
	int act;
	float slope;
	// dataflow version only: an extra plain 3x3 64->64 conv + act in front of the ResBlocks (the
	// generator's conv_1), reading this tensor [batch, h, w, cstride] and writing T0; its weights
	// and bias come first in `weights` / `bias`
	const void *lead_in;
	// launch with cudaLaunchAttributeCooperative: the driver then guarantees that all CTAs of the
	// grid are resident together, which the inter-CTA dependencies of this kernel need when other
	// work (another process under MPS, another stream) may occupy SMs
	int cooperative;
	// CTA pairs (clusters of two): one tcgen05.mma.cta_group::2 per two pixel tiles, each CTA holds
	// half of the weights
	int pair;
};
struct TrunkTcLaunch {
	alignas(64) unsigned char maps[8 * 128];
	alignas(8) unsigned char params[128];
	int grid;
	unsigned int smem_bytes;
	unsigned int *sync_counter;
	int cooperative;
	int pair;
};
// (trunk_df_tc.cu) per-wave completion counters between layers, no grid-wide barrier
cudaError_t trunk_df_tc_prepare(const TrunkArgs &a, TrunkTcLaunch *out);
cudaError_t trunk_df_tc_launch(const TrunkTcLaunch &l, TcStatus *status, cudaStream_t s);
// index (0 or 2) of the buffer holding the trunk output after n_layers
inline int trunk_output_buffer(int n_layers) { return ((n_layers / 2) & 1) ? 2 : 0; }

// ---- persistent flow net (flow_df_tc.cu): all layers of the flow autoencoder in one launch ----
struct FlowLayerSpec {
	int kind;                  // 0 = 3x3 convolution (ConvArgs as for conv_tc_prepare), 1 = legacy bilinear x2
	ConvArgs conv;
	const __half *up_src;      // kind 1: [batch, up_h, up_w, up_c] -> up_dst [batch, 2 up_h, 2 up_w, up_c]
	__half *up_dst;
	int up_h, up_w, up_c;
};
struct FlowDfLaunch {
	std::shared_ptr<const void> table;  // host copy of the layer table, passed by value as a kernel parameter
	int n_layers, batch, chunk;
	int grid;
	unsigned int smem_bytes;
	unsigned int *counters, *sync;
	int cooperative;
};
int flow_df_max_layers();
size_t flow_df_counter_words(int n_layers, int batch);    // zero-initialised row counters
// sync: two zeroed words (finished-CTA counter, launch epoch); chunk: streams per pass over all
// layers.  Fails (cudaErrorInvalidValue) when a layer cannot run with the TMA-store epilogue.
cudaError_t flow_df_tc_prepare(const FlowLayerSpec *specs, int n_layers, const ConvTcOptions &opt, int batch, int chunk,
    unsigned int *counters, unsigned int *sync, int cooperative, FlowDfLaunch *out);
cudaError_t flow_df_tc_launch(const FlowDfLaunch &l, TcStatus *status, cudaStream_t s);

cudaError_t launch_maxpool2(const __half *in, __half *out, int batch, int h, int w, int c, cudaStream_t s);
cudaError_t launch_upscale2(const __half *in, __half *out, int batch, int h, int w, int c, cudaStream_t s);

cudaError_t launch_warp_s2d(const __half *pre_gen, const float *flow_head, const FrameIO *io,
    __half *gen_in, float *taps, const float *brightness, int batch, int h, int w, int ph, int pw,
    int cstride, cudaStream_t s);

cudaError_t launch_final(const __half *mid, const float *w2, const float *bias2, const FrameIO *io,
    __half *pre_gen_next, float *out_raw, const float *brightness, int batch, int h, int w,
    cudaStream_t s);

// ---- output temporal filter (frame_filter.cu) ------------------------------
// Parameters of scripts/inference/onnx/frame_moving_avg.py:53-87.
struct FilterParams {
	float strength;   // -s
	float threshold;  // -t
	float gain;       // -g (0 = sign function)
	float c3;         // 1 - strength / 2, rounded once on the host like the script's constant
	int window;       // -w (0 = global mean)
	int norm_l2;      // -n: 0 = L1, 1 = L2
	int limit;        // -l
	int luma;         // --luma-normalize
};
void filter_geometry(const FilterParams &fp, int h, int w, int *cells_y, int *cells_x, int *pad_t, int *pad_l);
int filter_partials_per_stream(int h, int w);
// out_raw: generator output after Clip [batch, 4h, 4w, 4] fp16; gen_in: generator input [batch, h, w, 64]
// (warped previous output in channels 3..50); scratch: batch * (1 + max(partials, cells)) floats.
// Writes the filtered u8 BGRX image (io[b].out) and the recurrent state pre_gen_next.
cudaError_t launch_frame_filter(const __half *out_raw, const __half *gen_in, const FrameIO *io, __half *pre_gen_next,
    const float *brightness, float *scratch, const FilterParams &fp, int batch, int h, int w, cudaStream_t s);


}  // namespace ju
