// Persistent flow-net kernel: every layer of get_flow_autoencoder
// (scripts/training/models.py:334-481) - 3x3 Conv2D + BN + activation (+ MaxPool2D(2) fused into the
// epilogue), the legacy bilinear x2 UpscaleLayers between the up blocks and the 32-channel flow
// head - in ONE cooperative launch instead of 18.
//
// Why: at batch 1 the flow net is 14 small GEMMs (0.9 - 4.8 GFLOP); launched one by one each pays
// launch latency, barrier / TMEM set-up, a cold pipeline and a full drain: 141 us for 36 GFLOP
// (0.15 of the tensor peak) although the MMAs need ~40 us.
//
// How: each layer runs the same pipeline as conv_tc_kernel (conv_tc_body.cuh, PERSIST = true:
// identical arithmetic, hence bit-identical results).  The CTAs stay resident (cooperative launch,
// grid <= SM count), TMEM and the mbarriers are set up once, and consecutive layers are separated
// by a release / acquire counter instead of a kernel boundary: a CTA publishes "my part of layer l
// is stored" after its own CTA barrier and goes straight on to layer l+1 - requesting that layer's
// weights - while it waits for the other CTAs; only the first activation load of layer l+1 waits
// for the counter.  Shared memory is re-partitioned per layer (resident or streamed weights, 2-8
// halo stages); the mbarriers live for the whole kernel and their phases carry over.
//
// The x2 upsamples run as element-wise layers on the eight epilogue warps (same arithmetic as
// upscale2_kernel, resample.cu).  Large batches are processed in chunks of streams, all layers
// per chunk, so that a chunk's activations stay in L2.
//
// (A finer-grained version - one counter per 16-row tile row, so that layer l+1 starts under the
// still-running layer l - was measured SLOWER: 234 vs 141 us.  At these layer sizes an item is
// 0.3 - 1.5 us of MMA work, and a GPU-scope release per item on the store side plus an L2 round
// trip per item on the load side cost more than the overlap gains.)
#include <cstring>
#include <vector>

#include "conv_tc_body.cuh"
#include "kernels.h"
#include "tc_common.cuh"

namespace ju {

namespace {

using namespace tc;
using namespace tcconv;

constexpr uint32_t kFlowSmem = kSmemLimit - 1024u;  // per-layer budget; the rest covers static shared memory
constexpr uint32_t kFlowTmemCols = 128;

constexpr int kMaxFlowLayers = 24;

struct alignas(128) FlowLayerDev {
	CUtensorMap map_a, map_b, map_c;
	TcParams p;
	int kind;  // 0 conv, 1 upsample x2
	int epi;   // conv: shared-memory epilogue mode (1 fp16 N=64, 2 fp16 N=32, 3 fp32 N=32)
	int items_per_stream;
	const __half *up_src;
	__half *up_dst;
	int up_h, up_w, up_c8;
};

// The layer table travels as ONE __grid_constant__ kernel parameter (CUDA >= 12.1 allows 32 KB of
// parameters): the TMA descriptors then live in parameter space exactly like those of the
// per-layer kernels.  (Descriptors in global memory would need a tensormap-proxy acquire fence at
// system scope in every CTA before their first use.)
struct FlowTable {
	FlowLayerDev layers[kMaxFlowLayers];
};
static_assert(sizeof(FlowTable) <= 32000, "layer table exceeds the kernel parameter space");

// legacy bilinear x2 (keras_layers.py:46-52): input row y of stream b -> output rows 2y, 2y+1.
// One thread = 8 channels of one INPUT pixel -> the 2x2 outputs it anchors; operation order and
// rounding are those of upscale2_kernel (resample.cu), so both paths produce the same bytes.
__device__ __forceinline__ void upsample_layer(const FlowLayerDev &L, const PersistLayer &pl, Waiter &W, int b0, int nb) {
	const int tid = static_cast<int>(threadIdx.x) - 64;
	if (tid < 0 || tid >= 256) return;  // the eight epilogue warps
	const int h = L.up_h, w = L.up_w, c8 = L.up_c8;
	const int items = nb * h;
	if (static_cast<int>(blockIdx.x) >= items) return;
	const uint4 *src = reinterpret_cast<const uint4 *>(L.up_src);
	uint4 *dst = reinterpret_cast<uint4 *>(L.up_dst);
	if (tid == 0) persist_wait_layer(pl, W, 13);
	asm volatile("bar.sync 3, 256;" ::: "memory");
	for (int item = blockIdx.x; item < items; item += gridDim.x) {
		const int b = b0 + item / h, y = item % h;
		const int y1 = y + 1 < h ? y + 1 : h - 1;
		const size_t in0 = (static_cast<size_t>(b) * h + y) * w, in1 = (static_cast<size_t>(b) * h + y1) * w;
		const size_t ow = 2 * static_cast<size_t>(w);
		for (int idx = tid; idx < w * c8; idx += 256) {
			const int x = idx / c8, cv = idx - x * c8;
			const int x1 = x + 1 < w ? x + 1 : w - 1;
			const uint4 va = __ldcg(src + (in0 + x) * c8 + cv);
			const uint4 vb = __ldcg(src + (in0 + x1) * c8 + cv);
			const uint4 vc = __ldcg(src + (in1 + x) * c8 + cv);
			const uint4 vd = __ldcg(src + (in1 + x1) * c8 + cv);
			const __half *pa = reinterpret_cast<const __half *>(&va);
			const __half *pb = reinterpret_cast<const __half *>(&vb);
			const __half *pc = reinterpret_cast<const __half *>(&vc);
			const __half *pd = reinterpret_cast<const __half *>(&vd);
			__align__(16) __half r01[8], r10[8], r11[8];
#pragma unroll
			for (int e = 0; e < 8; ++e) {
				const float a = __half2float(pa[e]), bq = __half2float(pb[e]);
				const float c = __half2float(pc[e]), d = __half2float(pd[e]);
				const float top = __fadd_rn(a, __fmul_rn(__fsub_rn(bq, a), 0.5f));
				const float bot = __fadd_rn(c, __fmul_rn(__fsub_rn(d, c), 0.5f));
				r01[e] = __float2half_rn(top);
				r10[e] = __float2half_rn(__fadd_rn(a, __fmul_rn(__fsub_rn(c, a), 0.5f)));
				r11[e] = __float2half_rn(__fadd_rn(top, __fmul_rn(__fsub_rn(bot, top), 0.5f)));
			}
			const size_t o00 = ((static_cast<size_t>(b) * 2 * h + 2 * y) * ow + 2 * x) * c8 + cv;
			dst[o00] = va;
			dst[o00 + c8] = *reinterpret_cast<const uint4 *>(r01);
			dst[o00 + ow * c8] = *reinterpret_cast<const uint4 *>(r10);
			dst[o00 + ow * c8 + c8] = *reinterpret_cast<const uint4 *>(r11);
		}
	}
	__threadfence();  // ordered before the layer's release by the caller's CTA barrier
}

// One convolution layer.  Deliberately NOT inlined: each epilogue variant gets its own register
// allocation (inlining all three into the layer loop cost 168 registers plus spills and ran the
// same pipeline 40 % slower than conv_tc_kernel).  The layer's parameters are read from a
// shared-memory copy: register-indexed loads from the kernel-parameter table in the hot loops, or
// a by-value copy held in 44 registers, both cost more than the occasional LDS.
template <int EPI>
__device__ __noinline__ void flow_conv_layer(const FlowLayerDev *L, const TcParams &p, uint8_t *smem_raw, PersistLayer *pl_io,
    Waiter *w_io) {
	PersistLayer pl = *pl_io;
	Waiter W = *w_io;
	conv_tc_body<3, EPI, true>(&L->map_a, &L->map_b, &L->map_c, &L->map_c, p, smem_raw, pl, W);
	pl_io->stage_par = pl.stage_par;
	pl_io->acc_par = pl.acc_par;
	pl_io->w_par = pl.w_par;
	w_io->dead = W.dead;
}

__global__ void __launch_bounds__(kThreadsWide, 1)
flow_df_tc_kernel(const __grid_constant__ FlowTable table, int n_layers, int batch, int chunk, unsigned int *counters,
    unsigned int *sync, TcStatus *status) {
	extern __shared__ uint8_t smem_raw[];
	__shared__ uint32_t tmem_slot;
	// the pipeline's mbarriers: initialised once, phases carried from layer to layer (PersistLayer)
	__shared__ __align__(8) unsigned long long bars[kBarBytes / 8];
	__shared__ __align__(16) TcParams layer_params;  // the current layer's parameters (see flow_conv_layer)
	const int warp = threadIdx.x >> 5;
	if (threadIdx.x == 0) {
		// same block layout as conv_tc_body: full[8], empty[8], tfull[2], tempty[2], weights
		const uint32_t bar = smem_u32(&bars[0]);
		for (int s = 0; s < 2 * kMaxStages; ++s) mbar_init(bar + 8u * s, 1);
		for (int s = 0; s < 2; ++s) {
			mbar_init(bar + 8u * (2 * kMaxStages + s), 1);
			mbar_init(bar + 8u * (2 * kMaxStages + 2 + s), kPersistTemptyCount);
		}
		mbar_init(bar + 8u * (2 * kMaxStages + 4), 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	if (warp == 1) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
		             "r"(kFlowTmemCols)
		             : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	tcgen05_fence_before();
	__syncthreads();
	tcgen05_fence_after();
	const uint32_t tmem_base = tmem_slot;

	Waiter W(status, TC_KERNEL_FLOW);
	// launch epoch: identical for every CTA of this launch (advanced by the last CTA to finish)
	const unsigned int epoch = *reinterpret_cast<volatile unsigned int *>(sync + 1);
	// test hook: a launch whose status block names this kernel never publishes a row
	const int stall = status && *reinterpret_cast<volatile int *>(&status->inject) == TC_KERNEL_FLOW ? 1 : 0;

	PersistLayer pl;
	pl.stage_par = pl.acc_par = pl.w_par = 0u;
	pl.bar_base = smem_u32(&bars[0]);
	pl.tmem_base = tmem_base;
	pl.dep_want = (epoch + 1u) * gridDim.x;
	unsigned int *done = nullptr;  // completion counter of the layer this CTA has just finished
	int ci = 0;
	for (int b0 = 0; b0 < batch; b0 += chunk, ++ci) {
		const int nb = batch - b0 < chunk ? batch - b0 : chunk;
		for (int l = 0; l < n_layers; ++l) {
			const FlowLayerDev &L = table.layers[l];
			// Layer boundary inside the CTA: every role has drained (all accumulators read, so every
			// MMA and operand fetch is complete; the store leaders have waited for their bulk stores;
			// the upsample threads have fenced their stores): shared memory may be re-partitioned
			// and this CTA's part of the previous layer is published.
			tcgen05_fence_before();
			__syncthreads();
			tcgen05_fence_after();
			if (threadIdx.x == 0 && done && !W.dead && !stall) {
				asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(done) : "memory");
			}
			W.layer = l;
			unsigned int *mine = counters + ci * n_layers + l;
			pl.dep_counter = l > 0 ? mine - 1 : nullptr;  // layer 0 reads what an earlier kernel wrote
			done = l + 1 < n_layers ? mine : nullptr;     // nobody waits for the last layer
			pl.item_begin = b0 * L.items_per_stream;
			pl.item_end = (b0 + nb) * L.items_per_stream;
			if (L.kind == 0) {
				static_assert(sizeof(TcParams) % 4 == 0, "TcParams is copied word by word");
				if (threadIdx.x < sizeof(TcParams) / 4) {
					reinterpret_cast<uint32_t *>(&layer_params)[threadIdx.x] = reinterpret_cast<const uint32_t *>(&L.p)[threadIdx.x];
				}
				__syncthreads();
				switch (L.epi) {
				case 1: flow_conv_layer<1>(&L, layer_params, smem_raw, &pl, &W); break;
				case 2: flow_conv_layer<2>(&L, layer_params, smem_raw, &pl, &W); break;
				default: flow_conv_layer<3>(&L, layer_params, smem_raw, &pl, &W); break;
				}
			} else {
				upsample_layer(L, pl, W, b0, nb);
			}
		}
	}

	tcgen05_fence_before();
	__syncthreads();
	if (warp == 1) {
		tcgen05_fence_after();
		asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kFlowTmemCols)
		             : "memory");
	}
	if (threadIdx.x == 0) {
		__threadfence();
		// the last CTA to finish advances the epoch for the next launch
		const unsigned int old = atomicAdd(sync, 1u);
		if (old == gridDim.x - 1u) {
			atomicExch(sync, 0u);
			atomicAdd(sync + 1, 1u);
		}
	}
}

}  // namespace

int flow_df_max_layers() { return kMaxFlowLayers; }

size_t flow_df_counter_words(int n_layers, int batch) {
	return static_cast<size_t>(n_layers) * static_cast<size_t>(batch);  // one per (chunk of streams, layer)
}

cudaError_t flow_df_tc_prepare(const FlowLayerSpec *specs, int n_layers, const ConvTcOptions &opt, int batch, int chunk,
    unsigned int *counters, unsigned int *sync, int cooperative, FlowDfLaunch *out) {
	if (n_layers < 1 || n_layers > kMaxFlowLayers || batch < 1 || !counters || !sync) return cudaErrorInvalidValue;
	if (chunk < 1 || chunk > batch) chunk = batch;
	ConvTcOptions lopt = opt;
	lopt.smem_limit = kFlowSmem;
	lopt.pdl = 0;
	std::shared_ptr<FlowTable> holder(new FlowTable);
	std::memset(holder.get(), 0, sizeof(FlowTable));
	FlowLayerDev *table = holder->layers;
	int max_items = 1;
	uint32_t max_smem = 0;
	for (int l = 0; l < n_layers; ++l) {
		FlowLayerDev &d = table[l];
		std::memset(&d, 0, sizeof(d));
		const FlowLayerSpec &s = specs[l];
		d.kind = s.kind;
		if (s.kind == 0) {
			const ConvArgs &a = s.conv;
			if (a.ksize != 3 || a.shuffle2 || a.residual || a.batch != batch) return cudaErrorInvalidValue;
			ConvTcLaunch launch;
			cudaError_t e = conv_tc_prepare(a, lopt, &launch);
			if (e != cudaSuccess) return e;
			static_assert(sizeof(TcParams) <= sizeof(launch.params), "ConvTcLaunch::params too small");
			std::memcpy(&d.p, launch.params, sizeof(TcParams));
			if (d.p.tma_epi < 1 || d.p.tma_epi > 3) return cudaErrorInvalidValue;  // needs the TMA-store epilogue
			std::memcpy(&d.map_a, launch.map_a, 128);
			std::memcpy(&d.map_b, launch.map_b, 128);
			std::memcpy(&d.map_c, launch.map_c, 128);
			d.epi = d.p.tma_epi;
			d.items_per_stream = d.p.total_tiles / batch;
			if (launch.smem_bytes > max_smem) max_smem = launch.smem_bytes;
			if (chunk * d.items_per_stream > max_items) max_items = chunk * d.items_per_stream;
		} else {
			if (s.up_c % 8 || s.up_h < 1 || s.up_w < 1) return cudaErrorInvalidValue;
			d.up_src = s.up_src;
			d.up_dst = s.up_dst;
			d.up_h = s.up_h;
			d.up_w = s.up_w;
			d.up_c8 = s.up_c / 8;
			d.items_per_stream = s.up_h;
			if (chunk * s.up_h > max_items) max_items = chunk * s.up_h;
		}
	}
	cudaError_t e = cudaFuncSetAttribute(flow_df_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kFlowSmem));
	if (e != cudaSuccess) return e;
	int dev = 0, sms = 148;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	out->table = holder;
	out->n_layers = n_layers;
	out->batch = batch;
	out->chunk = chunk;
	out->grid = max_items < sms ? max_items : sms;
	out->smem_bytes = max_smem > 0 ? max_smem : 1024u;
	out->counters = counters;
	out->sync = sync;
	out->cooperative = cooperative;
	return cudaSuccess;
}

cudaError_t flow_df_tc_launch(const FlowDfLaunch &l, TcStatus *status, cudaStream_t s) {
	cudaLaunchConfig_t cfg{};
	cfg.gridDim = dim3(l.grid);
	cfg.blockDim = dim3(kThreadsWide);
	cfg.dynamicSmemBytes = l.smem_bytes;
	cfg.stream = s;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeCooperative;
	attr[0].val.cooperative = 1;
	cfg.attrs = attr;
	cfg.numAttrs = l.cooperative ? 1 : 0;
	int n_layers = l.n_layers, batch = l.batch, chunk = l.chunk;
	unsigned int *counters = l.counters, *sync = l.sync;
	void *args[7] = {const_cast<void *>(l.table.get()), &n_layers, &batch, &chunk, &counters, &sync, &status};
	return cudaLaunchKernelExC(&cfg, reinterpret_cast<const void *>(flow_df_tc_kernel), args);
}

}  // namespace ju
