// The sm_100a engine behind Runtime::processImage.
//
// Takes over every responsibility of the reference's TensorRTBackend
// (core/src/tensorrt_backend.cc:117-278): owning the device buffers, the
// double-buffered recurrent state (m_InterBuffers, 213-218), one CUDA stream,
// one captured CUDA graph per ping-pong parity (257-263), and the per-frame
// sequence convert-in -> graph launch -> convert-out -> synchronize -> flip
// (270-278).  The graph's nodes are this repo's hand-written kernels instead
// of TensorRT tactics.  Extension over the reference: `batch` independent
// streams advance in lockstep through one graph (batch folded into GEMM M).
#pragma once

#include <functional>
#include <map>
#include <memory>
#include <set>
#include <string>
#include <vector>

#include "../kernels/kernels.h"
#include "common.h"
#include "copy_pool.h"
#include "joshupscale_c.h"
#include "model.h"

namespace ju {

struct ConvLayer {
	std::string name;
	int ksize = 3;
	int cinReal = 0;  // channels in the Keras kernel
	int cin = 0;      // reduced-over channels (padded to 16)
	int cout = 0;
	bool shuffle2 = false;
	int act = ACT_NONE;
	float slope = 0.f;
	DeviceBuffer wSimt;
	DeviceBuffer wTc;
	DeviceBuffer bias;
	std::vector<float> biasHost;
};

struct Op {
	std::string name;
	std::function<cudaError_t(cudaStream_t)> run;
	double flops = 0;  // algorithmic (true channel counts)
	double bytes = 0;  // algorithmic bytes moved
	bool tensorBound = false;
	int layers = 1;  // network layers covered by this launch (the persistent trunk covers many)
	int kernels = 1;  // kernel launches issued by run()
};

struct NamedTensor {
	void *ptr[2] = {nullptr, nullptr};  // indexed by parity when `pingPong`
	bool pingPong = false;
	int dtype = 1;  // 0 f32, 1 f16, 2 u8
	std::vector<std::uint64_t> dims;
	std::size_t bytes = 0;
	bool writable = false;
};

// packed weights / bias and synchronisation words of one persistent ResBlock stack
struct TrunkState {
	DeviceBuffer weights, bias, counter, flags;
};

class Engine {
public:
	Engine(const ModelFile &model, int device, int batch);
	~Engine();
	Engine(const Engine &) = delete;
	Engine &operator=(const Engine &) = delete;

	const ModelSpec &spec() const { return m_Spec; }
	int batch() const { return m_Batch; }
	int device() const { return m_Device; }
	int convImpl() const { return m_ConvImpl; }
	std::size_t kernelsPerFrame() const {
		std::size_t n = 0;
		for (const Op &op : m_Plans[0][0]) n += static_cast<std::size_t>(op.kernels);
		return n;
	}

	// n <= batch images.  All `batch` streams advance in lockstep through one graph: streams
	// beyond n are advanced on whatever their staging buffer holds and their output is dropped,
	// so callers that feed fewer images keep using the same leading streams.
	void process(int n, const ju_image *inputs, const ju_image *outputs);
	// test hook: the next frame's kernel `kernelId` (TcKernelId) stalls on purpose; process() then
	// throws KernelStallException and the runtime stays usable
	void injectStall(int kernelId);
	void resetState();
	void readTensor(const std::string &name, void *dst, std::uint64_t capacity, ju_tensor_desc *desc);
	void writeState(const std::string &name, const void *src, std::uint64_t bytes);
	std::vector<ju_op_time> profileOps(int iters);

private:
	void buildLayers(const ModelFile &model);
	void allocate();
	void buildPlan(int parity, int variant);
	void capture(int parity, int variant);
	void destroyGraphsAndEvents();
	ConvLayer *addConv(const std::string &name, const FoldedConv &f, int act, float slope, bool shuffle2);
	Op filterOp(const FrameIO *io, __half *preGenNext, const float *bright, int b0, int nb);
	__half *emitTrunk(std::vector<Op> &plan, TrunkState &ts, const std::string &prefix, int nBlocks, __half *t0,
	    __half *t1, __half *t2, int cstride, int H, int W, ConvLayer *lead, const __half *leadIn,
	    const std::function<void(const __half *, int, int, bool)> &afterChunk);
	void emitTail(std::vector<Op> &plan, int parity, const __half *trunkOut, int gs, int b0, int nb);
	bool flowNetFusable() const;
	ConvArgs tcConvArgs(ConvLayer *layer, const __half *in, int cinStride, void *out, int coutStride, int h, int w,
	    bool outF32, bool pool) const;
	void emitFlowNet(std::vector<Op> &plan, const __half *input, const std::function<__half *(std::size_t)> &activation);
	Op chunkDoneOp(int b0, int nb, int row0, int row1);
	Op convOp(ConvLayer *layer, const __half *in, int cinStride, const __half *residual, void *out,
	    int coutStride, int h, int w, bool outF32, bool pool = false);
	DeviceBuffer &newActivation(std::size_t bytes);
	void registerTensor(const std::string &name, void *p0, void *p1, int dtype,
	    std::vector<std::uint64_t> dims, std::size_t bytes, bool writable);
	void bindImages(int n, const ju_image *inputs, const ju_image *outputs);
	void unmapResources();
	static bool isPageable(const void *ptr);
	void ensureHostStaging();
	void uploadStatus(int inject);
	void recoverFromStall();

	ModelSpec m_Spec;
	int m_Device = 0;
	int m_Batch = 1;
	int m_SmCount = 1;
	int m_ConvImpl = 0;
	bool m_UseGraph = true;
	bool m_TrunkCooperative = false;
	ConvTcOptions m_TcOpt{};
	int m_Parity = 0;
	cudaStream_t m_Stream = nullptr;
	cudaGraphExec_t m_GraphExec[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};  // [variant][parity]

	PinnedBuffer m_IoHost;
	DeviceBuffer m_IoDev;
	std::vector<unsigned char> m_IoShadow;  // last address table sent to the device
	DeviceBuffer m_Status;      // TcStatus, read by every tcgen05 kernel's waits
	PinnedBuffer m_StatusHost;  // host-mapped copy of TcStatus::code, checked after every frame
	int m_WaitTimeoutMs = 0;
	DeviceBuffer m_Brightness;
	TrunkState m_GenTrunk, m_FlowTrunk;
	DeviceBuffer m_FlowCounters, m_FlowSync;  // persistent flow kernel: row counters, {done, epoch}
	int m_TcOps = 0;
	std::set<std::string> m_WarnedSimt;
	DeviceBuffer m_InStage, m_OutStage;
	std::vector<ju_image> m_LastOutputs;
	std::vector<bool> m_OutputNeedsCopy;
	std::vector<bool> m_OutputPooled;  // pageable output: device -> m_OutPinned -> copy pool -> caller
	// pageable host images (the plugins' case): engine-owned pinned staging + a multi-threaded memcpy
	struct BandCopy {
		CopyJob job;
		cudaEvent_t event;  // the band has arrived in pinned memory
	};
	int m_CopyThreads = 4;
	int m_CopySpinUs = 100;
	std::unique_ptr<HostCopyPool> m_Pool;
	PinnedBuffer m_InPinned, m_OutPinned;
	std::vector<BandCopy> m_BandCopies;     // this frame's bands, in copy-stream order
	std::vector<cudaEvent_t> m_BandEvents;  // reused from frame to frame
	std::size_t m_BandEventsUsed = 0;
	std::vector<cudaArray_t> m_OutputArrays;                 // per stream: mapped output array or null
	std::vector<cudaGraphicsResource_t> m_MappedResources;  // mapped for the frame in flight

	DeviceBuffer m_FlowIn[2], m_PreGen[2];
	DeviceBuffer m_FlowHead, m_GenIn, m_Trunk[3], m_Mid, m_W2, m_B2;
	std::vector<float> m_W2Host, m_B2Host;
	bool m_FilterOn = false;
	FilterParams m_Filter{};
	DeviceBuffer m_OutRaw, m_FilterScratch;
	std::size_t m_FilterScratchPerStream = 0;
	// output rows [row0, row1) of streams [b0, b0 + nb) are complete when `ev` fires
	struct Region {
		int b0, nb, row0, row1;
		cudaEvent_t ev;
	};
	// plan variant 0: the frame as one pass (device-resident images); variant 1 (batch 1, host
	// images): tail kernel in bands so that the device-to-host copy overlaps it
	std::vector<Region> m_Regions[2];
	int m_BuildVariant = 0;
	int m_TailBands = 1;
	bool m_HostVariant = false;
	cudaStream_t m_CopyStream = nullptr;
	std::vector<std::unique_ptr<DeviceBuffer>> m_Activations;
	std::vector<std::unique_ptr<ConvLayer>> m_Layers;
	std::map<std::string, ConvLayer *> m_LayerByName;
	std::vector<Op> m_Plans[2][2];  // [variant][parity]
	std::map<std::string, NamedTensor> m_Tensors;
	int m_FlowCStride = 64;
};

}  // namespace ju
