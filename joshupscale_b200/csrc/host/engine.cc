#include "engine.h"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>

namespace ju {

const char *tc_kernel_name(int id) {
	switch (id) {
	case TC_KERNEL_TRUNK: return "trunk_df_tc_kernel";
	case TC_KERNEL_CONV: return "conv_tc_kernel";
	case TC_KERNEL_TAIL: return "tail_tc_kernel";
	case TC_KERNEL_FLOW: return "flow_df_tc_kernel";
	default: return "tcgen05 kernel";
	}
}

namespace {

// thrown by convOp when MaxPool fusion was requested but the tensor-core
// epilogue cannot provide it for this layer; the caller plans a separate pool
struct PoolFusionUnavailable {};

int pad64(int c) { return (c + 63) / 64 * 64; }
int pad16(int c) { return (c + 15) / 16 * 16; }

int envInt(const char *name, int fallback) {
	const char *v = std::getenv(name);
	return v ? std::atoi(v) : fallback;
}

// Frames of all engines on one device are serialised inside the process.  The persistent trunk
// kernel has inter-CTA dependencies (trunk_df_tc.cu): all of its CTAs must be resident together,
// which holds when a frame has the device to itself (grid <= SM count, one CTA per SM) but not
// when two engines driven from two threads interleave their 227 KB-per-CTA kernels.  The trunks
// fill every SM anyway, so nothing is lost by taking turns.  Other processes sharing the GPU
// (MPS) are covered by the cooperative launch of the persistent kernels (JU_TRUNK_COOP, default 1).
std::mutex &deviceMutex(int device) {
	static std::mutex mutexes[64];
	return mutexes[device >= 0 && device < 64 ? device : 63];
}

}  // namespace

Engine::Engine(const ModelFile &model, int device, int batch)
    : m_Spec(model.spec()), m_Device(device), m_Batch(batch) {
	if (batch < 1 || batch > 64) throw std::invalid_argument("batch must be in [1, 64]");
	int count = 0;
	JU_CUDA(cudaGetDeviceCount(&count));
	if (device < 0 || device >= count) throw std::invalid_argument("invalid CUDA device index");
	DeviceGuard guard(device);
	cudaDeviceProp prop{};
	JU_CUDA(cudaGetDeviceProperties(&prop, device));
	if (prop.major != 10) {
		JU_LOG_WARN << "device " << device << " is sm_" << prop.major << prop.minor
		            << "; kernels are built for sm_100a only";
	}
	m_SmCount = prop.multiProcessorCount > 0 ? prop.multiProcessorCount : 1;
	m_ConvImpl = envInt("JU_CONV_IMPL", 1);  // 1 = tcgen05 (default), 0 = SIMT reference kernels
	// kernel options belong to this engine: process defaults (ju_set_option), then the environment
	m_TcOpt = conv_tc_default_options();
	m_TcOpt.variant = envInt("JU_TC_VARIANT", m_TcOpt.variant);
	m_TcOpt.tma_epilogue = envInt("JU_TC_TMA_EPILOGUE", m_TcOpt.tma_epilogue);
	m_TcOpt.pdl = envInt("JU_TC_PDL", m_TcOpt.pdl);
	m_TcOpt.dual = envInt("JU_TC_DUAL", m_TcOpt.dual);
	m_UseGraph = envInt("JU_NO_GRAPH", 0) == 0;
	// cooperative launch of the persistent kernels (co-residency guaranteed by the driver, also
	// against other processes); measured free: 171.8 vs 171.9 us for the psp_fast trunk
	m_TrunkCooperative = envInt("JU_TRUNK_COOP", 1) != 0;
	m_WaitTimeoutMs = envInt("JU_WAIT_TIMEOUT_MS", 0);
	// Copy pool for pageable caller images: 0 = let the driver stage them.  Default: up to 4 threads,
	// but no more than this process' share of the host cores when there is one process per GPU
	// (8 ranks x 4 spinning threads on a 16-core host collapsed to half the pinned throughput).
	{
		int devices = 1;
		if (cudaGetDeviceCount(&devices) != cudaSuccess || devices < 1) devices = 1;
		const int cores = static_cast<int>(std::max(1u, std::thread::hardware_concurrency()));
		const int share = std::max(1, cores / devices);
		m_CopyThreads = envInt("JU_COPY_THREADS", std::min(4, share));
	}
	// how long the calling thread polls for the next output band before it parks the pool and blocks
	// on the event: a batch-1 frame (0.4 - 0.7 ms) is polled through, a many-stream frame (10 ms)
	// leaves the cores to other processes between its tail groups
	m_CopySpinUs = envInt("JU_COPY_SPIN_US", m_Batch == 1 ? 2000 : 100);
	JU_CUDA(cudaStreamCreateWithFlags(&m_Stream, cudaStreamNonBlocking));
	JU_CUDA(cudaStreamCreateWithFlags(&m_CopyStream, cudaStreamNonBlocking));
	try {
		buildLayers(model);
		if (model.has("meta/frame_moving_avg")) {
			// output temporal filter baked into the model file, like the reference bakes it into the
			// exported graph (scripts/inference/onnx/frame_moving_avg.py); float32[8] =
			// {enabled, strength, window, threshold, gain, norm (0 L1 / 1 L2), limit, luma_normalize}
			const HostTensor &t = model.tensor("meta/frame_moving_avg");
			if (t.data.size() < 8) throw ModelException("meta/frame_moving_avg must hold 8 values");
			if (t.data[0] != 0.f) {
				m_FilterOn = true;
				m_Filter.strength = t.data[1];
				m_Filter.window = static_cast<int>(t.data[2]);
				m_Filter.threshold = t.data[3];
				m_Filter.gain = t.data[4];
				m_Filter.norm_l2 = t.data[5] != 0.f;
				m_Filter.limit = t.data[6] != 0.f;
				m_Filter.luma = t.data[7] != 0.f;
				m_Filter.c3 = static_cast<float>(1.0 - static_cast<double>(t.data[1]) / 2.0);
				if (m_Filter.window < 0) throw ModelException("frame_moving_avg window must be >= 0");
			}
		}
		allocate();
		buildPlan(0, 0);
		buildPlan(1, 0);
		// Second plan for frames whose images live in host memory: at batch 1 the tail kernel runs
		// as JU_TAIL_BANDS bands of tile rows, each followed by a completion event, so the 8 MB
		// device-to-host copy starts after the first band instead of after the whole frame.
		m_TailBands = envInt("JU_TAIL_BANDS", 3);
		m_HostVariant = m_TailBands > 1 && m_Batch == 1 && !m_FilterOn;
		if (m_HostVariant) {
			buildPlan(0, 1);
			buildPlan(1, 1);
			m_HostVariant = m_Regions[1].size() > 1;  // the plan may not have been able to band the tail
		}
		if (m_UseGraph) {
			for (int v = 0; v < (m_HostVariant ? 2 : 1); ++v) {
				capture(0, v);
				capture(1, v);
			}
		}
		JU_CUDA(cudaStreamSynchronize(m_Stream));
	} catch (...) {
		destroyGraphsAndEvents();
		cudaStreamDestroy(m_CopyStream);
		cudaStreamDestroy(m_Stream);
		throw;
	}
	JU_LOG_INFO << "engine ready: " << m_Spec.frameW << "x" << m_Spec.frameH << " -> "
	            << 4 * m_Spec.frameW << "x" << 4 * m_Spec.frameH << ", batch " << m_Batch << ", "
	            << m_Plans[0][0].size() << " kernels/frame, conv impl " << m_ConvImpl << " (" << m_TcOps / 2
	            << " tcgen05 launches/frame)";
}

Engine::~Engine() {
	int prev = -1;
	cudaGetDevice(&prev);
	cudaSetDevice(m_Device);
	struct Restore {  // the caller's current device survives the destructor (reference cuda.h:297-308)
		int dev;
		~Restore() {
			if (dev >= 0) cudaSetDevice(dev);
		}
	} restore{prev};
	cudaStreamSynchronize(m_Stream);
	cudaStreamSynchronize(m_CopyStream);
	for (cudaEvent_t ev : m_BandEvents) cudaEventDestroy(ev);
	destroyGraphsAndEvents();
	cudaStreamDestroy(m_CopyStream);
	cudaStreamDestroy(m_Stream);
}

void Engine::destroyGraphsAndEvents() {
	for (auto &pair : m_GraphExec)
		for (auto &g : pair)
			if (g) {
				cudaGraphExecDestroy(g);
				g = nullptr;
			}
	for (auto &regions : m_Regions) {
		for (Region &r : regions) cudaEventDestroy(r.ev);
		regions.clear();
	}
}

// ---------------------------------------------------------------------------
// layers
// ---------------------------------------------------------------------------

ConvLayer *Engine::addConv(const std::string &name, const FoldedConv &f, int act, float slope,
    bool shuffle2) {
	auto layer = std::make_unique<ConvLayer>();
	layer->name = name;
	layer->ksize = f.ksize;
	layer->cinReal = f.cin;
	layer->cin = pad16(f.cin);
	layer->cout = f.cout;
	layer->shuffle2 = shuffle2;
	layer->act = act;
	layer->slope = slope;
	std::vector<__half> packed(conv_simt_weight_bytes(f.ksize, layer->cin, f.cout) / sizeof(__half));
	conv_simt_pack_weights(f.kernel.data(), f.scale.data(), f.ksize, f.cin, layer->cin, f.cout,
	    packed.data());
	layer->wSimt = DeviceBuffer(packed.size() * sizeof(__half));
	layer->wSimt.upload(packed.data(), packed.size() * sizeof(__half));
	{
		// tcgen05 layout: Cin padded to the 64-channel K block of the activation buffers
		const int cinTc = pad64(f.cin);
		ConvArgs shape{};
		shape.ksize = f.ksize;
		shape.cin_stride = cinTc;
		shape.cin = cinTc;
		shape.cout = f.cout;
		shape.cout_stride = 8;
		shape.shuffle2 = shuffle2 ? 1 : 0;
		if (conv_tc_supported(shape)) {
			std::vector<__half> tc(conv_tc_weight_bytes(f.ksize, cinTc, f.cout) / sizeof(__half));
			conv_tc_pack_weights(f.kernel.data(), f.scale.data(), f.ksize, f.cin, cinTc, f.cout, tc.data());
			layer->wTc = DeviceBuffer(tc.size() * sizeof(__half));
			layer->wTc.upload(tc.data(), tc.size() * sizeof(__half));
		}
	}
	layer->bias = DeviceBuffer(f.bias.size() * sizeof(float));
	layer->bias.upload(f.bias.data(), f.bias.size() * sizeof(float));
	layer->biasHost = f.bias;
	ConvLayer *raw = layer.get();
	m_LayerByName[name] = raw;
	m_Layers.push_back(std::move(layer));
	return raw;
}

void Engine::buildLayers(const ModelFile &model) {
	const ModelSpec &s = m_Spec;
	const int actF = s.actFlow == 0 ? ACT_RELU : ACT_LRELU;
	const int actG = s.actGen == 0 ? ACT_RELU : ACT_LRELU;
	if (3 * s.flowInputs > 64) throw ModelException("too many flow inputs");
	if (s.flowArch == 0) {
		int n = static_cast<int>(s.flowFilters.size()) / 2;
		for (int i = 0; i < 2 * n; ++i) {
			std::string p = "flow/block_" + std::to_string(i + 1);
			addConv(p + "/conv_1", model.foldConv(p + "/conv_1", p + "/bn_1"), actF, s.slopeFlow, false);
			addConv(p + "/conv_2", model.foldConv(p + "/conv_2", p + "/bn_2"), actF, s.slopeFlow, false);
		}
		if (s.flowFilters.size() % 2) {
			addConv("flow/conv_1", model.foldConv("flow/conv_1", "flow/bn_1"), actF, s.slopeFlow, false);
		}
	} else {
		addConv("flow/conv_1", model.foldConv("flow/conv_1", "flow/bn_1"), actF, s.slopeFlow, false);
		for (int i = 0; i < s.flowFilters[1]; ++i) {
			std::string p = "flow/block_" + std::to_string(i + 1);
			addConv(p + "/conv_1", model.foldConv(p + "/conv_1", p + "/bn_1"), actF, s.slopeFlow, false);
			addConv(p + "/conv_2", model.foldConv(p + "/conv_2", p + "/bn_2"), actF, s.slopeFlow, false);
		}
	}
	ConvLayer *head = addConv("flow/conv_2", model.foldConv("flow/conv_2", ""), ACT_NONE, 0.f, false);
	if (head->cout != 32) throw ModelException("flow head must have 32 channels");

	addConv("generator/conv_1", model.foldConv("generator/conv_1", "generator/bn_1"), actG, s.slopeGen, false);
	for (int i = 0; i < s.genBlocks; ++i) {
		std::string p = "generator/block_" + std::to_string(i + 1);
		addConv(p + "/conv_1", model.foldConv(p + "/conv_1", p + "/bn_1"), actG, s.slopeGen, false);
		addConv(p + "/conv_2", model.foldConv(p + "/conv_2", p + "/bn_2"), actG, s.slopeGen, false);
	}
	ConvLayer *ct1 = addConv("generator/conv_trans_1",
	    model.foldConvTranspose("generator/conv_trans_1", "generator/bn_2"), actG, s.slopeGen, true);
	if (ct1->cout != 128) throw ModelException("conv_trans_1 must have 32 filters");

	// conv_trans_2 (2,2,3,32) + bias -> [q][o][c] fp32 for the fused final kernel;
	// weights are rounded to fp16 precision (engine storage contract)
	const HostTensor &k2 = model.tensor("generator/conv_trans_2/kernel");
	if (k2.dims != std::vector<int>{2, 2, 3, 32}) throw ModelException("conv_trans_2 must be (2,2,3,32)");
	std::vector<float> w2(k2.data.size());
	for (std::size_t i = 0; i < w2.size(); ++i) w2[i] = __half2float(__float2half_rn(k2.data[i]));
	m_W2 = DeviceBuffer(w2.size() * sizeof(float));
	m_W2.upload(w2.data(), w2.size() * sizeof(float));
	const auto &b2 = model.tensor("generator/conv_trans_2/bias").data;
	if (b2.size() != 3) throw ModelException("conv_trans_2 bias must have 3 entries");
	m_B2 = DeviceBuffer(3 * sizeof(float));
	m_B2.upload(b2.data(), 3 * sizeof(float));
	m_W2Host = w2;  // the fused tail kernel takes both as kernel parameters
	m_B2Host = b2;
}

// ---------------------------------------------------------------------------
// buffers
// ---------------------------------------------------------------------------

void Engine::registerTensor(const std::string &name, void *p0, void *p1, int dtype,
    std::vector<std::uint64_t> dims, std::size_t bytes, bool writable) {
	NamedTensor t;
	t.ptr[0] = p0;
	t.ptr[1] = p1;
	t.pingPong = p1 != nullptr;
	t.dtype = dtype;
	t.dims = std::move(dims);
	t.bytes = bytes;
	t.writable = writable;
	m_Tensors[name] = std::move(t);
}

void Engine::allocate() {
	const ModelSpec &s = m_Spec;
	const std::uint64_t B = m_Batch, H = s.frameH, W = s.frameW, PH = s.padH, PW = s.padW;
	m_Status = DeviceBuffer(sizeof(TcStatus));
	m_StatusHost = PinnedBuffer(2 * sizeof(int));
	m_StatusHost.as<int>()[0] = m_StatusHost.as<int>()[1] = 0;
	uploadStatus(0);
	m_Brightness = DeviceBuffer(sizeof(float) * B);
	m_IoHost = PinnedBuffer(sizeof(FrameIO) * B);
	m_IoDev = DeviceBuffer(sizeof(FrameIO) * B);
	m_InStage = DeviceBuffer(B * H * W * 4);
	m_OutStage = DeviceBuffer(B * 16 * H * W * 4);
	m_FlowCStride = 64;
	for (int p = 0; p < 2; ++p) {
		m_FlowIn[p] = DeviceBuffer(B * PH * PW * m_FlowCStride * sizeof(__half));
		m_PreGen[p] = DeviceBuffer(B * 16 * H * W * 4 * sizeof(__half));
	}
	m_FlowHead = DeviceBuffer(B * PH * PW * 32 * sizeof(float));
	const int gstride = pad64(std::max(m_Spec.genFilters, 51));
	m_GenIn = DeviceBuffer(B * H * W * 64 * sizeof(__half));
	for (auto &t : m_Trunk) t = DeviceBuffer(B * H * W * gstride * sizeof(__half));
	m_Mid = DeviceBuffer(B * 4 * H * W * 32 * sizeof(__half));
	if (m_FilterOn) {
		m_OutRaw = DeviceBuffer(B * 16 * H * W * 4 * sizeof(__half));
		int cy, cx, pt, pl;
		filter_geometry(m_Filter, static_cast<int>(H), static_cast<int>(W), &cy, &cx, &pt, &pl);
		const std::size_t work = std::max<std::size_t>(
		    filter_partials_per_stream(static_cast<int>(H), static_cast<int>(W)), static_cast<std::size_t>(cy) * cx);
		m_FilterScratchPerStream = 1 + work;
		m_FilterScratch = DeviceBuffer(B * m_FilterScratchPerStream * sizeof(float));
	}

	registerTensor("flow_in", m_FlowIn[0].get(), m_FlowIn[1].get(), 1,
	    {B, PH, PW, static_cast<std::uint64_t>(m_FlowCStride)}, m_FlowIn[0].bytes(), true);
	registerTensor("pre_gen", m_PreGen[0].get(), m_PreGen[1].get(), 1, {B, 4 * H, 4 * W, 4},
	    m_PreGen[0].bytes(), true);
	registerTensor("flow_head", m_FlowHead.get(), nullptr, 0, {B, PH, PW, 32}, m_FlowHead.bytes(), false);
	registerTensor("gen_in", m_GenIn.get(), nullptr, 1, {B, H, W, 64}, m_GenIn.bytes(), false);
	registerTensor("mid", m_Mid.get(), nullptr, 1, {B, 2 * H, 2 * W, 32}, m_Mid.bytes(), false);

	// default frame table: staging buffers (used until the first bindImages)
	FrameIO *io = m_IoHost.as<FrameIO>();
	for (std::uint64_t b = 0; b < B; ++b) {
		io[b].in = m_InStage.as<std::uint8_t>() + b * H * W * 4;
		io[b].in_stride = static_cast<long long>(W * 4);
		io[b].out = m_OutStage.as<std::uint8_t>() + b * 16 * H * W * 4;
		io[b].out_stride = static_cast<long long>(4 * W * 4);
	}
	JU_CUDA(cudaMemcpy(m_IoDev.get(), io, sizeof(FrameIO) * B, cudaMemcpyHostToDevice));
}

DeviceBuffer &Engine::newActivation(std::size_t bytes) {
	m_Activations.push_back(std::make_unique<DeviceBuffer>(bytes));
	return *m_Activations.back();
}

// ---------------------------------------------------------------------------
// plan
// ---------------------------------------------------------------------------

Op Engine::convOp(ConvLayer *L, const __half *in, int cinStride, const __half *residual, void *out,
    int coutStride, int h, int w, bool outF32, bool pool) {
	ConvArgs a{};
	a.in = in;
	a.weights = L->wSimt.get();
	a.bias = L->bias.as<float>();
	a.residual = residual;
	a.out = out;
	a.batch = m_Batch;
	a.h = h;
	a.w = w;
	a.cin_stride = cinStride;
	a.cin = L->cin;
	a.cout = L->cout;
	a.cout_stride = coutStride;
	a.ksize = L->ksize;
	a.act = L->act;
	a.slope = L->slope;
	a.out_f32 = outF32 ? 1 : 0;
	a.shuffle2 = L->shuffle2 ? 1 : 0;
	a.pool = pool ? 1 : 0;
	if (a.cin > cinStride) throw ModelException("channel stride too small for " + L->name);
	Op op;
	op.name = L->name;
	op.tensorBound = true;
	op.flops = 2.0 * m_Batch * h * w * L->ksize * L->ksize * L->cinReal * L->cout;
	op.bytes = static_cast<double>(m_Batch) * h * w *
	           (L->cinReal * 2.0 + L->cout * (outF32 ? 4.0 : 2.0) + (residual ? L->cout * 2.0 : 0.0));
	if (m_ConvImpl == 1 && L->wTc.get()) {
		ConvArgs t = a;
		t.weights = L->wTc.get();
		t.cin = pad64(L->cinReal);
		t.cin_live = L->cinReal;
		if (t.cin <= cinStride && conv_tc_supported(t)) {
			ConvTcLaunch launch;
			cudaError_t prep = conv_tc_prepare(t, m_TcOpt, &launch);
			if (pool && prep != cudaSuccess) throw PoolFusionUnavailable();
			checkCuda(prep, "conv_tc_prepare");
			TcStatus *status = m_Status.as<TcStatus>();
			op.run = [launch, status](cudaStream_t s) { return conv_tc_launch(launch, status, s); };
			++m_TcOps;
			return op;
		}
	}
	if (pool) throw PoolFusionUnavailable();
	if (m_ConvImpl == 1 && m_BuildVariant == 0 && m_WarnedSimt.insert(L->name).second) {
		// never silent: this layer runs at a few percent of the tensor-core kernel's speed
		JU_LOG_WARN << L->name << " (" << L->ksize << "x" << L->ksize << ", " << L->cinReal << " -> " << L->cout
		            << " channels) has no tcgen05 kernel (needs Cout = 32 or a multiple of 64): running the "
		               "CUDA-core reference convolution";
	}
	op.run = [a](cudaStream_t s) { return launch_conv_simt(a, s); };
	return op;
}

void Engine::buildPlan(int parity, int variant) {
	m_BuildVariant = variant;
	const ModelSpec &s = m_Spec;
	const int B = m_Batch, H = s.frameH, W = s.frameW, PH = s.padH, PW = s.padW;
	std::vector<Op> &plan = m_Plans[variant][parity];
	plan.clear();
	std::size_t actCursor = 0;
	auto activation = [&](std::size_t bytes) -> __half * {
		if (parity == 0 && variant == 0) return newActivation(bytes).as<__half>();
		DeviceBuffer &buf = *m_Activations.at(actCursor++);
		if (buf.bytes() != bytes) throw std::logic_error("activation plan mismatch");
		return buf.as<__half>();
	};
	auto layer = [&](const std::string &name) -> ConvLayer * { return m_LayerByName.at(name); };
	const FrameIO *io = m_IoDev.as<FrameIO>();

	// parity p reads state set p and writes set p^1 (tensorrt_backend.cc:247-256)
	const __half *flowPrev = m_FlowIn[parity].as<__half>();
	__half *flowNext = m_FlowIn[parity ^ 1].as<__half>();
	const __half *preGenPrev = m_PreGen[parity].as<__half>();
	__half *preGenNext = m_PreGen[parity ^ 1].as<__half>();

	const float *bright = nullptr;
	if (s.normalizeBrightness) {
		// models.py:772-779: scalar per stream, subtracted from the flow input and the
		// recurrent state, added back to the warped frame
		float *bout = m_Brightness.as<float>();
		bright = bout;
		Op op;
		op.name = "brightness";
		op.bytes = static_cast<double>(B) * H * W * 4.0;
		op.run = [=](cudaStream_t st) { return launch_brightness(io, bout, B, H, W, st); };
		plan.push_back(std::move(op));
	}
	{
		Op op;
		op.name = "preprocess";
		const int k = s.flowInputs, cs = m_FlowCStride;
		op.bytes = static_cast<double>(B) * (H * W * 4.0 + PH * PW * (3.0 * (k - 1) * 2 + 3.0 * k * 2));
		op.run = [=](cudaStream_t st) {
			return launch_preprocess(io, flowPrev, flowNext, bright, B, H, W, PH, PW, k, cs, st);
		};
		plan.push_back(std::move(op));
	}

	// ---- flow net ------------------------------------------------------
	const __half *x = flowNext;
	int xs = m_FlowCStride, h = PH, w = PW;
	auto conv = [&](const std::string &name, const __half *residual = nullptr) {
		ConvLayer *L = layer(name);
		int os = pad64(L->cout);
		__half *out = activation(static_cast<std::size_t>(B) * h * w * os * sizeof(__half));
		plan.push_back(convOp(L, x, xs, residual, out, os, h, w, false));
		x = out;
		xs = os;
	};
	bool flowDone = false;
	if (s.flowArch == 0 && flowNetFusable()) {
		// the whole autoencoder incl. the head as one persistent dataflow kernel (flow_df_tc.cu)
		emitFlowNet(plan, flowNext, activation);
		flowDone = true;
	} else if (s.flowArch == 0) {
		int n = static_cast<int>(s.flowFilters.size()) / 2;
		for (int i = 0; i < 2 * n; ++i) {
			std::string p = "flow/block_" + std::to_string(i + 1);
			conv(p + "/conv_1");
			__half *pooled = nullptr;  // MaxPool output, allocated first so both paths share it
			if (i < n) {
				const int os = pad64(layer(p + "/conv_2")->cout);
				pooled = activation(static_cast<std::size_t>(B) * (h / 2) * (w / 2) * os * sizeof(__half));
			}
			if (i < n && m_ConvImpl == 1 && envInt("JU_FUSED_POOL", 1) != 0 && h % 2 == 0 && w % 2 == 0) {
				// conv_2 + BN + act + MaxPool2D(2) in one kernel (models.py:386-409)
				ConvLayer *L = layer(p + "/conv_2");
				const int os = pad64(L->cout);
				try {
					Op fused = convOp(L, x, xs, nullptr, pooled, os, h, w, false, true);
					fused.name = p + "/conv_2+max_pool";
					plan.push_back(std::move(fused));
					x = pooled;
					xs = os;
					h /= 2;
					w /= 2;
					continue;
				} catch (const PoolFusionUnavailable &) {
					// layer shape not covered by the tensor-core epilogue: separate kernels below
				}
			}
			conv(p + "/conv_2");
			const __half *src = x;
			const int c = xs, hh = h, ww = w;
			Op op;
			if (i < n) {
				__half *out = pooled;
				op.name = p + "/max_pool";
				op.bytes = static_cast<double>(B) * hh * ww * s.flowFilters[i] * 2.0 * 1.25;
				op.run = [=](cudaStream_t st) { return launch_maxpool2(src, out, B, hh, ww, c, st); };
				h /= 2;
				w /= 2;
				x = out;
			} else {
				__half *out = activation(static_cast<std::size_t>(B) * (h * 2) * (w * 2) * c * sizeof(__half));
				op.name = p + "/upscale";
				op.bytes = static_cast<double>(B) * hh * ww * s.flowFilters[i] * 2.0 * 5.0;
				op.run = [=](cudaStream_t st) { return launch_upscale2(src, out, B, hh, ww, c, st); };
				h *= 2;
				w *= 2;
				x = out;
			}
			plan.push_back(std::move(op));
		}
		if (s.flowFilters.size() % 2) conv("flow/conv_1");
	} else {
		const bool flowTrunk = m_ConvImpl == 1 && s.flowFilters[0] == 64 && s.flowFilters[1] > 0 &&
		                       envInt("JU_FUSED_TRUNK", 1) != 0;
		if (flowTrunk) {
			// get_flow_resnet (models.py:257-331) is conv_1 + a ResBlock stack of the generator's shape:
			// the same persistent trunk kernel runs it
			const std::size_t bytes = static_cast<std::size_t>(B) * h * w * 64 * sizeof(__half);
			__half *f0 = activation(bytes), *f1 = activation(bytes), *f2 = activation(bytes);
			plan.push_back(convOp(layer("flow/conv_1"), x, xs, nullptr, f0, 64, h, w, false));
			x = emitTrunk(plan, m_FlowTrunk, "flow", s.flowFilters[1], f0, f1, f2, 64, h, w, nullptr, nullptr, nullptr);
			xs = 64;
		} else {
			conv("flow/conv_1");
			for (int i = 0; i < s.flowFilters[1]; ++i) {
				std::string p = "flow/block_" + std::to_string(i + 1);
				const __half *shortcut = x;
				conv(p + "/conv_1");
				conv(p + "/conv_2", shortcut);
			}
		}
	}
	if (!flowDone) {
		if (h != PH || w != PW) throw ModelException("flow net does not return to input resolution");
		plan.push_back(convOp(layer("flow/conv_2"), x, xs, nullptr, m_FlowHead.get(), 32, PH, PW, true));
	}

	// ---- warp + space-to-depth + concat --------------------------------
	{
		Op op;
		op.name = "warp_s2d";
		__half *genIn = m_GenIn.as<__half>();
		const float *head = m_FlowHead.as<float>();
		// read state (3ch fp16) + read flow (2 x fp32 here) + write warped (3ch fp16), SURVEY 8(d)
		op.bytes = static_cast<double>(B) * 16.0 * H * W * (3 * 2 + 2 * 4 + 3 * 2);
		op.run = [=](cudaStream_t st) {
			return launch_warp_s2d(preGenPrev, head, io, genIn, nullptr, bright, B, H, W, PH, PW, 64, st);
		};
		plan.push_back(std::move(op));
	}

	// ---- generator -----------------------------------------------------
	const int gs = pad64(std::max(s.genFilters, 51));
	__half *t0 = m_Trunk[0].as<__half>(), *t1 = m_Trunk[1].as<__half>(), *t2 = m_Trunk[2].as<__half>();
	__half *cur = t0, *tmp = t1, *nxt = t2;
	const bool fusedTrunk = m_ConvImpl == 1 && s.genFilters == 64 && s.genBlocks > 0 && gs % 64 == 0 &&
	                        envInt("JU_FUSED_TRUNK", 1) != 0;
	// conv_1 runs as layer 0 of the dataflow trunk when it can (emitTrunk decides the same way)
	ConvLayer *gc1 = layer("generator/conv_1");
	const bool conv1InTrunk = fusedTrunk && envInt("JU_TRUNK_LEAD", 1) != 0 &&
	                          gc1->wTc.get() && gc1->cout == 64 && gc1->ksize == 3 && gs == 64;
	if (!conv1InTrunk) plan.push_back(convOp(gc1, m_GenIn.as<__half>(), 64, nullptr, t0, gs, H, W, false));
	if (fusedTrunk) {
		// all 2 x genBlocks ResBlock convolutions as persistent launches (trunk_df_tc.cu / trunk_tc.cu)
		ConvLayer *ct1c = layer("generator/conv_trans_1");
		const bool tailPerChunk = m_ConvImpl == 1 && ct1c->wTc.get() && gs % 64 == 0 && s.genFilters == 64 &&
		                          envInt("JU_FUSED_TAIL", 1) != 0;
		bool tailsEmitted = false;
		// the tail kernel wants >= 2 streams per launch (1020 tiles of one PSP stream are 6.9 waves on
		// 148 SMs): finished trunk launches are collected until that many streams are waiting
		const int tailGroup = std::max(1, envInt("JU_TAIL_GROUP", 2));
		// ... but the LAST group's device-to-host copy has nothing left to hide behind, so the group
		// boundaries are shifted to leave that many streams (default 1) for the final tail
		const int tailLast = std::max(1, std::min(tailGroup, envInt("JU_TAIL_LAST", 1)));
		int tailB0 = 0, tailNb = 0;
		cur = emitTrunk(plan, m_GenTrunk, "generator", s.genBlocks, t0, t1, t2, gs, H, W, conv1InTrunk ? gc1 : nullptr,
		    m_GenIn.as<__half>(), [&](const __half *out, int b0, int nb, bool wholeBatch) {
			    // With several sub-batches each one is finished right away (tail kernel, output filter)
			    // and its completion is published as an event, so that process() can copy these streams'
			    // images to the host while the trunk of the next sub-batch is still running.
			    if (wholeBatch || !tailPerChunk) return;
			    if (tailNb == 0) tailB0 = b0;
			    tailNb += nb;
			    const int remaining = B - (b0 + nb);
			    const bool boundary = remaining >= tailLast && (remaining - tailLast) % tailGroup == 0;
			    if (tailNb < tailGroup && remaining > 0 && !boundary) return;
			    emitTail(plan, parity, out, gs, tailB0, tailNb);
			    tailNb = 0;
			    tailsEmitted = true;
		    });
		if (tailsEmitted) {
			if (parity == 0) {
				registerTensor("trunk", cur, nullptr, 1,
				    {static_cast<std::uint64_t>(B), static_cast<std::uint64_t>(H), static_cast<std::uint64_t>(W),
				        static_cast<std::uint64_t>(gs)},
				    m_Trunk[0].bytes(), false);
			}
			return;
		}
	} else {
		for (int i = 0; i < s.genBlocks; ++i) {
			std::string p = "generator/block_" + std::to_string(i + 1);
			plan.push_back(convOp(layer(p + "/conv_1"), cur, gs, nullptr, tmp, gs, H, W, false));
			plan.push_back(convOp(layer(p + "/conv_2"), tmp, gs, cur, nxt, gs, H, W, false));
			std::swap(cur, nxt);
		}
	}
	if (parity == 0) {
		registerTensor("trunk", cur, nullptr, 1,
		    {static_cast<std::uint64_t>(B), static_cast<std::uint64_t>(H), static_cast<std::uint64_t>(W),
		        static_cast<std::uint64_t>(gs)},
		    m_Trunk[0].bytes(), false);
	}
	ConvLayer *ct1 = layer("generator/conv_trans_1");
	const bool fusedTail = m_ConvImpl == 1 && ct1->wTc.get() && gs % 64 == 0 && s.genFilters == 64 &&
	                       envInt("JU_FUSED_TAIL", 1) != 0;
	if (fusedTail) {
		emitTail(plan, parity, cur, gs, 0, B);
		return;
	}
	plan.push_back(convOp(ct1, cur, gs, nullptr, m_Mid.get(), 32, H, W, false));
	{
		Op op;
		op.name = "final";
		const __half *mid = m_Mid.as<__half>();
		const float *w2 = m_W2.as<float>(), *b2 = m_B2.as<float>();
		// read mid (32ch fp16 @2Hx2W) + LR input + write BGRX u8 + fp16 state (3ch), SURVEY 8(d)
		op.bytes = static_cast<double>(B) * (4.0 * H * W * 32 * 2 + H * W * 4.0 + 16.0 * H * W * (4 + 3 * 2));
		op.flops = 2.0 * B * 4.0 * H * W * 32 * 12;
		__half *stateOut = m_FilterOn ? m_OutRaw.as<__half>() : preGenNext;
		const float *stateBright = m_FilterOn ? nullptr : bright;
		op.run = [=](cudaStream_t st) {
			return launch_final(mid, w2, b2, io, stateOut, nullptr, stateBright, B, H, W, st);
		};
		plan.push_back(std::move(op));
	}
	if (m_FilterOn) plan.push_back(filterOp(io, preGenNext, bright, 0, B));
	plan.push_back(chunkDoneOp(0, B, 0, 4 * H));
}

// get_flow_autoencoder (models.py:334-481) can run as the persistent flow kernel when every
// convolution has a tcgen05 kernel with the TMA-store epilogue (Cout = 32 or a multiple of 64)
// and every MaxPool sees even sizes.
bool Engine::flowNetFusable() const {
	const ModelSpec &s = m_Spec;
	// Opt-in (JU_FUSED_FLOW=1): measured SLOWER than one launch per layer on B200 - 222 vs 141 us at
	// batch 1, 1536 vs 1126 us at batch 16 (profiles/r02_flow_df_probe.txt) - because a layer
	// boundary inside the kernel (bulk-store completion, GPU-scope release, acquire, cold pipeline)
	// costs ~6 us against ~5 us for a kernel boundary in the captured graph; see flow_df_tc.cu.
	if (m_ConvImpl != 1 || !m_TcOpt.tma_epilogue || envInt("JU_FUSED_FLOW", 0) == 0 || envInt("JU_FUSED_POOL", 1) == 0) {
		return false;
	}
	const int n = static_cast<int>(s.flowFilters.size()) / 2;
	int h = s.padH, w = s.padW;
	auto usable = [&](const std::string &name) {
		auto it = m_LayerByName.find(name);
		return it != m_LayerByName.end() && it->second->wTc.get() && it->second->ksize == 3 &&
		       (it->second->cout == 32 || it->second->cout % 64 == 0);
	};
	for (int i = 0; i < 2 * n; ++i) {
		const std::string p = "flow/block_" + std::to_string(i + 1);
		if (!usable(p + "/conv_1") || !usable(p + "/conv_2")) return false;
		if (i < n) {
			if ((h | w) & 1) return false;
			h /= 2;
			w /= 2;
		} else {
			h *= 2;
			w *= 2;
		}
	}
	if (s.flowFilters.size() % 2 && !usable("flow/conv_1")) return false;
	if (2 * n + n + 2 > flow_df_max_layers()) return false;
	return usable("flow/conv_2") && h == s.padH && w == s.padW;
}

// ConvArgs of layer L on the tcgen05 path (what convOp passes to conv_tc_prepare)
ConvArgs Engine::tcConvArgs(ConvLayer *L, const __half *in, int cinStride, void *out, int coutStride, int h, int w,
    bool outF32, bool pool) const {
	ConvArgs a{};
	a.in = in;
	a.weights = L->wTc.get();
	a.bias = L->bias.as<float>();
	a.residual = nullptr;
	a.out = out;
	a.batch = m_Batch;
	a.h = h;
	a.w = w;
	a.cin_stride = cinStride;
	a.cin = pad64(L->cinReal);
	a.cin_live = L->cinReal;
	a.cout = L->cout;
	a.cout_stride = coutStride;
	a.ksize = L->ksize;
	a.act = L->act;
	a.slope = L->slope;
	a.out_f32 = outF32 ? 1 : 0;
	a.shuffle2 = 0;
	a.pool = pool ? 1 : 0;
	if (a.cin > cinStride) throw ModelException("channel stride too small for " + L->name);
	return a;
}

void Engine::emitFlowNet(std::vector<Op> &plan, const __half *input, const std::function<__half *(std::size_t)> &activation) {
	const ModelSpec &s = m_Spec;
	const int B = m_Batch, PH = s.padH, PW = s.padW;
	const int n = static_cast<int>(s.flowFilters.size()) / 2;
	std::vector<FlowLayerSpec> specs;
	double flops = 0, bytes = 0;
	const __half *x = input;
	int xs = m_FlowCStride, h = PH, w = PW;
	auto addConv = [&](const std::string &name, void *out, int os, bool outF32, bool pool) {
		ConvLayer *L = m_LayerByName.at(name);
		FlowLayerSpec spec{};
		spec.kind = 0;
		spec.conv = tcConvArgs(L, x, xs, out, os, h, w, outF32, pool);
		specs.push_back(spec);
		flops += 2.0 * B * h * w * 9.0 * L->cinReal * L->cout;
		bytes += static_cast<double>(B) * h * w * (L->cinReal * 2.0 + L->cout * (outF32 ? 4.0 : 2.0) * (pool ? 0.25 : 1.0));
	};
	auto bytesOf = [&](int hh, int ww, int c) { return static_cast<std::size_t>(B) * hh * ww * c * sizeof(__half); };
	for (int i = 0; i < 2 * n; ++i) {
		const std::string p = "flow/block_" + std::to_string(i + 1);
		const int os1 = pad64(m_LayerByName.at(p + "/conv_1")->cout), os2 = pad64(m_LayerByName.at(p + "/conv_2")->cout);
		__half *o1 = activation(bytesOf(h, w, os1));
		addConv(p + "/conv_1", o1, os1, false, false);
		x = o1;
		xs = os1;
		if (i < n) {
			// conv_2 + BN + act + MaxPool2D(2) (models.py:386-409): pooled in the epilogue
			__half *pooled = activation(bytesOf(h / 2, w / 2, os2));
			addConv(p + "/conv_2", pooled, os2, false, true);
			x = pooled;
			h /= 2;
			w /= 2;
		} else {
			// conv_2, then UpscaleLayer(bilinear, 2) (models.py:441-446) as an element-wise layer
			__half *o2 = activation(bytesOf(h, w, os2));
			addConv(p + "/conv_2", o2, os2, false, false);
			__half *up = activation(bytesOf(2 * h, 2 * w, os2));
			FlowLayerSpec spec{};
			spec.kind = 1;
			spec.up_src = o2;
			spec.up_dst = up;
			spec.up_h = h;
			spec.up_w = w;
			spec.up_c = os2;
			specs.push_back(spec);
			bytes += static_cast<double>(B) * h * w * s.flowFilters[i] * 2.0 * 5.0;
			x = up;
			h *= 2;
			w *= 2;
		}
		xs = os2;
	}
	if (s.flowFilters.size() % 2) {
		const int os = pad64(m_LayerByName.at("flow/conv_1")->cout);
		__half *o = activation(bytesOf(h, w, os));
		addConv("flow/conv_1", o, os, false, false);
		x = o;
		xs = os;
	}
	addConv("flow/conv_2", m_FlowHead.get(), 32, true, false);

	const int nLayers = static_cast<int>(specs.size());
	if (!m_FlowCounters.get()) {
		m_FlowCounters = DeviceBuffer(flow_df_counter_words(nLayers, B) * sizeof(unsigned int));
		m_FlowSync = DeviceBuffer(2 * sizeof(unsigned int));
	}
	// JU_FLOW_SUBBATCH: streams per pass over all layers (-1: see below, 0: the whole batch).  A pass
	// costs ~3 us of pipeline fill / drain per layer and CTA, a whole-batch pass streams every
	// activation through HBM once the batch no longer fits the L2
	int chunk = envInt("JU_FLOW_SUBBATCH", -1);
	if (chunk < 0) chunk = B <= 4 ? B : 4;
	if (chunk == 0 || chunk > B) chunk = B;
	FlowDfLaunch launch{};
	// JU_FLOW_DEBUG_LAYERS=n: run only the first n layers (bisecting a stalled pipeline; the frame is wrong)
	const int debugLayers = envInt("JU_FLOW_DEBUG_LAYERS", 0);
	checkCuda(flow_df_tc_prepare(specs.data(), debugLayers > 0 && debugLayers < nLayers ? debugLayers : nLayers, m_TcOpt, B, chunk,
	              m_FlowCounters.as<unsigned int>(), m_FlowSync.as<unsigned int>(), m_TrunkCooperative ? 1 : 0, &launch),
	    "flow_df_tc_prepare");
	TcStatus *status = m_Status.as<TcStatus>();
	Op op;
	op.name = "flow/*(persistent)";
	op.tensorBound = true;
	op.layers = nLayers;
	op.flops = flops;
	op.bytes = bytes;
	op.run = [launch, status](cudaStream_t st) { return flow_df_tc_launch(launch, status, st); };
	plan.push_back(std::move(op));
	++m_TcOps;
}

// A stack of `nBlocks` ResBlocks (3x3 64->64 conv + BN + act, conv + BN + shortcut + act;
// scripts/training/models.py:193-254) named <prefix>/block_<i>/conv_<j>, input in t0, as persistent
// trunk launches.  Returns the buffer that holds the result; `afterChunk(result, b0, nb, whole)`
// is called after the launch of every sub-batch.
__half *Engine::emitTrunk(std::vector<Op> &plan, TrunkState &ts, const std::string &prefix, int nBlocks, __half *t0,
    __half *t1, __half *t2, int cstride, int H, int W,
    ConvLayer *lead, const __half *leadIn, const std::function<void(const __half *, int, int, bool)> &afterChunk) {
	const int B = m_Batch;
	const int nBlockLayers = 2 * nBlocks;
	// optional plain conv in front of the ResBlocks (the generator's conv_1): same 3x3 64->64 shape
	// once its 51 input channels are padded, so it becomes layer 0 of the dataflow trunk
	if (lead && !(envInt("JU_TRUNK_LEAD", 1) != 0 && lead->wTc.get() && lead->cout == 64 && lead->ksize == 3 &&
	                 lead->cinReal <= 64 && cstride == 64)) {
		lead = nullptr;
	}
	const int nLead = lead ? 1 : 0;
	const int nLayers = nBlockLayers + nLead;
	auto blockLayer = [&](int l) {
		if (lead && l == 0) return lead;
		l -= nLead;
		return m_LayerByName.at(prefix + "/block_" + std::to_string(l / 2 + 1) + "/conv_" + std::to_string(l % 2 + 1));
	};
	if (!ts.weights.get()) {
		const std::size_t per = conv_tc_weight_bytes(3, 64, 64);
		ts.weights = DeviceBuffer(per * nLayers);
		ts.bias = DeviceBuffer(sizeof(float) * 64 * nLayers);
		ts.counter = DeviceBuffer(sizeof(unsigned int) * 2 * static_cast<std::size_t>(B));
		ts.flags = DeviceBuffer(sizeof(unsigned int) * static_cast<std::size_t>(nLayers) * B * ((H + 15) / 16) * ((W + 7) / 8));
		for (int l = 0; l < nLayers; ++l) {
			ConvLayer *L = blockLayer(l);
			if (!L->wTc.get() || L->cout != 64 || (L != lead && L->cinReal != 64) || L->ksize != 3) {
				throw ModelException("unexpected ResBlock layer shape");
			}
			JU_CUDA(cudaMemcpy(ts.weights.as<char>() + per * l, L->wTc.get(), per, cudaMemcpyDeviceToDevice));
			JU_CUDA(cudaMemcpy(ts.bias.as<float>() + 64 * l, L->bias.get(), 64 * sizeof(float), cudaMemcpyDeviceToDevice));
		}
	}
	ConvLayer *first = blockLayer(nLead);
	if (lead && (lead->act != first->act || lead->slope != first->slope)) {
		throw ModelException("lead convolution and ResBlocks use different activations");
	}
	TrunkArgs ta{};
	ta.buffers[0] = t0;
	ta.buffers[1] = t1;
	ta.buffers[2] = t2;
	ta.cstride = cstride;
	ta.weights = ts.weights.get();
	ta.bias = ts.bias.as<float>();
	ta.sync_counter = ts.counter.as<unsigned int>();
	ta.flags = ts.flags.as<unsigned int>();
	ta.batch = B;
	ta.h = H;
	ta.w = W;
	ta.n_layers = nLayers;
	ta.act = first->act;
	ta.slope = first->slope;
	ta.lead_in = lead ? leadIn : nullptr;
	// JU_TRUNK_SUBBATCH: a large batch runs as consecutive launches of this
	// many streams, each through ALL layers, so that the three trunk tensors of one launch stay
	// resident in L2 (one launch over 16 PSP streams would stream ~800 MB per layer through HBM).
	// Default (-1): as many streams as fit 45 % of the L2, i.e. ONE PSP stream (50 MB): with two
	// (100 MB of the 126 MB) the reads still hit, but ~80 % of every layer's output is written back
	// to DRAM (ncu: 1.32 GB per launch, profiles/r02_ncu_trunk_df_b2_of_16.json) and, the 16-stream
	// step being power-capped, that costs 3 % (1594 vs 1644 fps, 1700 vs 1747 MHz).
	// 0 = one launch for the whole batch.
	const std::size_t perStream = static_cast<std::size_t>(H) * W * cstride;
	int chunk = envInt("JU_TRUNK_SUBBATCH", -1);
	if (chunk < 0) {
		int l2 = 0;
		cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, m_Device);
		const double budget = 0.45 * static_cast<double>(l2);
		chunk = static_cast<int>(budget / (3.0 * perStream * sizeof(__half)));
		if (chunk < 1) chunk = 1;
	}
	if (chunk <= 0 || chunk > B) chunk = B;
	const std::size_t tilesPerStream = static_cast<std::size_t>((H + 15) / 16) * ((W + 7) / 8);
	__half *result = trunk_output_buffer(nBlockLayers) == 0 ? t0 : t2;
	const double leadFlops = lead ? 2.0 * H * W * 9.0 * lead->cinReal * 64 : 0.0;
	TcStatus *status = m_Status.as<TcStatus>();
	ta.cooperative = m_TrunkCooperative ? 1 : 0;
	ta.pair = envInt("JU_TRUNK_PAIR", 0) != 0 ? 1 : 0;
	for (int b0 = 0; b0 < B; b0 += chunk) {
		TrunkArgs sub = ta;
		sub.batch = std::min(chunk, B - b0);
		for (int i = 0; i < 3; ++i) sub.buffers[i] = static_cast<__half *>(ta.buffers[i]) + perStream * b0;
		sub.sync_counter = ta.sync_counter + 2 * (b0 / chunk);
		sub.flags = ta.flags + static_cast<std::size_t>(nLayers) * tilesPerStream * b0;
		if (lead) sub.lead_in = leadIn + perStream * b0;
		TrunkTcLaunch launch;
		checkCuda(trunk_df_tc_prepare(sub, &launch), "trunk_df_tc_prepare");
		Op op;
		op.name = prefix + "/block_*(persistent)";
		op.tensorBound = true;
		op.layers = b0 == 0 ? nLayers : 0;  // network layers are counted once, not once per sub-batch
		op.flops = 2.0 * sub.batch * H * W * 9.0 * 64 * 64 * nBlockLayers + sub.batch * leadFlops;
		op.bytes = static_cast<double>(sub.batch) * H * W * 64 * 2.0 * (2.0 * nLayers + 0.5 * nLayers);
		op.run = [launch, status](cudaStream_t st) { return trunk_df_tc_launch(launch, status, st); };
		plan.push_back(std::move(op));
		++m_TcOps;
		if (afterChunk) afterChunk(result, b0, sub.batch, chunk >= B);
	}
	return result;
}

// Fused tail (+ output filter) for streams [b0, b0 + nb), followed by the event that tells
// process() these streams' images are complete.
void Engine::emitTail(std::vector<Op> &plan, int parity, const __half *trunkOut, int gs, int b0, int nb) {
	const int H = m_Spec.frameH, W = m_Spec.frameW;
	const FrameIO *io = m_IoDev.as<FrameIO>() + b0;
	const std::size_t hrStream = static_cast<std::size_t>(16) * H * W * 4;  // fp16 elements of one HR state
	__half *preGenNext = m_PreGen[parity ^ 1].as<__half>() + hrStream * b0;
	const float *bright = m_Spec.normalizeBrightness ? m_Brightness.as<float>() + b0 : nullptr;
	ConvLayer *ct1 = m_LayerByName.at("generator/conv_trans_1");
	// conv_trans_1 + conv_trans_2 + tanh + upscale + add + clip + pack + state in ONE kernel
	TailArgs ta{};
	ta.in = trunkOut + static_cast<std::size_t>(H) * W * gs * b0;
	ta.cin_stride = gs;
	ta.weights1 = ct1->wTc.get();
	ta.bias1_host = ct1->biasHost.data();
	ta.w2_host = m_W2Host.data();
	ta.bias2_host = m_B2Host.data();
	ta.io = io;
	ta.pre_gen_next = m_FilterOn ? m_OutRaw.as<__half>() + hrStream * b0 : preGenNext;
	ta.out_raw = nullptr;
	ta.brightness = m_FilterOn ? nullptr : bright;
	ta.batch = nb;
	ta.h = H;
	ta.w = W;
	ta.act = ct1->act;
	ta.slope = ct1->slope;
	ta.pdl = 1;
	TcStatus *status = m_Status.as<TcStatus>();
	const int tileRows = (H + 15) / 16;
	const int bands = (m_BuildVariant == 1 && nb == 1 && m_Batch == 1 && !m_FilterOn) ? std::min(m_TailBands, tileRows) : 1;
	for (int band = 0; band < bands; ++band) {
		const int r0 = tileRows * band / bands, r1 = tileRows * (band + 1) / bands;
		TailArgs tb = ta;
		if (bands > 1) {
			tb.tile_row_begin = r0;
			tb.tile_row_end = r1;
		}
		TailTcLaunch launch;
		checkCuda(tail_tc_prepare(tb, &launch), "tail_tc_prepare");
		const double share = static_cast<double>(r1 - r0) / tileRows;
		Op op;
		op.name = "tail_fused";
		op.tensorBound = false;
		// read trunk (64ch fp16) + LR input + write BGRX u8 + fp16 state (3ch), SURVEY 8(d)
		op.bytes = share * nb * (H * W * 64 * 2.0 + H * W * 4.0 + 16.0 * H * W * (4 + 3 * 2));
		op.flops = share * 2.0 * nb * H * W * (64.0 * 128 + 4.0 * 32 * 12);
		op.run = [launch, status](cudaStream_t st) { return tail_tc_launch(launch, status, st); };
		plan.push_back(std::move(op));
		++m_TcOps;
		if (bands > 1) plan.push_back(chunkDoneOp(b0, nb, 64 * r0, std::min(64 * r1, 4 * H)));
	}
	if (bands > 1) return;
	if (m_FilterOn) plan.push_back(filterOp(io, preGenNext, bright, b0, nb));
	plan.push_back(chunkDoneOp(b0, nb, 0, 4 * H));
}

// Marks streams [b0, b0 + nb) complete: an event that process() makes the copy stream wait on.
// Inside stream capture it becomes an external event-record node of the frame graph.
Op Engine::chunkDoneOp(int b0, int nb, int row0, int row1) {
	std::vector<Region> &regions = m_Regions[m_BuildVariant];
	std::size_t idx = 0;
	for (; idx < regions.size(); ++idx) {
		const Region &r = regions[idx];
		if (r.b0 == b0 && r.nb == nb && r.row0 == row0 && r.row1 == row1) break;
	}
	if (idx == regions.size()) {  // both parities share the events
		Region r{b0, nb, row0, row1, nullptr};
		JU_CUDA(cudaEventCreateWithFlags(&r.ev, cudaEventDisableTiming));
		regions.push_back(r);
	}
	cudaEvent_t ev = regions[idx].ev;
	Op op;
	op.name = "sync:streams_done";
	op.kernels = 0;
	op.layers = 0;
	op.run = [ev](cudaStream_t st) {
		cudaStreamCaptureStatus status = cudaStreamCaptureStatusNone;
		cudaError_t e = cudaStreamIsCapturing(st, &status);
		if (e != cudaSuccess) return e;
		return cudaEventRecordWithFlags(ev, st,
		    status == cudaStreamCaptureStatusActive ? cudaEventRecordExternal : cudaEventRecordDefault);
	};
	return op;
}

// frame_moving_avg.py:142-307: blend the generator output (left in m_OutRaw by the tail kernel,
// un-normalised) with the warped previous output still sitting in the generator input tensor;
// rewrites the u8 image and produces the recurrent state
// for streams [b0, b0 + nb); io / preGenNext / bright already point at stream b0
Op Engine::filterOp(const FrameIO *io, __half *preGenNext, const float *bright, int b0, int nb) {
	const int B = nb, H = m_Spec.frameH, W = m_Spec.frameW;
	Op op;
	op.name = "frame_moving_avg";
	op.kernels = m_Filter.window == 0 ? 3 : 2;
	// two passes over out (4 x fp16) and pw (fp16) + u8 image + fp16 state
	op.bytes = static_cast<double>(B) * (2.0 * (16.0 * H * W * 8 + H * W * 128.0) + 16.0 * H * W * (4 + 8));
	const __half *outRaw = m_OutRaw.as<__half>() + static_cast<std::size_t>(16) * H * W * 4 * b0;
	const __half *genIn = m_GenIn.as<__half>() + static_cast<std::size_t>(H) * W * 64 * b0;
	float *scratch = m_FilterScratch.as<float>() + m_FilterScratchPerStream * b0;
	const FilterParams fp = m_Filter;
	op.run = [=](cudaStream_t st) {
		return launch_frame_filter(outRaw, genIn, io, preGenNext, bright, scratch, fp, B, H, W, st);
	};
	return op;
}

void Engine::capture(int parity, int variant) {
	cudaGraph_t graph = nullptr;
	JU_CUDA(cudaStreamBeginCapture(m_Stream, cudaStreamCaptureModeThreadLocal));
	cudaError_t err = cudaSuccess;
	for (const Op &op : m_Plans[variant][parity]) {
		err = op.run(m_Stream);
		if (err != cudaSuccess) break;
	}
	cudaError_t endErr = cudaStreamEndCapture(m_Stream, &graph);
	checkCuda(err, "kernel launch during graph capture");
	checkCuda(endErr, "cudaStreamEndCapture");
	cudaError_t instErr = cudaGraphInstantiate(&m_GraphExec[variant][parity], graph, 0);
	cudaGraphDestroy(graph);
	checkCuda(instErr, "cudaGraphInstantiate");
}

// ---------------------------------------------------------------------------
// per-frame
// ---------------------------------------------------------------------------

void Engine::bindImages(int n, const ju_image *inputs, const ju_image *outputs) {
	const std::size_t H = m_Spec.frameH, W = m_Spec.frameW;
	const std::size_t inRow = W * 4, outRow = 4 * W * 4;
	FrameIO *io = m_IoHost.as<FrameIO>();
	m_LastOutputs.assign(outputs, outputs + n);
	m_OutputNeedsCopy.assign(n, false);
	m_OutputPooled.assign(n, false);
	m_OutputArrays.assign(n, nullptr);
	std::vector<int> pooledInputs;
	// GRAPHICS_RESOURCE images (reference core/src/cuda_convert.cc.cu:381-397, 420-436): ptr is a
	// registered cudaGraphicsResource_t (getGLImage / getD3D11Image); it is mapped for the duration
	// of the frame and its array copied to / from the staging buffers on the frame's stream
	auto mappedArray = [&](void *ptr, std::size_t w, std::size_t h) -> cudaArray_t {
		auto res = static_cast<cudaGraphicsResource_t>(ptr);
		JU_CUDA(cudaGraphicsMapResources(1, &res, m_Stream));
		m_MappedResources.push_back(res);
		cudaArray_t array = nullptr;
		JU_CUDA(cudaGraphicsSubResourceGetMappedArray(&array, res, 0, 0));
		cudaChannelFormatDesc fmt{};
		cudaExtent extent{};
		JU_CUDA(cudaArrayGetInfo(&fmt, &extent, nullptr, array));
		if (extent.width != w || extent.height != h || fmt.f != cudaChannelFormatKindUnsigned || fmt.x != 8 ||
		    fmt.y != 8 || fmt.z != 8 || fmt.w != 8) {
			throw std::invalid_argument("graphics resource must be a 4 x 8-bit image of the model's size");
		}
		return array;
	};
	for (int s = 0; s < m_Batch; ++s) {
		std::uint8_t *inStage = m_InStage.as<std::uint8_t>() + s * H * inRow;
		std::uint8_t *outStage = m_OutStage.as<std::uint8_t>() + s * 4 * H * outRow;
		FrameIO &f = io[s];
		if (s >= n) {
			f.in = inStage;
			f.in_stride = static_cast<long long>(inRow);
			f.out = outStage;
			f.out_stride = static_cast<long long>(outRow);
			continue;
		}
		const ju_image &in = inputs[s];
		const ju_image &out = outputs[s];
		if (in.width != W || in.height != H || out.width != 4 * W || out.height != 4 * H) {
			throw std::invalid_argument("image size does not match the model (input " +
			                            std::to_string(W) + "x" + std::to_string(H) + ", output " +
			                            std::to_string(4 * W) + "x" + std::to_string(4 * H) + ")");
		}
		if (!in.ptr || !out.ptr) throw std::invalid_argument("null image pointer");
		auto absStride = [](std::int64_t v) { return static_cast<std::size_t>(v < 0 ? -v : v); };
		if (in.location == JU_LOC_CPU) {
			if (absStride(in.stride) < inRow) throw std::invalid_argument("input stride smaller than a row");
			const auto *p = static_cast<const std::uint8_t *>(in.ptr);
			if (m_CopyThreads > 0 && isPageable(p)) {
				// pageable caller memory: gather the rows into the engine's pinned buffer on the copy
				// pool (all streams in parallel), then one asynchronous host-to-device copy below
				ensureHostStaging();
				CopyJob job;
				job.dst = m_InPinned.as<std::uint8_t>() + s * H * inRow;
				job.src = p;
				job.dstStride = static_cast<std::ptrdiff_t>(inRow);
				job.srcStride = static_cast<std::ptrdiff_t>(in.stride);  // image row i lives at ptr + i * stride
				job.rowBytes = inRow;
				job.rows = H;
				m_Pool->submit(job);
				pooledInputs.push_back(s);
				f.in = inStage;
				f.in_stride = static_cast<long long>(inRow);
			} else if (in.stride == static_cast<std::int64_t>(inRow)) {
				JU_CUDA(cudaMemcpyAsync(inStage, p, inRow * H, cudaMemcpyHostToDevice, m_Stream));
				f.in = inStage;
				f.in_stride = static_cast<long long>(inRow);
			} else if (in.stride >= 0) {
				JU_CUDA(cudaMemcpy2DAsync(inStage, inRow, p, in.stride, inRow, H, cudaMemcpyHostToDevice, m_Stream));
				f.in = inStage;
				f.in_stride = static_cast<long long>(inRow);
			} else {
				// bottom-up: ptr addresses the last memory row (avisynth main.cc:125-142)
				const std::uint8_t *lowest = p + static_cast<std::int64_t>(H - 1) * in.stride;
				JU_CUDA(cudaMemcpy2DAsync(inStage, inRow, lowest, -in.stride, inRow, H, cudaMemcpyHostToDevice, m_Stream));
				f.in = inStage + (H - 1) * inRow;
				f.in_stride = -static_cast<long long>(inRow);
			}
		} else if (in.location == JU_LOC_CUDA) {
			// the kernels read / write whole BGRX pixels: 4-byte aligned rows of at least one row of pixels
			if (absStride(in.stride) < inRow) throw std::invalid_argument("input stride smaller than a row");
			if ((reinterpret_cast<std::uintptr_t>(in.ptr) | absStride(in.stride)) & 3u) {
				throw std::invalid_argument("CUDA input image must be 4-byte aligned (pointer and stride)");
			}
			f.in = static_cast<const std::uint8_t *>(in.ptr);
			f.in_stride = in.stride;
		} else if (in.location == JU_LOC_GRAPHICS_RESOURCE) {
			cudaArray_t array = mappedArray(in.ptr, W, H);
			JU_CUDA(cudaMemcpy2DFromArrayAsync(inStage, inRow, array, 0, 0, inRow, H, cudaMemcpyDeviceToDevice, m_Stream));
			f.in = inStage;
			f.in_stride = static_cast<long long>(inRow);
		} else {
			throw std::invalid_argument("unknown image location");
		}
		if (out.location == JU_LOC_CPU) {
			if (absStride(out.stride) < outRow) throw std::invalid_argument("output stride smaller than a row");
			m_OutputNeedsCopy[s] = true;
			if (m_CopyThreads > 0 && isPageable(out.ptr)) {
				// pageable caller memory: device -> pinned (asynchronous, band by band), then the copy
				// pool scatters the rows into the caller's image (process())
				ensureHostStaging();
				m_OutputPooled[s] = true;
				f.out = outStage;
				f.out_stride = static_cast<long long>(outRow);
			} else if (out.stride >= 0) {
				f.out = outStage;
				f.out_stride = static_cast<long long>(outRow);
			} else {
				f.out = outStage + (4 * H - 1) * outRow;
				f.out_stride = -static_cast<long long>(outRow);
			}
		} else if (out.location == JU_LOC_CUDA) {
			if (absStride(out.stride) < outRow) throw std::invalid_argument("output stride smaller than a row");
			if ((reinterpret_cast<std::uintptr_t>(out.ptr) | absStride(out.stride)) & 3u) {
				throw std::invalid_argument("CUDA output image must be 4-byte aligned (pointer and stride)");
			}
			f.out = static_cast<std::uint8_t *>(out.ptr);
			f.out_stride = out.stride;
		} else if (out.location == JU_LOC_GRAPHICS_RESOURCE) {
			m_OutputArrays[s] = mappedArray(out.ptr, 4 * W, 4 * H);
			f.out = outStage;
			f.out_stride = static_cast<long long>(outRow);
		} else {
			throw std::invalid_argument("unknown image location");
		}
	}
	if (!pooledInputs.empty()) {
		m_Pool->end();
		for (int s : pooledInputs) {
			JU_CUDA(cudaMemcpyAsync(m_InStage.as<std::uint8_t>() + s * H * inRow, m_InPinned.as<std::uint8_t>() + s * H * inRow,
			    H * inRow, cudaMemcpyHostToDevice, m_Stream));
		}
	}
	// the device-side address table only changes when the caller's pointers do: host images always
	// go through the same staging buffers, so their frames skip this copy
	const std::size_t tableBytes = sizeof(FrameIO) * m_Batch;
	if (m_IoShadow.size() != tableBytes || std::memcmp(m_IoShadow.data(), io, tableBytes) != 0) {
		JU_CUDA(cudaMemcpyAsync(m_IoDev.get(), io, tableBytes, cudaMemcpyHostToDevice, m_Stream));
		m_IoShadow.assign(reinterpret_cast<const unsigned char *>(io), reinterpret_cast<const unsigned char *>(io) + tableBytes);
	}
}

void Engine::uploadStatus(int inject) {
	TcStatus st{};
	st.code = 0;
	st.timeout_ms = m_WaitTimeoutMs;
	st.inject = inject;
	st.host_code = m_StatusHost.as<int>();  // pinned host memory is device-addressable (unified addressing)
	JU_CUDA(cudaMemcpy(m_Status.get(), &st, sizeof(st), cudaMemcpyHostToDevice));
}

void Engine::injectStall(int kernelId) {
	DeviceGuard guard(m_Device);
	JU_CUDA(cudaStreamSynchronize(m_Stream));
	uploadStatus(kernelId);
}

// A pipeline wait expired during the frame (TcStatus): the kernels drained without side effects.
// Re-arm everything the aborted frame may have left half-way - the status block and the
// never-reset dataflow counters of the persistent trunks - so that the next frame starts clean.
void Engine::recoverFromStall() {
	m_StatusHost.as<int>()[0] = m_StatusHost.as<int>()[1] = 0;
	uploadStatus(0);
	for (TrunkState *ts : {&m_GenTrunk, &m_FlowTrunk}) {
		if (ts->counter.get()) JU_CUDA(cudaMemset(ts->counter.get(), 0, ts->counter.bytes()));
		if (ts->flags.get()) JU_CUDA(cudaMemset(ts->flags.get(), 0, ts->flags.bytes()));
	}
	if (m_FlowCounters.get()) {
		JU_CUDA(cudaMemset(m_FlowCounters.get(), 0, m_FlowCounters.bytes()));
		JU_CUDA(cudaMemset(m_FlowSync.get(), 0, m_FlowSync.bytes()));
	}
}

// cudaPointerGetAttributes: ordinary heap memory is "unregistered"; pinned / registered host memory
// can be the target of a truly asynchronous copy and keeps the direct path
bool Engine::isPageable(const void *ptr) {
	cudaPointerAttributes attr{};
	const cudaError_t e = cudaPointerGetAttributes(&attr, ptr);
	if (e != cudaSuccess) {
		(void) cudaGetLastError();
		return true;
	}
	return attr.type == cudaMemoryTypeUnregistered;
}

void Engine::ensureHostStaging() {
	if (m_Pool) return;
	const std::size_t H = m_Spec.frameH, W = m_Spec.frameW;
	m_InPinned = PinnedBuffer(static_cast<std::size_t>(m_Batch) * H * W * 4);
	m_OutPinned = PinnedBuffer(static_cast<std::size_t>(m_Batch) * 16 * H * W * 4);
	m_Pool = std::make_unique<HostCopyPool>(m_CopyThreads);
}

void Engine::unmapResources() {
	if (m_MappedResources.empty()) return;
	std::vector<cudaGraphicsResource_t> res;
	res.swap(m_MappedResources);
	JU_CUDA(cudaGraphicsUnmapResources(static_cast<int>(res.size()), res.data(), m_Stream));
}

void Engine::process(int n, const ju_image *inputs, const ju_image *outputs) {
	if (n < 1 || n > m_Batch) throw std::invalid_argument("image count must be in [1, batch]");
	m_BandEventsUsed = 0;
	DeviceGuard guard(m_Device);
	std::lock_guard<std::mutex> turn(deviceMutex(m_Device));  // see deviceMutex: one frame at a time per device
	const std::size_t H = m_Spec.frameH, W = m_Spec.frameW, outRow = 4 * W * 4;
	try {
		bindImages(n, inputs, outputs);
		bool anyCopy = false;
		for (int s = 0; s < n; ++s) anyCopy = anyCopy || m_OutputNeedsCopy[s];
		const int variant = (m_HostVariant && anyCopy) ? 1 : 0;
		if (m_UseGraph) {
			JU_CUDA(cudaGraphLaunch(m_GraphExec[variant][m_Parity], m_Stream));
		} else {
			for (const Op &op : m_Plans[variant][m_Parity]) checkCuda(op.run(m_Stream), op.name.c_str());
		}
		// Device-to-host copies of staged images run on a second stream, one group of streams at a
		// time as soon as the frame graph has recorded that group's completion event: with several
		// sub-batches the copies of the first streams overlap the trunk of the later ones.
		bool copied = false;
		m_BandCopies.clear();
		for (const Region &r : m_Regions[variant]) {
			const int b0 = r.b0, b1 = std::min(n, r.b0 + r.nb);
			bool any = false;
			for (int s = b0; s < b1; ++s) any = any || m_OutputNeedsCopy[s];
			if (!any) continue;
			JU_CUDA(cudaStreamWaitEvent(m_CopyStream, r.ev, 0));
			const std::size_t rows = static_cast<std::size_t>(r.row1 - r.row0);
			for (int s = b0; s < b1; ++s) {
				if (!m_OutputNeedsCopy[s]) continue;
				const ju_image &out = outputs[s];
				const std::uint8_t *stage = m_OutStage.as<std::uint8_t>() + s * 4 * H * outRow;
				auto *p = static_cast<std::uint8_t *>(out.ptr);
				if (m_OutputPooled[s]) {
					std::uint8_t *pinned = m_OutPinned.as<std::uint8_t>() + s * 4 * H * outRow + r.row0 * outRow;
					JU_CUDA(cudaMemcpyAsync(pinned, stage + r.row0 * outRow, rows * outRow, cudaMemcpyDeviceToHost, m_CopyStream));
					// the rows are scattered into the caller's image as soon as the band has arrived in
					// pinned memory (event below); image row i lives at ptr + i * stride
					BandCopy band;
					band.job.dst = p + static_cast<std::ptrdiff_t>(r.row0) * static_cast<std::ptrdiff_t>(out.stride);
					band.job.src = pinned;
					band.job.dstStride = static_cast<std::ptrdiff_t>(out.stride);
					band.job.srcStride = static_cast<std::ptrdiff_t>(outRow);
					band.job.rowBytes = outRow;
					band.job.rows = rows;
					band.event = nullptr;
					m_BandCopies.push_back(band);
				} else if (out.stride == static_cast<std::int64_t>(outRow)) {
					// dense image: one linear copy
					JU_CUDA(cudaMemcpyAsync(p + r.row0 * outRow, stage + r.row0 * outRow, rows * outRow,
					    cudaMemcpyDeviceToHost, m_CopyStream));
				} else if (out.stride >= 0) {
					JU_CUDA(cudaMemcpy2DAsync(p + static_cast<std::int64_t>(r.row0) * out.stride, out.stride,
					    stage + r.row0 * outRow, outRow, outRow, rows, cudaMemcpyDeviceToHost, m_CopyStream));
				} else {
					// bottom-up: image row i lives in memory row 4H-1-i of both the staging buffer and the
					// caller's buffer, so image rows [row0, row1) are memory rows [4H-row1, 4H-row0)
					const std::size_t m0 = 4 * H - static_cast<std::size_t>(r.row1);
					std::uint8_t *lowest = p + static_cast<std::int64_t>(4 * H - 1) * out.stride;
					JU_CUDA(cudaMemcpy2DAsync(lowest + m0 * static_cast<std::size_t>(-out.stride), -out.stride,
					    stage + m0 * outRow, outRow, outRow, rows, cudaMemcpyDeviceToHost, m_CopyStream));
				}
				copied = true;
			}
			if (!m_BandCopies.empty() && !m_BandCopies.back().event) {
				// one event per region: every pinned copy of the region queued above has landed
				const std::size_t used = m_BandEventsUsed++;
				if (used == m_BandEvents.size()) {
					cudaEvent_t ev = nullptr;
					JU_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming | cudaEventBlockingSync));
					m_BandEvents.push_back(ev);
				}
				JU_CUDA(cudaEventRecord(m_BandEvents[used], m_CopyStream));
				for (auto it = m_BandCopies.rbegin(); it != m_BandCopies.rend() && !it->event; ++it) it->event = m_BandEvents[used];
			}
		}
		// graphics-resource outputs: staging buffer -> mapped array, then unmap, all in stream order
		for (int s = 0; s < n; ++s) {
			if (!m_OutputArrays[s]) continue;
			const std::uint8_t *stage = m_OutStage.as<std::uint8_t>() + s * 4 * H * outRow;
			JU_CUDA(cudaMemcpy2DToArrayAsync(m_OutputArrays[s], 0, 0, stage, outRow, outRow, 4 * H,
			    cudaMemcpyDeviceToDevice, m_Stream));
		}
		unmapResources();
		if (!m_BandCopies.empty()) {
			// pageable outputs: poll the bands in stream order, hand each to the copy pool as it
			// arrives and copy along while waiting for the next one
			m_Pool->begin();
			cudaEvent_t ready = nullptr;
			for (const BandCopy &band : m_BandCopies) {
				if (band.event != ready) {
					const auto t0 = std::chrono::steady_clock::now();
					for (;;) {
						const cudaError_t q = cudaEventQuery(band.event);
						if (q == cudaSuccess) break;
						if (q != cudaErrorNotReady) checkCuda(q, "cudaEventQuery");
						if (m_Pool->help()) continue;
						if (std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(m_CopySpinUs)) {
							// nothing to copy and the band is far away: park the workers and sleep on the event
							m_Pool->end();
							JU_CUDA(cudaEventSynchronize(band.event));
							m_Pool->begin();
							break;
						}
					}
					ready = band.event;
				}
				m_Pool->submit(band.job);
			}
			m_Pool->end();
		}
		JU_CUDA(cudaStreamSynchronize(m_Stream));
		if (copied) JU_CUDA(cudaStreamSynchronize(m_CopyStream));
		const int stall = *static_cast<volatile int *>(m_StatusHost.as<int>());
		if (stall != 0) {
			const int where = static_cast<volatile int *>(m_StatusHost.as<int>())[1];
			TcStatus snapshot{};
			cudaMemcpy(&snapshot, m_Status.get(), sizeof(snapshot), cudaMemcpyDeviceToHost);
			std::string pending;
			for (int c = 0; c < 32; ++c) {
				if (snapshot.pending & (1u << c)) pending += (pending.empty() ? "" : ",") + std::to_string(c);
			}
			recoverFromStall();
			throw KernelStallException(std::string("frame aborted: ") + tc_kernel_name(stall >> 8) +
			                           " pipeline wait " + std::to_string(stall & 0xff) + " expired (layer " +
			                           std::to_string(where >> 16) + ", CTA " + std::to_string(where & 0xffff) + "; blocked waits: " + pending + ")");
		}
	} catch (...) {
		// a failed frame leaves the ping-pong index unchanged (the reference flips only after the
		// synchronize, tensorrt_backend.cc:276-277); nothing may still be writing into the caller's
		// images when the exception leaves this function
		try {
			unmapResources();
		} catch (...) {
		}
		cudaStreamSynchronize(m_Stream);
		cudaStreamSynchronize(m_CopyStream);
		if (m_Pool) m_Pool->end();
		throw;
	}
	m_Parity ^= 1;
}

void Engine::resetState() {
	DeviceGuard guard(m_Device);
	JU_CUDA(cudaStreamSynchronize(m_Stream));
	for (int p = 0; p < 2; ++p) {
		JU_CUDA(cudaMemset(m_FlowIn[p].get(), 0, m_FlowIn[p].bytes()));
		JU_CUDA(cudaMemset(m_PreGen[p].get(), 0, m_PreGen[p].bytes()));
	}
	m_Parity = 0;
}

void Engine::readTensor(const std::string &name, void *dst, std::uint64_t capacity, ju_tensor_desc *desc) {
	auto it = m_Tensors.find(name);
	if (it == m_Tensors.end()) throw std::invalid_argument("unknown tensor " + name);
	const NamedTensor &t = it->second;
	if (desc) {
		desc->dtype = static_cast<std::uint32_t>(t.dtype);
		desc->ndim = static_cast<std::uint32_t>(t.dims.size());
		for (std::size_t i = 0; i < 4; ++i) desc->dims[i] = i < t.dims.size() ? t.dims[i] : 1;
		desc->bytes = t.bytes;
	}
	if (!dst) return;
	if (capacity < t.bytes) throw std::invalid_argument("destination too small for " + name);
	DeviceGuard guard(m_Device);
	// ping-pong tensors: the set the NEXT frame will read (= newest state)
	void *src = t.pingPong ? t.ptr[m_Parity] : t.ptr[0];
	JU_CUDA(cudaStreamSynchronize(m_Stream));
	JU_CUDA(cudaMemcpy(dst, src, t.bytes, cudaMemcpyDeviceToHost));
}

void Engine::writeState(const std::string &name, const void *srcHost, std::uint64_t bytes) {
	auto it = m_Tensors.find(name);
	if (it == m_Tensors.end() || !it->second.writable) throw std::invalid_argument("not a state tensor: " + name);
	const NamedTensor &t = it->second;
	if (bytes != t.bytes) throw std::invalid_argument("size mismatch for " + name);
	DeviceGuard guard(m_Device);
	JU_CUDA(cudaStreamSynchronize(m_Stream));
	JU_CUDA(cudaMemcpy(t.ptr[m_Parity], srcHost, bytes, cudaMemcpyHostToDevice));
}

std::vector<ju_op_time> Engine::profileOps(int iters) {
	if (iters < 1) iters = 1;
	DeviceGuard guard(m_Device);
	std::lock_guard<std::mutex> turn(deviceMutex(m_Device));
	JU_CUDA(cudaStreamSynchronize(m_Stream));
	const std::vector<Op> &plan = m_Plans[0][m_Parity];
	std::vector<cudaEvent_t> ev(plan.size() + 1);
	for (auto &e : ev) JU_CUDA(cudaEventCreate(&e));
	std::vector<double> total(plan.size(), 0.0);
	// state is advanced in place on parity m_Parity without flipping: the
	// timings are data-independent, and the caller resets state afterwards
	for (int it = -2; it < iters; ++it) {
		JU_CUDA(cudaEventRecord(ev[0], m_Stream));
		for (std::size_t i = 0; i < plan.size(); ++i) {
			checkCuda(plan[i].run(m_Stream), plan[i].name.c_str());
			JU_CUDA(cudaEventRecord(ev[i + 1], m_Stream));
		}
		JU_CUDA(cudaStreamSynchronize(m_Stream));
		if (it < 0) continue;  // warm-up
		for (std::size_t i = 0; i < plan.size(); ++i) {
			float ms = 0.f;
			JU_CUDA(cudaEventElapsedTime(&ms, ev[i], ev[i + 1]));
			total[i] += ms * 1000.0;
		}
	}
	std::vector<ju_op_time> result(plan.size());
	for (std::size_t i = 0; i < plan.size(); ++i) {
		ju_op_time &o = result[i];
		std::memset(&o, 0, sizeof(o));
		std::strncpy(o.name, plan[i].name.c_str(), sizeof(o.name) - 1);
		o.usec = total[i] / iters;
		o.flops = plan[i].flops;
		o.bytes = plan[i].bytes;
		o.tensor_bound = plan[i].tensorBound ? 1 : 0;
		o.reserved = plan[i].layers;
	}
	// Second pass: events only at group boundaries, so kernels inside a group
	// run back to back exactly as in the replayed graph (programmatic dependent
	// launch overlaps their prologues).  "group:<label>" entries carry the
	// group's total time per frame and its launch count in `reserved`.
	auto groupOf = [](const std::string &n) -> std::string {
		if (n.rfind("generator/block_", 0) == 0) return "resblocks";
		if (n.rfind("flow/", 0) == 0) return "flow";
		if (n.rfind("sync:", 0) == 0) return "sync";
		return n;
	};
	std::vector<std::string> labels;
	std::vector<std::size_t> firstOp;
	for (std::size_t i = 0; i < plan.size(); ++i) {
		std::string g = groupOf(plan[i].name);
		if (labels.empty() || labels.back() != g) {
			labels.push_back(g);
			firstOp.push_back(i);
		}
	}
	firstOp.push_back(plan.size());
	std::vector<double> gtotal(labels.size(), 0.0);
	// each group is captured into its own CUDA graph, so the timing includes the
	// same device-side launch path (and programmatic edges) as the frame graph
	// and none of the host's per-launch cost
	std::vector<cudaGraphExec_t> execs(labels.size(), nullptr);
	auto destroyExecs = [&] {
		for (auto &e : execs)
			if (e) cudaGraphExecDestroy(e);
	};
	try {
		for (std::size_t g = 0; g < labels.size(); ++g) {
			cudaGraph_t graph = nullptr;
			JU_CUDA(cudaStreamBeginCapture(m_Stream, cudaStreamCaptureModeThreadLocal));
			cudaError_t err = cudaSuccess;
			for (std::size_t i = firstOp[g]; i < firstOp[g + 1] && err == cudaSuccess; ++i) err = plan[i].run(m_Stream);
			cudaError_t endErr = cudaStreamEndCapture(m_Stream, &graph);
			checkCuda(err, "kernel launch during group capture");
			checkCuda(endErr, "cudaStreamEndCapture");
			cudaError_t instErr = cudaGraphInstantiate(&execs[g], graph, 0);
			cudaGraphDestroy(graph);
			checkCuda(instErr, "cudaGraphInstantiate");
		}
		for (int it = -2; it < iters; ++it) {
			for (std::size_t g = 0; g < labels.size(); ++g) {
				JU_CUDA(cudaEventRecord(ev[g], m_Stream));
				JU_CUDA(cudaGraphLaunch(execs[g], m_Stream));
			}
			JU_CUDA(cudaEventRecord(ev[labels.size()], m_Stream));
			JU_CUDA(cudaStreamSynchronize(m_Stream));
			if (it < 0) continue;
			for (std::size_t g = 0; g < labels.size(); ++g) {
				float ms = 0.f;
				JU_CUDA(cudaEventElapsedTime(&ms, ev[g], ev[g + 1]));
				gtotal[g] += ms * 1000.0;
			}
		}
	} catch (...) {
		destroyExecs();
		throw;
	}
	destroyExecs();
	// a label can occur several times in the plan (per sub-batch trunk / tail): one entry per label
	std::vector<std::string> unique;
	for (std::size_t g = 0; g < labels.size(); ++g) {
		std::size_t u = 0;
		for (; u < unique.size(); ++u)
			if (unique[u] == labels[g]) break;
		if (u == unique.size()) {
			unique.push_back(labels[g]);
			ju_op_time o;
			std::memset(&o, 0, sizeof(o));
			std::snprintf(o.name, sizeof(o.name), "group:%s", labels[g].c_str());
			result.push_back(o);
		}
		ju_op_time &o = result[plan.size() + u];
		o.usec += gtotal[g] / iters;
		for (std::size_t i = firstOp[g]; i < firstOp[g + 1]; ++i) {
			o.flops += plan[i].flops;
			o.bytes += plan[i].bytes;
			o.tensor_bound |= plan[i].tensorBound ? 1 : 0;
			o.reserved += plan[i].layers;
		}
	}
	for (auto &e : ev) cudaEventDestroy(e);
	return result;
}

}  // namespace ju
