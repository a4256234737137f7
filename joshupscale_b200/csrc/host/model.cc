#include "model.h"

#include <cmath>
#include <cstring>
#include <fstream>

#include "common.h"

namespace ju {

namespace {

#pragma pack(push, 1)
struct RawHeader {
	char magic[8];
	std::uint32_t version, headerBytes;
	std::uint32_t frameH, frameW, padH, padW;
	std::uint32_t flowArch, flowInputs;
	std::uint32_t nFilters;
	std::uint32_t filters[16];
	std::uint32_t genFilters, genBlocks;
	std::uint32_t actFlow;
	float slopeFlow;
	std::uint32_t actGen;
	float slopeGen;
	std::uint32_t normalizeBrightness;
	float bnEps;
	std::uint32_t nTensors;
};
struct RawEntry {
	char name[96];
	std::uint32_t dtype, ndim;
	std::uint32_t dims[4];
	std::uint64_t offset, nbytes;
};
#pragma pack(pop)

constexpr char kMagic[8] = {'J', 'U', 'P', 'M', 'D', 'L', 0, 1};

}  // namespace

double ModelSpec::flowGmacs() const {
	double macs = 0;
	double cin = 3.0 * flowInputs;
	double h = padH, w = padW;
	if (flowArch == 0) {
		const auto &f = flowFilters;
		int n = static_cast<int>(f.size()) / 2;
		for (int i = 0; i < 2 * n; ++i) {
			macs += h * w * 9 * (cin * f[i] + double(f[i]) * f[i]);
			cin = f[i];
			if (i < n) {
				h /= 2;
				w /= 2;
			} else {
				h *= 2;
				w *= 2;
			}
		}
		if (f.size() % 2) {
			macs += h * w * 9 * cin * f.back();
			cin = f.back();
		}
		macs += h * w * 9 * cin * 32;
	} else {
		double nf = flowFilters[0], nb = flowFilters[1];
		macs += h * w * 9 * cin * nf + h * w * 9 * nf * nf * 2 * nb + h * w * nf * 32;
	}
	return macs / 1e9;
}

double ModelSpec::genGmacs() const {
	double h = frameH, w = frameW, nf = genFilters;
	double macs = h * w * 9 * 51 * nf + h * w * 9 * nf * nf * 2 * genBlocks + h * w * nf * 32 * 4 +
	              (2 * h) * (2 * w) * 32 * 3 * 4;
	return macs / 1e9;
}

ModelFile ModelFile::load(const std::string &path) {
	std::ifstream file(path, std::ifstream::binary | std::ifstream::ate);
	if (!file) throw ModelException("cannot open model file: " + path);
	auto size = static_cast<std::size_t>(file.tellg());
	std::vector<char> data(size);
	file.seekg(0);
	file.read(data.data(), static_cast<std::streamsize>(size));
	if (!file) throw ModelException("cannot read model file: " + path);
	if (size < sizeof(RawHeader)) throw ModelException("model file too small: " + path);
	RawHeader h;
	std::memcpy(&h, data.data(), sizeof(h));
	if (std::memcmp(h.magic, kMagic, 8) != 0 || h.version != 1) {
		throw ModelException(
		    "not a .jup model container (the B200 build does not load TensorRT engines): " + path);
	}
	// Nothing in the file is trusted: every size is bounded and every offset checked without
	// arithmetic that can wrap, so a corrupt or crafted container ends in a ModelException.
	if (h.headerBytes < sizeof(RawHeader) || h.headerBytes > size) throw ModelException("invalid header size");
	constexpr std::uint32_t kMaxFrame = 8192, kMaxChannels = 1024, kMaxBlocks = 256, kMaxTensors = 8192;
	if (h.frameH > kMaxFrame || h.frameW > kMaxFrame || h.padH > kMaxFrame + 256 || h.padW > kMaxFrame + 256 ||
	    h.flowInputs > 21 || h.genFilters < 1 || h.genFilters > kMaxChannels || h.genBlocks > kMaxBlocks ||
	    h.nTensors > kMaxTensors) {
		throw ModelException("model header out of range");
	}
	ModelFile m;
	ModelSpec &s = m.m_Spec;
	s.frameH = h.frameH;
	s.frameW = h.frameW;
	s.padH = h.padH;
	s.padW = h.padW;
	s.flowArch = h.flowArch;
	s.flowInputs = h.flowInputs;
	if (h.nFilters > 16) throw ModelException("bad filter count");
	s.flowFilters.assign(h.filters, h.filters + h.nFilters);
	s.genFilters = h.genFilters;
	s.genBlocks = h.genBlocks;
	s.actFlow = h.actFlow;
	s.slopeFlow = h.slopeFlow;
	s.actGen = h.actGen;
	s.slopeGen = h.slopeGen;
	s.normalizeBrightness = h.normalizeBrightness != 0;
	s.bnEps = h.bnEps;
	if (s.frameH < 2 || s.frameW < 2 || s.padH < s.frameH || s.padW < s.frameW ||
	    s.flowInputs < 1 || s.flowArch > 1) {
		throw ModelException("invalid model header");
	}
	if (s.flowArch == 1) {
		// get_flow_resnet: {filters, blocks}
		if (s.flowFilters.size() != 2 || s.flowFilters[0] < 1 || s.flowFilters[0] > static_cast<int>(kMaxChannels) ||
		    s.flowFilters[1] < 0 || s.flowFilters[1] > static_cast<int>(kMaxBlocks)) {
			throw ModelException("flow-resnet header needs {filters, blocks}");
		}
	} else {
		for (int f : s.flowFilters) {
			if (f < 1 || f > static_cast<int>(kMaxChannels)) throw ModelException("flow filter count out of range");
		}
		// get_flow_autoencoder: n down blocks + n up blocks (+ one trailing conv); the padded frame
		// must survive n MaxPool2D(2)
		if (s.flowFilters.size() < 2) throw ModelException("flow-autoencoder header needs at least two filters");
		const int n = static_cast<int>(s.flowFilters.size()) / 2;
		if ((s.padH % (1 << n)) != 0 || (s.padW % (1 << n)) != 0) {
			throw ModelException("padded frame not divisible by 2^blocks");
		}
	}
	const std::size_t tableBytes = static_cast<std::size_t>(h.nTensors) * sizeof(RawEntry);
	if (tableBytes > size - h.headerBytes) throw ModelException("truncated tensor table");
	for (std::uint32_t i = 0; i < h.nTensors; ++i) {
		RawEntry e;
		std::memcpy(&e, data.data() + h.headerBytes + i * sizeof(RawEntry), sizeof(e));
		if (e.dtype != 0 || e.ndim > 4 || e.offset > size || e.nbytes > size - e.offset) {
			throw ModelException("invalid tensor entry");
		}
		HostTensor t;
		std::uint64_t count = 1;
		for (std::uint32_t d = 0; d < e.ndim; ++d) {
			if (e.dims[d] == 0 || e.dims[d] > (1u << 24)) throw ModelException("tensor dimension out of range");
			t.dims.push_back(static_cast<int>(e.dims[d]));
			count *= e.dims[d];
			if (count > (std::uint64_t{1} << 32)) throw ModelException("tensor too large");
		}
		if (count * sizeof(float) != e.nbytes) throw ModelException("tensor size mismatch");
		t.data.resize(count);
		std::memcpy(t.data.data(), data.data() + e.offset, e.nbytes);
		std::string name(e.name, strnlen(e.name, sizeof(e.name)));
		m.m_Tensors.emplace(std::move(name), std::move(t));
	}
	return m;
}

const HostTensor &ModelFile::tensor(const std::string &name) const {
	auto it = m_Tensors.find(name);
	if (it == m_Tensors.end()) throw ModelException("model file lacks tensor " + name);
	return it->second;
}

namespace {

// s = gamma / sqrt(var + eps); b = beta - mean * s   (fp32)
void bnFold(const ModelFile &m, const std::string &bn, int cout, std::vector<float> *scale,
    std::vector<float> *bias) {
	scale->assign(cout, 1.0f);
	bias->assign(cout, 0.0f);
	if (bn.empty()) return;
	const auto &gamma = m.tensor(bn + "/gamma").data;
	const auto &beta = m.tensor(bn + "/beta").data;
	const auto &mean = m.tensor(bn + "/moving_mean").data;
	const auto &var = m.tensor(bn + "/moving_variance").data;
	if (static_cast<int>(gamma.size()) != cout) throw ModelException("BN size mismatch: " + bn);
	const float eps = m.spec().bnEps;
	for (int o = 0; o < cout; ++o) {
		volatile float denom = std::sqrt(var[o] + eps);
		volatile float s = gamma[o] / denom;
		volatile float ms = mean[o] * s;  // separate rounding, matches the oracle
		(*scale)[o] = s;
		(*bias)[o] = beta[o] - ms;
	}
}

}  // namespace

FoldedConv ModelFile::foldConv(const std::string &conv, const std::string &bn) const {
	const HostTensor &k = tensor(conv + "/kernel");
	if (k.dims.size() != 4 || k.dims[0] != k.dims[1]) throw ModelException("bad kernel: " + conv);
	FoldedConv f;
	f.ksize = k.dims[0];
	f.cin = k.dims[2];
	f.cout = k.dims[3];
	f.kernel = k.data;  // (kh, kw, Cin, Cout) == [tap][cin][cout]
	bnFold(*this, bn, f.cout, &f.scale, &f.bias);
	if (has(conv + "/bias")) {
		const auto &b = tensor(conv + "/bias").data;
		for (int o = 0; o < f.cout; ++o) f.bias[o] += b[o] * f.scale[o];
	}
	return f;
}

FoldedConv ModelFile::foldConvTranspose(const std::string &conv, const std::string &bn) const {
	const HostTensor &k = tensor(conv + "/kernel");  // (2, 2, Cout, Cin)
	if (k.dims.size() != 4 || k.dims[0] != 2 || k.dims[1] != 2) {
		throw ModelException("bad transpose kernel: " + conv);
	}
	int cout = k.dims[2], cin = k.dims[3];
	FoldedConv f;
	f.ksize = 1;
	f.cin = cin;
	f.cout = 4 * cout;
	f.kernel.resize(static_cast<std::size_t>(cin) * 4 * cout);
	for (int q = 0; q < 4; ++q)
		for (int o = 0; o < cout; ++o)
			for (int c = 0; c < cin; ++c)
				f.kernel[static_cast<std::size_t>(c) * 4 * cout + q * cout + o] =
				    k.data[(static_cast<std::size_t>(q) * cout + o) * cin + c];
	std::vector<float> s, b;
	bnFold(*this, bn, cout, &s, &b);
	if (has(conv + "/bias")) {
		const auto &cb = tensor(conv + "/bias").data;
		for (int o = 0; o < cout; ++o) b[o] += cb[o] * s[o];
	}
	f.scale.resize(4 * cout);
	f.bias.resize(4 * cout);
	for (int q = 0; q < 4; ++q)
		for (int o = 0; o < cout; ++o) {
			f.scale[q * cout + o] = s[o];
			f.bias[q * cout + o] = b[o];
		}
	return f;
}

}  // namespace ju
