// `.jup` model container reader and BatchNorm folding.
//
// The container stands where the reference keeps a serialized TensorRT engine
// (file read in TensorRTRuntime's ctor, core/src/core.cc:154-167): it holds
// the hyper-parameters of get_flow_autoencoder / get_flow_resnet /
// get_generator_resnet / get_inference_model (scripts/training/models.py:
// 257-263, 334-339, 484-491, 680-689) and the raw Keras-layout fp32 tensors.
// Written by joshupscale_b200/weights.py.
#pragma once

#include <cstdint>
#include <map>
#include <string>
#include <vector>

namespace ju {

struct ModelSpec {
	int frameH = 0, frameW = 0, padH = 0, padW = 0;
	int flowArch = 0;  // 0 autoencoder, 1 resnet
	int flowInputs = 4;
	std::vector<int> flowFilters;  // autoencoder: filter list; resnet: {filters, blocks}
	int genFilters = 64, genBlocks = 24;
	int actFlow = 0, actGen = 0;  // 0 relu, 1 lrelu
	float slopeFlow = 0.3f, slopeGen = 0.3f;
	bool normalizeBrightness = false;
	float bnEps = 1e-3f;

	double flowGmacs() const;
	double genGmacs() const;
};

struct HostTensor {
	std::vector<int> dims;
	std::vector<float> data;
};

// conv (+ folded BN) in the generic (tap, Cin, Cout) fp32 form all packers take
struct FoldedConv {
	int ksize = 3, cin = 0, cout = 0;
	std::vector<float> kernel;  // [ksize*ksize][cin][cout], UNscaled
	std::vector<float> scale;   // [cout]  gamma / sqrt(var + eps)   (1 if no BN)
	std::vector<float> bias;    // [cout]  beta - mean * scale (+ conv bias)
};

class ModelFile {
public:
	static ModelFile load(const std::string &path);

	const ModelSpec &spec() const { return m_Spec; }
	const HostTensor &tensor(const std::string &name) const;
	bool has(const std::string &name) const { return m_Tensors.count(name) != 0; }

	// Conv2D `conv` (+ BatchNormalization `bn` if non-empty; + bias if present)
	FoldedConv foldConv(const std::string &conv, const std::string &bn) const;
	// Conv2DTranspose(k2,s2) `conv` (+BN) expressed as a 1x1 conv to 4*Cout
	// channels ordered (i*2+j)*Cout + o (pixel-shuffle store)
	FoldedConv foldConvTranspose(const std::string &conv, const std::string &bn) const;

private:
	ModelSpec m_Spec;
	std::map<std::string, HostTensor> m_Tensors;
};

}  // namespace ju
