// Drives the REFERENCE's AviSynth filter (avisynth_plugin/src/main.cc, compiled unmodified
// against oracle/ref_avisynth/avisynth.h and this repo's include/JoshUpscale/core.h) with the
// frame requests given on the command line and prints, per request,
//   <n> -> <output id> | <source frame indices passed to processImage>
// for comparison with include/JoshUpscale/sequencer.h (tests/test_sequencer.py).
// TEST INFRASTRUCTURE ONLY: the runtime below is a recording fake, not the CUDA engine.
#include <JoshUpscale/core.h>
#include <avisynth.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

extern "C" const char *AvisynthPluginInit3(IScriptEnvironment *env, const AVS_Linkage *const vectors);

namespace {

constexpr int kW = 8, kH = 4;
std::vector<int> g_Processed;
int g_Calls = 0;

struct FakeRuntime : JoshUpscale::core::Runtime {
	FakeRuntime() {
		m_InputWidth = kW;
		m_InputHeight = kH;
		m_OutputWidth = 4 * kW;
		m_OutputHeight = 4 * kH;
	}
	void processImage(const JoshUpscale::core::Image &in, const JoshUpscale::core::Image &out) override {
		int src = 0;
		std::memcpy(&src, in.ptr, sizeof(src));  // first pixel of image row 0
		g_Processed.push_back(src);
		const int id = 1000 * (++g_Calls) + src;
		std::memcpy(out.ptr, &id, sizeof(id));
	}
};

// a clip whose frame n is filled with the 32-bit value n
class SourceClip : public IClip {
public:
	SourceClip() {
		m_Vi.width = kW;
		m_Vi.height = kH;
		m_Vi.num_frames = 1 << 20;
	}
	PVideoFrame __stdcall GetFrame(int n, IScriptEnvironment *) override {
		auto f = std::make_shared<VideoFrame>(kW * 4, kH);
		int *p = reinterpret_cast<int *>(f->GetWritePtr());
		for (int i = 0; i < kW * kH; ++i) p[i] = n;
		return f;
	}
	const VideoInfo &__stdcall GetVideoInfo() override { return m_Vi; }

private:
	VideoInfo m_Vi;
};

}  // namespace

namespace JoshUpscale {
namespace core {
Runtime *createRuntime(int, const std::filesystem::path &) { return new FakeRuntime(); }
std::string getExceptionString() { return "exception"; }
}  // namespace core
}  // namespace JoshUpscale

int main(int argc, char **argv) {
	IScriptEnvironment env;
	AvisynthPluginInit3(&env, nullptr);
	if (env.m_Name != "JoshUpscale" || !env.m_Apply) return 2;
	AVSValue args(std::vector<AVSValue>{AVSValue(PClip(new SourceClip())), AVSValue("model.jup"), AVSValue()});
	PClip filter = env.m_Apply(args, env.m_UserData, &env).AsClip();
	const VideoInfo &vi = filter->GetVideoInfo();
	if (vi.width != 4 * kW || vi.height != 4 * kH) return 3;
	for (int i = 1; i < argc; ++i) {
		g_Processed.clear();
		const int n = std::atoi(argv[i]);
		PVideoFrame out = filter->GetFrame(n, &env);
		int id = 0;
		// bottom-up RGB32: image row 0 is the last memory row (main.cc:133-142)
		std::memcpy(&id, out->GetReadPtr() + static_cast<size_t>(vi.height - 1) * out->GetPitch(), sizeof(id));
		std::printf("%d -> %d |", n, id);
		for (int p : g_Processed) std::printf(" %d", p);
		std::printf("\n");
	}
	return 0;
}
