// Minimal stand-in for the AviSynth+ SDK header, just enough to COMPILE THE REFERENCE'S
// avisynth_plugin/src/main.cc UNMODIFIED on Linux (oracle/Makefile, target _ref/avisynth_trace).
// TEST INFRASTRUCTURE ONLY.  Written from the way the plugin uses the SDK, not from SDK sources:
// only the members main.cc touches exist.
#pragma once

#include <cstdarg>
#include <cstdio>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#ifndef _WIN32
#define __stdcall
#define __cdecl
#define __declspec(x)
#endif

typedef unsigned char BYTE;

enum { DEV_TYPE_NONE = 0, DEV_TYPE_CPU = 1, DEV_TYPE_CUDA = 2 };
enum {
	CACHE_GETCHILD_COST = 1,
	CACHE_GETCHILD_THREAD_MODE,
	CACHE_GETCHILD_ACCESS_COST,
	CACHE_GET_DEV_TYPE,
	CACHE_GET_CHILD_DEV_TYPE,
	CACHE_GET_MTMODE,
	CACHE_COST_HI = 100,
	CACHE_THREAD_CLASS,
	CACHE_ACCESS_SEQ1,
};
enum { MT_NICE_FILTER = 1, MT_MULTI_INSTANCE = 2, MT_SERIALIZED = 3 };

struct AVS_Linkage {};

struct VideoInfo {
	int width = 0, height = 0, num_frames = 0;
	bool rgb32 = true;
	bool IsRGB32() const { return rgb32; }
};

class Device {
public:
	int GetType() const { return DEV_TYPE_CPU; }
};

class VideoFrame {
public:
	VideoFrame(int rowBytes, int height) : m_Pitch(rowBytes), m_Data(static_cast<size_t>(rowBytes) * height, 0) {}
	const BYTE *GetReadPtr() const { return m_Data.data(); }
	BYTE *GetWritePtr() { return m_Data.data(); }
	int GetPitch() const { return m_Pitch; }
	Device GetDevice() const { return Device(); }

private:
	int m_Pitch;
	std::vector<BYTE> m_Data;
};
using PVideoFrame = std::shared_ptr<VideoFrame>;

class IScriptEnvironment;

class IClip {
public:
	virtual ~IClip() {}
	virtual PVideoFrame __stdcall GetFrame(int n, IScriptEnvironment *env) = 0;
	virtual const VideoInfo &__stdcall GetVideoInfo() = 0;
	virtual int __stdcall GetVersion() { return 8; }
	virtual int __stdcall SetCacheHints(int, int) { return 0; }
};
using PClip = std::shared_ptr<IClip>;

class GenericVideoFilter : public IClip {
public:
	explicit GenericVideoFilter(PClip c) : child(std::move(c)), vi(child->GetVideoInfo()) {}
	const VideoInfo &__stdcall GetVideoInfo() override { return vi; }

protected:
	PClip child;
	VideoInfo vi;
};

class AVSValue {
public:
	AVSValue() {}
	AVSValue(IClip *c) : m_Clip(c), m_Kind('c') {}       // NOLINT: the plugin returns `new Filter(...)`
	AVSValue(PClip c) : m_Clip(std::move(c)), m_Kind('c') {}  // NOLINT
	AVSValue(const char *s) : m_Str(s), m_Kind('s') {}    // NOLINT
	AVSValue(int i) : m_Int(i), m_Kind('i') {}            // NOLINT
	explicit AVSValue(std::vector<AVSValue> a) : m_Array(std::move(a)), m_Kind('a') {}
	bool Defined() const { return m_Kind != 0; }
	PClip AsClip() const { return m_Clip; }
	const char *AsString() const { return m_Str.c_str(); }
	int AsInt() const { return m_Int; }
	const AVSValue &operator[](int i) const { return m_Array.at(static_cast<size_t>(i)); }

private:
	PClip m_Clip;
	std::string m_Str;
	int m_Int = 0;
	std::vector<AVSValue> m_Array;
	char m_Kind = 0;
};

class IScriptEnvironment {
public:
	using ApplyFunc = AVSValue(__cdecl *)(AVSValue args, void *user_data, IScriptEnvironment *env);
	virtual ~IScriptEnvironment() {}
	[[noreturn]] void ThrowError(const char *fmt, ...) {
		char buf[1024];
		va_list ap;
		va_start(ap, fmt);
		std::vsnprintf(buf, sizeof(buf), fmt, ap);
		va_end(ap);
		throw std::runtime_error(buf);
	}
	void CheckVersion(int) {}
	PVideoFrame NewVideoFrameP(const VideoInfo &vi, PVideoFrame *) {
		return std::make_shared<VideoFrame>(vi.width * 4, vi.height);
	}
	void AddFunction(const char *name, const char *params, ApplyFunc apply, void *user_data) {
		m_Name = name;
		m_Params = params;
		m_Apply = apply;
		m_UserData = user_data;
	}
	std::string m_Name, m_Params;
	ApplyFunc m_Apply = nullptr;
	void *m_UserData = nullptr;
};
