// Symbol visibility for libJoshUpscale.  The reference generates this header
// with CMake's generate_export_header (core/CMakeLists.txt:37-39); everything
// not marked JOSHUPSCALE_EXPORT is hidden (CXX_VISIBILITY_PRESET hidden).
#pragma once

#if defined(_WIN32)
#  if defined(JoshUpscale_EXPORTS)
#    define JOSHUPSCALE_EXPORT __declspec(dllexport)
#  else
#    define JOSHUPSCALE_EXPORT __declspec(dllimport)
#  endif
#else
#  define JOSHUPSCALE_EXPORT __attribute__((visibility("default")))
#endif
