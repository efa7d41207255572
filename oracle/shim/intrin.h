/* TEST INFRASTRUCTURE ONLY. MSVC's <intrin.h> is the single Windows-only header
 * reachable from the reference render TU (DTRendererPlatform.h:5,
 * DTRendererRender.cpp:10); on Linux the same intrinsics live here. */
#pragma once
#include <x86intrin.h>
