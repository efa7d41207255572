// dtr_kernels.h -- parameter blocks and launchers of the sm_100a kernels (dtr_kernels.cu).
#pragma once
#include <cuda_runtime_api.h>

#include "dtr_records.h"

namespace dtr
{

constexpr int SETUP_THREADS = 128;

struct SetupParams
{
	const DrawItem *items;
	const uint32_t *blockItem; // [ceil(numPrims / SETUP_THREADS)] item of each CTA's first primitive
	int             numItems;
	uint32_t        numPrims;
	PrimRecord     *prims;
	PrimBounds     *bounds;
	uint32_t       *tileCount; // [numFrames * bandTiles]
	uint32_t       *coarseCount; // [numFrames * coarseBins * coarseSegs] when two-level binning is on
	const FrameState *frames;
	const TexDesc  *textures; // texture table: pointer and size are copied into textured records
	Geometry        g;
};

constexpr int SCAN_THREADS = 1024;
constexpr int SCAN_CHUNK   = SCAN_THREADS * 4; // counts scanned per CTA

struct ScanParams
{
	const uint32_t     *counts0;  // per-tile counts [n0]
	uint32_t           *offsets0; // [n0 + 1]
	uint32_t            n0;
	const uint32_t     *counts1;  // coarse counts [n1] (two-level binning), n1 may be 0
	uint32_t           *offsets1; // [n1 + 1]
	uint32_t            n1;
	uint32_t           *order;    // [n0] raster work order: busy tiles first, untouched tiles last
	unsigned long long *status;   // [chunks0 + chunks1] look-back words, zeroed before the launch
	unsigned long long *totals;   // [0] list total, [1] coarse list total
	uint32_t           *workCounter;
	uint32_t           *numBusy;  // number of tiles with primitives
	uint32_t            chunks0;  // filled by launch_scan
};

inline uint32_t scan_status_words(uint32_t n0, uint32_t n1)
{
	return (n0 + SCAN_CHUNK - 1) / SCAN_CHUNK + 1 + (n1 + SCAN_CHUNK - 1) / SCAN_CHUNK;
}

struct BinParams
{
	const PrimBounds *bounds;
	const FrameState *frames;
	const uint32_t   *tileCount;
	const uint32_t   *tileOffset;
	uint32_t         *lists;
	uint32_t          listCapacity;
	// two-level: coarse lists (offsets have one extra trailing entry = total)
	const uint32_t   *coarseOffset;
	uint32_t         *coarseLists;
	uint32_t          coarseCapacity;
	Geometry          g;
};

struct RasterParams
{
	uint32_t           *color; // [F][H][W]
	float              *depth;
	const FrameState   *frames;
	const PrimRecord   *prims;
	const PrimBounds   *bounds;
	const uint32_t     *tileCount;
	const uint32_t     *tileOffset;
	const uint32_t     *order; // work order written by scan_kernel
	const uint32_t     *lists;
	const TexDesc      *textures;
	unsigned long long *setPixels;
	uint32_t           *workCounter;    // zeroed by scan_kernel; items handed out by atomicAdd
	const uint32_t     *numBusy;        // written by scan_kernel
	uint32_t            numTiles;       // filled by launch_raster
	uint32_t            smallTilesMin;  // filled by launch_raster: busy tiles rasterised as fine-grained items
	Geometry            g;
};

void launch_setup(const SetupParams &P, cudaStream_t s);
// Scans tile counts (-> total[0]) and, when nCoarse > 0, coarse counts (-> total[1]); offsets arrays
// get one extra trailing entry holding the total.  Also zeroes the raster work counter.
void launch_scan(const ScanParams &P, cudaStream_t s);
void launch_bin_coarse(const BinParams &P, cudaStream_t s);
void launch_selftest_sqrt(unsigned long long *mismatches, cudaStream_t s);
void launch_bin(const BinParams &P, cudaStream_t s);
void launch_raster(const RasterParams &P, cudaStream_t s);

} // namespace dtr
