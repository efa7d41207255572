// dtr_kernels.h -- parameter blocks and launchers of the sm_100a kernels (dtr_kernels.cu).
#pragma once
#include <cuda_runtime_api.h>

#include "dtr_records.h"

namespace dtr
{

struct SetupParams
{
	const DrawItem *items;
	int             numItems;
	uint32_t        numPrims;
	PrimRecord     *prims;
	PrimBounds     *bounds;
	uint32_t       *tileCount; // [numFrames * bandTiles]
	uint32_t       *coarseCount; // [numFrames * coarseBins * coarseSegs] when two-level binning is on
	const FrameState *frames;
	Geometry        g;
};

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
	const uint32_t     *lists;
	const TexDesc      *textures;
	unsigned long long *setPixels;
	uint32_t           *workCounter;    // zeroed by scan_kernel; items handed out by atomicAdd
	uint32_t            numItems;       // filled by launch_raster
	uint32_t            regionsPerItem; // 8 (a whole tile per pull) or 1
	Geometry            g;
};

void launch_setup(const SetupParams &P, cudaStream_t s);
// Scans tile counts (-> total[0]) and, when nCoarse > 0, coarse counts (-> total[1]); offsets arrays
// get one extra trailing entry holding the total.  Also zeroes the raster work counter.
void launch_scan(const uint32_t *counts, uint32_t *offsets, uint32_t n, const uint32_t *coarseCounts,
                 uint32_t *coarseOffsets, uint32_t nCoarse, unsigned long long *totals, uint32_t *workCounter,
                 cudaStream_t s);
void launch_bin_coarse(const BinParams &P, cudaStream_t s);
void launch_selftest_sqrt(unsigned long long *mismatches, cudaStream_t s);
void launch_bin(const BinParams &P, cudaStream_t s);
void launch_raster(const RasterParams &P, cudaStream_t s);

} // namespace dtr
