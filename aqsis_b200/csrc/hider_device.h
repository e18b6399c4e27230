// hider_device.h -- structures shared by the host driver (hider_api.cpp) and the
// sm_100a kernels (hider_kernels.cu).
#ifndef AQSIS_B200_HIDER_DEVICE_H
#define AQSIS_B200_HIDER_DEVICE_H

#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/aqsis_b200_hider.h"

namespace aqh {

#define AQH_DEEP_INLINE 8          /* in-line transparent hit slots per sample (hider_kernels.cu: DEEP_INLINE) */

// Per-grid record in HBM (32 B).
struct GridRec
{
	uint32_t vbase;     // first vertex of the grid in the Ci/Oi/culled arrays
	uint32_t pbase;     // first position of the grid in the P arrays (key-major inside the grid)
	uint32_t nverts;    // (cu+1)*(cv+1)
	uint32_t cu_cv;     // cu | cv << 16
	uint32_t flags;     // AQH_GRID_*
	uint32_t nkeys_koff;// nkeys | key offset << 8  (key offset indexes keyTimes / splitLines)
	float lod0, lod1;
};

// info word stored in P4.w (as float bits) by the project kernel
enum : uint32_t
{
	VINFO_GRID_MASK  = 0x0fffffffu,
	VINFO_OPAQUE     = 1u << 28,   // Oi >= 1 in all channels at this vertex (or no Oi)
	VINFO_MP_VALID   = 1u << 29,   // a micropolygon starts at this vertex (iu<cu, iv<cv, not culled; every vertex of a points grid)
	VINFO_MP_TRIMMED = 1u << 30    // a trim curve crosses the micropolygon: its hits are tested against the curves (MarkTrimmed)
};

// Everything the kernels need about the frame; passed by value (<= 4 KB of kernel params).
struct DevFrame
{
	// options
	int xres, yres;
	int cropX0, cropX1, cropY0, cropY1;
	int xs, ys, n;
	int shiftX, shiftY;            // m_DiscreteShiftX/Y
	float xfwo2, yfwo2;            // ceil(filterwidth)*0.5f
	float clipNear, clipFar;
	float shutterOpen, shutterClose;
	int useDof;
	float dofMult, dofInvFocal, dofScaleX, dofScaleY;
	float zthr[3];
	int depthFilter;               // AQH_DEPTHFILTER_*
	int cullable;                  // 0: display mode has z and the depth filter is max/average -> every hit joins the sample's list (bucketprocessor.cpp:1074-1079)
	int midpointZ;                 // display mode has z and the depth filter is midpoint -> occlZ is the SECOND nearest opaque depth
	float expGain, expGamma;
	int jitter;                    // 0: pattern index is always 0 and ncache == 1
	float camToRaster[16];
	int anyMotion, anyTransparent, anyLod, anyTriangular, anyCamera;
	// sample region and tiling
	int sx0, sy0, sw, sh;          // global sample region [crop-shift, crop+shift)
	int tileW, tileH, ntx, nty;    // tiles cover the sample region
	int nActiveTiles;
	// inputs
	const float* Praw;             // n_pos*3
	const float* Ci;               // n_verts*3 or null
	const float* Oi;               // n_verts*3 or null
	const uint8_t* culled;         // n_verts or null
	const float* Ng;               // n_verts*3 camera-space geometric normals (AQH_GRID_CULL_BACKFACING) or null
	const float* Nn;               // n_verts*3 user normals that decide the facing of Ng, or null
	const float* radius;           // n_pos raster radii (AQH_GRID_POINTS) or null
	const float* aov;              // n_verts*aovFloats arbitrary output variables, vertex-major, or null
	int aovFloats, nch;            // floats per pixel of the channel buffer: 9 + aovFloats
	uint32_t* gridTail;            // per grid: 1 + the last vertex whose Oi is not black (AQH_GRID_CULL_TRANSPARENT), 0 = none
	// CSG (aqh_set_csg_tree): per grid its primitive node; per node type, parent, slot among the parent's children and
	// child count; csgOrder lists the non-primitive nodes children-before-parents (the order ProcessSampleList recurses in)
	int anyCSG, nCsgNodes, nCsgOrder;
	const int32_t* gridCsg;        // nGrids or null
	const int32_t* csgType; const int32_t* csgParent; const int32_t* csgSlot; const int32_t* csgKids; const int32_t* csgOrder;
	int cullTransparentOk;         // limits:zthreshold is not black (micropolygon.cpp:496)
	// trim curves (aqh_set_trim_loops): per grid 1 + its trim set (0 = untrimmed); set s owns loops [trimSetLoop[s], trimSetLoop[s+1])
	// (entry 0 is the empty set), loop l the points [trimLoopPoint[l], trimLoopPoint[l+1]); trimUV: surface parameters per vertex
	int anyTrim;
	int plain;                     // none of: discs, level-of-detail ranges, trim curves, triangular grids, > 2 motion keys, CSG, AOVs, incremental flush / occlusion-only pass, midpoint depth filter: the short kernels (k_hide<..., PLAIN>)
	const int32_t* gridTrim; const int32_t* trimSetLoop; const int32_t* trimLoopPoint; const float2* trimPoints; const float2* trimUV;
	// incremental flushes: per-sample occlusion keys kept in HBM between aqh_flush calls ([row][pixel][sample] of the sample
	// region; null when the frame was never flushed).  A flush hides only the opaque micropolygons at positions >= binFrom
	// against them and writes them back; the final frame starts from them and skips the opaque micropolygons below flushedPos.
	unsigned long long* zKeys; unsigned long long* zKeys2;
	int64_t flushedPos;
	int deferDisplay;              // the filter only writes the float channel buffer; k_finish exposes (deferExpose) and quantises
	int deferExpose;
	const GridRec* grids;
	const uint32_t* chunkGrid;     // grid index of the first position of each 256-position chunk (+1 sentinel)
	const float* keyTimes;         // per grid nkeys floats at key offset
	float4* splitLines;            // per (grid,key): Ax,Ay,Bx,By (written by the project kernel)
	int64_t nPos, nVerts;
	int nGrids;
	// derived geometry
	float4* P4;                    // n_pos: raster x, raster y, camera z, info bits
	float4* CO;                    // n_verts x 2: (Ci.r, Ci.g, Ci.b, Oi.r), (Oi.g, Oi.b, -, -): 16-byte loads for the shading of hits
	// frame tables
	const float2* posTab;          // ncache*n
	const float* val1d;            // ncache*n
	const uint8_t* shufTab;        // ncache*n (n <= 256)
	const uint8_t* shufInvTab;     // ncache*n: the inverse permutations (sample index -> lens cell)
	const uint8_t* patPlanes;      // 5 planes of sw*sh
	const float* filterTab;        // (2*shiftX+1)*(2*shiftY+1)*n
	const float4* dofBounds;       // n: minx, miny, maxx, maxy
	const float* dither;           // n_displays * xres*yres
	// tiles / bins
	const int32_t* tileSlot;       // ntx*nty: index into the active tile list or -1
	const uint32_t* activeTiles;   // nActiveTiles tile ids
	uint32_t* binCount;            // nActiveTiles
	uint32_t* binOffset;           // nActiveTiles+1
	unsigned long long* binEntries; // per tile: (transparent pass << 63 | depthKey(zmin) >> 1 << 32 | position index), sorted ascending
	int sortRun;                   // bins are sorted ascending in runs of this many entries
	int binPartition;              // frames of the static kernel: bin entries carry the pass in their top bit (k_bin<fill>, k_tile_flags)
	uint32_t* tileFlags;           // per active tile: index of the first bin entry of the transparent pass (k_tile_flags), 0xffffffff = not partitioned
	// resolved samples: planes [k][y][chunk][x][planeSC] over the sample region: the samples of a pixel are
	// split in planeChunks chunks of planeSC (<= 64, a multiple of 4) slots, pixel rows padded to planeW pixels
	// with slack, so that one (row, chunk) of a span of pixels is ONE contiguous, 16-byte aligned piece for the
	// filter's bulk copies.  Slots past n hold mask 0.
	float* planes;                 // 7 planes: R G B Or Og Ob Z
	unsigned char* maskPlane;      // per slot one compact word of maskBytes (1, 2 or 4) bytes: bits [0, 2*shiftX] x-tap inclusion,
	                               // the next 2*shiftY+1 bits y-tap inclusion, then one bit "holds a valid hit"
	int maskBytes;
	int64_t planeStride;           // ringRows*planeChunks*planeW*planeSC
	int planeW, planeSC, planeChunks;
	int ringRows;                  // sample rows the planes hold at a time: row y lives at (y - sy0) % ringRows.  The whole
	                               // sample region when the frame fits AqhFrameParams::plane_budget_mb, else a ring over bands
	// tile-partials filter mode: per (tap, value, y, x) partial sums over a pixel's samples;
	// values: 0 gTot, 1 hit count, 2..8 R G B Or Og Ob Z
	int filterMode;                // AQH_FILTER_*
	float* partials;               // ntaps*9 planes of sw*sh
	int ntaps;                     // (2*shiftX+1)*(2*shiftY+1)
	// transparent hit pool, per persistent CTA: AQH_DEEP_INLINE in-line records per sample of the tile (rank-major), then
	// deepCapPerCta overflow records; a hit is A = (depth bits, position index) and B = (u, v, first vertex, cu | shading flags)
	uint2* deepA;
	uint4* deepB;
	uint32_t* deepNext;            // overflow chains: next overflow slot + 1
	uint32_t deepCapPerCta;
	uint32_t* errorFlags;          // bit0: deep pool overflow
	unsigned long long* counters;  // [0] MPs binned, [1] bin entries, [2] deep hits, [3] longest bin
	// output
	float* channels;               // xres*yres*9
	// occlusion feedback (aqh_flush): per image pixel the farthest occlusion depth over the pixel's samples
	// (FLT_MAX while any sample is uncovered); with zOnly the tile stops there -- no resolve, no filter input
	float* occlImage;              // sw*sh (the sample region) or null
	int zOnly;
	const uint8_t* rowOwned;       // yres: 1 when this rank owns the pixel row
	int tune[8];                   // development knobs (environment AQH_TUNE="a,b,..."; 0 = the built-in default): [0] bin entries per grab, opaque pass; [1] deep pass; [2] entries between hierarchical-z refreshes; [4] 1 = the general kernels even for a plain frame; [5] factor on the index window up to which a moving micropolygon is enumerated by (pixel, index)
};

struct DevDisplay
{
	int nChannels;
	int channel[AQH_MAX_DISPLAY_CHANNELS];
	int type;
	int entrySize;
	float qZero, qOne, qMin, qMax, qDither;
	unsigned char* out;            // xres*yres*entrySize
};
struct DevDisplays
{
	int n;
	DevDisplay d[AQH_MAX_DISPLAYS];
};

struct LaunchCfg
{
	int smCount;
	int hideThreads;       // threads per CTA of the hide kernel
	int hideCtas;          // persistent CTAs
	size_t hideSmemBytes;
	int batchMPs;
};

// kernel launchers (hider_kernels.cu); all asynchronous on `st`.
cudaError_t launchProjectCount(const DevFrame& f, int64_t pA, int64_t pB, cudaStream_t st);
cudaError_t launchSplitLines(const DevFrame& f, cudaStream_t st);
cudaError_t launchBinScan(const DevFrame& f, cudaStream_t st);
cudaError_t launchBinFill(const DevFrame& f, int64_t pA, int64_t pB, cudaStream_t st);
cudaError_t launchBinCount(const DevFrame& f, int64_t pA, int64_t pB, cudaStream_t st);
cudaError_t launchProject(const DevFrame& f, int64_t pA, int64_t pB, cudaStream_t st);
cudaError_t launchFillKeys(unsigned long long* keys, size_t n, cudaStream_t st);      // every sample "empty" (occlZ = FLT_MAX, no hit)
cudaError_t launchHide(const DevFrame& f, const LaunchCfg& cfg, uint32_t slotBeg, uint32_t slotEnd, uint32_t* cursor, cudaStream_t st);
cudaError_t launchFilter(const DevFrame& f, const DevDisplays& disp, const float* hostFilterTab, int yBeg, int yEnd, bool uploadTable, cudaStream_t st);
cudaError_t launchFinish(const DevFrame& f, const DevDisplays& disp, int expose, cudaStream_t st);   // expose and/or quantise the channel buffer
int filterLaunchCount(const DevFrame& f);     // kernel launches of one launchFilter call
cudaError_t hideKernelConfig(const DevFrame& f, int smCount, LaunchCfg& cfg);
int kernelsArchOk();   // 1 when the loaded kernel image can run on the current device

} // namespace aqh
#endif
