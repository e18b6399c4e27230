// hider_internal.h -- the hider object behind the C ABI, shared by the host-side translation units
// (hider_api.cpp: frames; hider_shard.cpp: sharding over ranks and the NCCL gather).
#ifndef AQSIS_B200_HIDER_INTERNAL_H
#define AQSIS_B200_HIDER_INTERNAL_H

#include "hider_device.h"
#include "host_sampling.h"

#include <algorithm>
#include <cstring>
#include <string>
#include <thread>
#include <utility>
#include <vector>

namespace aqh {

struct DevBuf
{
	void* p = nullptr;
	size_t cap = 0;
	cudaError_t reserve(size_t bytes)
	{
		if(bytes <= cap) return cudaSuccess;
		if(p) cudaFree(p);
		p = nullptr; cap = 0;
		size_t want = bytes + bytes/8 + 256;
		cudaError_t e = cudaMalloc(&p, want);
		if(e != cudaSuccess) { want = bytes; e = cudaMalloc(&p, want); }
		if(e == cudaSuccess) cap = want;
		return e;
	}
	// grow, keeping the first `keep` bytes (device-to-device copy on `st`; the old block is freed after the copy is queued:
	// cudaFree synchronises)
	cudaError_t reserveKeep(size_t bytes, size_t keep, cudaStream_t st)
	{
		if(bytes <= cap) return cudaSuccess;
		void* q = nullptr;
		size_t want = bytes + bytes/4 + 256;
		cudaError_t e = cudaMalloc(&q, want);
		if(e != cudaSuccess) { want = bytes; e = cudaMalloc(&q, want); }
		if(e != cudaSuccess) return e;
		if(p && keep) e = cudaMemcpyAsync(q, p, keep < cap ? keep : cap, cudaMemcpyDeviceToDevice, st);
		if(e == cudaSuccess && p) { cudaStreamSynchronize(st); cudaFree(p); }
		if(e != cudaSuccess) { cudaFree(q); return e; }
		p = q; cap = want;
		return cudaSuccess;
	}
	void release() { if(p) cudaFree(p); p = nullptr; cap = 0; }
	template<class T> T* as() const { return static_cast<T*>(p); }
};

struct PinnedBuf
{
	void* p = nullptr;
	size_t cap = 0;
	bool reserve(size_t bytes, bool keep = false, size_t used = 0)
	{
		if(bytes <= cap) return true;
		size_t want = std::max(bytes, cap*2);
		void* q = nullptr;
		if(cudaHostAlloc(&q, want, cudaHostAllocDefault) != cudaSuccess) return false;
		if(keep && p && used) std::memcpy(q, p, used);
		if(p) cudaFreeHost(p);
		p = q; cap = want;
		return true;
	}
	void release() { if(p) cudaFreeHost(p); p = nullptr; cap = 0; }
	template<class T> T* as() const { return static_cast<T*>(p); }
};

// A growable array in pinned host memory: what is uploaded from it can be copied asynchronously.
template<class T> struct PinnedVec
{
	PinnedBuf b;
	size_t n = 0;
	bool ensure(size_t m)
	{
		if(m*sizeof(T) <= b.cap) return true;
		return b.reserve(std::max(m*sizeof(T), size_t(1) << 16), true, n*sizeof(T));
	}
	bool push(const T& v) { if(!ensure(n + 1)) return false; b.as<T>()[n++] = v; return true; }
	T* data() const { return b.as<T>(); }
	void clear() { n = 0; }
	void release() { b.release(); n = 0; }
};

struct Segment
{
	int64_t firstGrid = 0, nGrids = 0, nVerts = 0, nPos = 0;
	const float* P = nullptr; const float* Ci = nullptr; const float* Oi = nullptr; const uint8_t* culled = nullptr;
	// rarely used per-vertex / per-position arrays (block route: the caller's pointers; staged route: null, see AqhHider::hx*)
	const float* aov = nullptr; const float* Ng = nullptr; const float* N = nullptr; const float* radius = nullptr;
	const float* trimUV = nullptr;
	int memorySpace = 0;     // 0 host, 1 device
	bool staged = false;     // lives in the hider's own pinned staging (pointers are offsets to fix up)
	bool closed = false;     // a flush consumed it: aqh_add_grid opens a new staged run
};

} // namespace aqh

struct AqhHider
{
	using DevBuf = aqh::DevBuf; using PinnedBuf = aqh::PinnedBuf; using Segment = aqh::Segment;
	template<class T> using PinnedVec = aqh::PinnedVec<T>;
	using GridRec = aqh::GridRec; using ReplayLayout = aqh::ReplayLayout; using SamplerTables = aqh::SamplerTables;
	int device = 0;
	int smCount = 0;
	cudaStream_t stream = nullptr;
	bool ownStream = false;
	std::string lastError;
	bool inFrame = false, rendered = false;
	AqhFrameParams params{};
	ReplayLayout layout{};
	// frame tables (host) and the key they were built for
	std::string tableKey, maskLayoutKey;
	SamplerTables tables;
	std::vector<uint8_t> patPlanes;
	std::vector<float> dither, filterTab, dofBounds;
	std::vector<uint8_t> shuf8;
	bool tablesUploaded = false;
	std::thread tablesJob;         // builds the frame tables while the caller submits grids (hider_api.cpp: buildTables / joinTables)
	double prepareMs = 0.0;
	// grids of the frame (host tables)
	std::vector<int32_t> gcu, gcv, gnkeys;
	std::vector<uint32_t> gflags;
	std::vector<float> glod, gkeyTimes;
	std::vector<int32_t> gcsg;                      // per grid: its primitive node in the CSG tree or -1
	std::vector<int32_t> gtrim;                     // per grid: 1 + its trim set, 0 = untrimmed
	std::vector<int32_t> trimSetLoop, trimLoopPoint; std::vector<float> trimPoints;   // aqh_set_trim_loops
	bool anyTrim = false, anyTrimUV = false;
	// CSG tree of the frame (aqh_set_csg_tree): type, parent, slot among the parent's children, child count per node,
	// and the non-primitive nodes children-before-parents
	std::vector<int32_t> csgType, csgParent, csgSlot, csgKids, csgOrder;
	// rarely used arrays of grids handed over one by one (aqh_add_grid): frame-global host copies, indexed like the
	// device arrays (vertex / position offsets), grown on first use
	std::vector<float> hxAov, hxNg, hxN, hxRadius, hxTrimUV;
	bool anyAov = false, anyNg = false, anyN = false, anyRadius = false, anyCSG = false, anyPoints = false;
	int maxKeysG = 1;               // largest key count of the frame's grids
	int aovFloats = 0;
	std::vector<Segment> segments;
	// device-side grid table and the 256-position chunk index, built as the grids are submitted
	PinnedVec<GridRec> recs;
	PinnedVec<uint32_t> chunk;
	uint64_t recVb = 0, recPb = 0, recKo = 0;
	bool anyMotionG = false, anyLodG = false, anyTriG = false, anyCamG = false;
	int64_t nVerts = 0, nPos = 0;
	bool anyCi = false, anyOi = false, anyCulled = false, allCi = true, allOi = true;
	// pinned staging for aqh_add_grid
	PinnedBuf stP, stCi, stOi, stCulled;
	size_t stPUsed = 0, stVUsed = 0;   // floats*3 units: positions / vertices staged
	// device buffers
	DevBuf dPraw, dCi, dOi, dCulled, dP4, dCO, dGrids, dChunk, dKeyTimes, dSplit;
	DevBuf dPosTab, dVal1d, dShuf, dPat, dFilt, dDofB, dDither;
	DevBuf dTileSlot, dActive, dBinCount, dBinOffset, dBinEntries, dMisc, dTileFlags;
	DevBuf dPlanes, dMask, dPartials, dDeepA, dDeepB, dDeepUV, dChannels, dRowOwned;
	DevBuf dDisplay[AQH_MAX_DISPLAYS];
	DevBuf dOccl, dBandCursor;
	DevBuf dAov, dNg, dNn, dRadius, dGridTail, dGridCsg, dCsgTab, dTrimUV, dGridTrim, dTrimTab;
	// incremental flushes (aqh_flush): what is already on the device and projected, and the per-sample occlusion keys kept
	// in HBM between flushes -- (depth key << 32 | position of the winning opaque hit) per sample of the sample region,
	// [row][pixel][sample]; zKeys2: the nearest hit when the midpoint depth filter keeps the second nearest depth as occlZ
	DevBuf dZKeys, dZKeys2;
	int64_t upPos = 0, upVerts = 0, upGrids = 0, flushedPos = 0;
	size_t upSegs = 0;
	bool haveZ = false, sawTransparent = false;
	int64_t flushBinEntries = 0, nFlushes = 0;
	std::vector<cudaEvent_t> bandEv;    // one event after every hide / filter launch of a frame
	PinnedBuf hOccl;
	bool haveOccl = false;
	// tiling
	int tileW = 0, tileH = 0, ntx = 0, nty = 0;
	std::vector<int32_t> tileSlot;
	std::vector<uint32_t> activeTiles;
	std::vector<uint8_t> rowOwned;
	std::vector<std::pair<int,int>> strips;
	std::string stripKey, hostStripKey;   // image geometry + row ownership of the device images / of what the host images hold
	// outputs (host)
	PinnedBuf hChannels;
	PinnedBuf hDisplay[AQH_MAX_DISPLAYS];
	int dispType[AQH_MAX_DISPLAYS] = {0}, dispEntry[AQH_MAX_DISPLAYS] = {0};
	bool haveHostImage = false;
	cudaEvent_t ev[8] = {nullptr};
	// pipelined upload: host grids travel in chunks on copyStream while the main stream projects/bins the previous chunk
	cudaStream_t copyStream = nullptr;
	std::vector<cudaEvent_t> chunkEv;
	std::vector<cudaEvent_t> readyEv;   // streamed delivery: a band's finished rows have arrived in the host images
	int deliveredRows = -1;             // bucket rows already handed to the display callbacks (-1: none, not even started)
	int deliveredBuckets = 0;
	std::vector<unsigned char> bucketScratch;
	AqhFrameStats stats{};
	int nChannels = AQH_NUM_CHANNELS;   // floats per pixel of the channel buffer: 9 + the frame's AOV floats
	// NCCL communicator for the gather of the strips (hider_shard.cpp)
	void* comm = nullptr;
	int commRank = 0, commWorld = 1;
	bool gatherPending = false;         // ev[4]..ev[5] bracket a gather whose time has not been read yet

	int fail(int status, const std::string& msg) { lastError = msg; return status; }
	int cudaFail(cudaError_t e, const char* what)
	{
		lastError = std::string(what) + ": " + cudaGetErrorString(e);
		return (e == cudaErrorMemoryAllocation) ? AQH_ERR_NO_MEMORY : AQH_ERR_CUDA;
	}
};


namespace aqh {
/// The pixel-row strips `rank` of p.world_size owns (see AqhFrameParams::strip_rows).
void computeStrips(const AqhFrameParams& p, int rank, std::vector<std::pair<int,int>>& strips);
}

#endif
