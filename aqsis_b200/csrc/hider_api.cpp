// hider_api.cpp -- host driver behind the C ABI of include/aqsis_b200_hider.h.
//
// Owns device memory and streams, replays the renderer-global random stream into per-pixel
// pattern planes, lays the shaded grids out in HBM and launches the sm_100a kernels of
// hider_kernels.cu.  There is deliberately no CPU implementation of the path in here: without
// a usable device every entry point that would compute returns AQH_ERR_NO_DEVICE.
#include "hider_internal.h"

#include <algorithm>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

using namespace aqh;

namespace {

// The span filter reads its weights from ONE __constant__ table per device (hider_kernels.cu): hiders that share
// a device and render from different threads take turns from the table upload to the end of their frame.
std::mutex g_filterTableMutex[64];

double nowMs()
{
	return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int typeSize(int type)
{
	switch(type)
	{
		case AQH_FLOAT32: case AQH_UNSIGNED32: case AQH_SIGNED32: return 4;
		case AQH_UNSIGNED16: case AQH_SIGNED16: return 2;
		default: return 1;
	}
}
// selectDataFormat, ddmanager.cpp:249-283
int selectDataFormat(float oneVal, float minVal, float maxVal)
{
	if(oneVal == 0) return AQH_FLOAT32;
	if(minVal >= 0)
	{
		if(maxVal <= 255.0f) return AQH_UNSIGNED8;
		if(maxVal <= 65535.0f) return AQH_UNSIGNED16;
		return AQH_UNSIGNED32;
	}
	if(minVal >= -128.0f && maxVal <= 127.0f) return AQH_SIGNED8;
	if(minVal >= -32768.0f && maxVal <= 32767.0f) return AQH_SIGNED16;
	return AQH_SIGNED32;
}

} // namespace

#define CU(call, what) do { cudaError_t e__ = (call); if(e__ != cudaSuccess) return h->cudaFail(e__, what); } while(0)

namespace {

void resetFrameGrids(AqhHider* h)
{
	h->gcu.clear(); h->gcv.clear(); h->gnkeys.clear(); h->gflags.clear(); h->glod.clear(); h->gkeyTimes.clear();
	h->segments.clear();
	h->recs.clear(); h->chunk.clear(); h->recVb = h->recPb = h->recKo = 0;
	h->anyMotionG = h->anyLodG = h->anyTriG = h->anyCamG = false; h->maxKeysG = 1;
	h->nVerts = h->nPos = 0;
	h->anyCi = h->anyOi = h->anyCulled = false; h->allCi = h->allOi = true;
	h->gcsg.clear(); h->csgType.clear(); h->csgParent.clear(); h->csgSlot.clear(); h->csgKids.clear(); h->csgOrder.clear();
	h->hxAov.clear(); h->hxNg.clear(); h->hxN.clear(); h->hxRadius.clear(); h->hxTrimUV.clear();
	h->gtrim.clear(); h->trimSetLoop.clear(); h->trimLoopPoint.clear(); h->trimPoints.clear(); h->anyTrim = h->anyTrimUV = false;
	h->anyAov = h->anyNg = h->anyN = h->anyRadius = h->anyCSG = h->anyPoints = false;
	h->upPos = h->upVerts = h->upGrids = h->flushedPos = 0; h->upSegs = 0;
	h->haveZ = h->sawTransparent = false; h->flushBinEntries = h->nFlushes = 0;
	h->stPUsed = h->stVUsed = 0;
}

int validateParams(AqhHider* h, const AqhFrameParams& p)
{
	if(p.abi_version != AQH_ABI_VERSION) return h->fail(AQH_ERR_BAD_PARAMS, "abi_version mismatch");
	if(p.xres <= 0 || p.yres <= 0) return h->fail(AQH_ERR_BAD_PARAMS, "resolution must be positive");
	if(p.crop_xmin < 0 || p.crop_ymin < 0 || p.crop_xmax > p.xres || p.crop_ymax > p.yres ||
	   p.crop_xmax <= p.crop_xmin || p.crop_ymax <= p.crop_ymin)
		return h->fail(AQH_ERR_BAD_PARAMS, "crop window must be a non-empty sub-rectangle of the image");
	if(p.xsamples < 1 || p.ysamples < 1 || p.xsamples > 255 || p.ysamples > 255 || p.xsamples*p.ysamples > 256)
		return h->fail(AQH_ERR_BAD_PARAMS, "PixelSamples must be >= 1, at most 255 per axis and 256 samples per pixel");
	if(!(p.filter_xwidth > 0.f) || !(p.filter_ywidth > 0.f) || p.filter_xwidth >= 16.f || p.filter_ywidth >= 16.f)
		return h->fail(AQH_ERR_BAD_PARAMS, "filter widths must be in (0,16)");
	if(p.bucket_xsize < 1 || p.bucket_ysize < 1) return h->fail(AQH_ERR_BAD_PARAMS, "bucket size must be positive");
	if(p.n_displays < 0 || p.n_displays > AQH_MAX_DISPLAYS) return h->fail(AQH_ERR_BAD_PARAMS, "too many displays");
	int aovFloats = 0;
	if(p.n_aovs < 0 || p.n_aovs > AQH_MAX_AOVS) return h->fail(AQH_ERR_BAD_PARAMS, "too many arbitrary output variables");
	for(int a = 0; a < p.n_aovs; ++a)
	{
		if(p.aov[a].n_floats != 1 && p.aov[a].n_floats != 3 && p.aov[a].n_floats != 16)
			return h->fail(AQH_ERR_BAD_PARAMS, "an arbitrary output variable has 1, 3 or 16 floats");
		aovFloats += p.aov[a].n_floats;
	}
	if(aovFloats > AQH_MAX_AOV_FLOATS) return h->fail(AQH_ERR_BAD_PARAMS, "too many floats of arbitrary output variables");
	for(int d = 0; d < p.n_displays; ++d)
	{
		const AqhDisplayDesc& dd = p.display[d];
		if(dd.n_channels < 1 || dd.n_channels > AQH_MAX_DISPLAY_CHANNELS) return h->fail(AQH_ERR_BAD_PARAMS, "display channel count");
		for(int c = 0; c < dd.n_channels; ++c)
			if(dd.channel[c] < 0 || dd.channel[c] >= AQH_NUM_CHANNELS + aovFloats) return h->fail(AQH_ERR_BAD_PARAMS, "display channel index");
		if(dd.type < 0 || dd.type > AQH_SIGNED8) return h->fail(AQH_ERR_BAD_PARAMS, "display data type");
	}
	if(p.depth_filter < AQH_DEPTHFILTER_MIN || p.depth_filter > AQH_DEPTHFILTER_AVERAGE)
		return h->fail(AQH_ERR_BAD_PARAMS, "depth_filter");
	if(p.filter_mode != AQH_FILTER_TILE_PARTIALS && p.filter_mode != AQH_FILTER_REFERENCE_ORDER)
		return h->fail(AQH_ERR_BAD_PARAMS, "filter_mode");
	if(p.filter_xwidth >= 16.f || p.filter_ywidth >= 16.f)
		return h->fail(AQH_ERR_BAD_PARAMS, "filter widths must be below 16");
	if(p.world_size < 0 || p.world_size > AQH_MAX_RANKS || (p.world_size > 0 && (p.rank < 0 || p.rank >= p.world_size)))
		return h->fail(AQH_ERR_BAD_PARAMS, "rank/world_size");
	if(p.strip_rows < -2) return h->fail(AQH_ERR_BAD_PARAMS, "strip_rows");
	if(p.use_dof && !(p.dof_one_over_focal_distance != 0.f))
		return h->fail(AQH_ERR_BAD_PARAMS, "depth of field needs a finite focal distance");
	return AQH_OK;
}

void chooseTile(const AqhFrameParams& p, bool mbdofHint, int& tw, int& th)
{
	const int n = p.xsamples*p.ysamples;
	// Static frames: tiles of 2048 samples (256-thread CTAs, four per SM) unless that leaves fewer than 32 pixels per tile --
	// then 4096 samples (512-thread CTAs, two per SM).  Measured: config 2 (64 samples per pixel) hides 6.5 % faster on the
	// small tiles, config 4 (256 samples per pixel: 8 instead of 16 pixels per tile) 13 % slower (profiles/README.md).
	int staticTarget = (32*n <= 2048) ? 2048 : 4096;
	if(const char* e = std::getenv("AQH_ST_TILE")) staticTarget = std::atoi(e);      // (development: A/B of the tile size)
	const int target = mbdofHint ? 2048 : staticTarget;
	tw = 16; th = 16;
	while(tw*p.xsamples > 255 && tw > 1) tw >>= 1;
	while(th*p.ysamples > 255 && th > 1) th >>= 1;
	while(tw*th*n > target && (tw > 1 || th > 1))
	{
		if(th >= tw && th > 1) th >>= 1; else tw >>= 1;
	}
}

// Build (or reuse) the frame tables: jitter patterns, per-pixel pattern planes, dither, filter
// weights, lens-cell bounds.  Everything here depends on the options only, not on the geometry.  The replay of the
// renderer's random stream over the frame (the expensive part: one draw sequence per pixel in bucket order) runs on
// a host thread of its own from aqh_begin_frame on -- while the caller submits its grids -- and is joined by the first
// thing that needs the tables (joinTables).
void buildTablesJob(AqhHider* h, std::string key)
{
	const double t0 = nowMs();
	const AqhFrameParams& p = h->params;
	const ReplayLayout& L = h->layout;
	// RiWorldBegin reseeds (ri.cpp:660); RenderImage always constructs the jittered sampler
	// (imagebuffer.cpp:694), which consumes the stream and reseeds with 19.
	Random rng(p.rng_seed);
	rng.discard(p.rng_predraws);
	SamplerTables jit;
	buildJitterTables(rng, p.xsamples, p.ysamples, jit);
	if(p.jitter) h->tables = std::move(jit);
	else buildGridTables(p.xsamples, p.ysamples, h->tables);
	const size_t plane = size_t(L.sw)*L.sh;
	h->patPlanes.assign(5*plane, 0);
	h->dither.assign(size_t(std::max(1, p.n_displays))*p.xres*p.yres, 0.f);
	replayFrame(p, L, rng, p.jitter != 0, h->patPlanes.data(), h->dither.data());
	buildFilterTable(p, h->filterTab);
	buildDofBounds(p.xsamples, p.ysamples, h->dofBounds);
	// the shuffle patterns as bytes, followed by their inverses (sample index -> lens cell)
	{
		const size_t N = h->tables.shuffled.size(), nS = size_t(p.xsamples)*p.ysamples;
		h->shuf8.assign(2*N, 0);
		for(size_t i = 0; i < N; ++i)
		{
			h->shuf8[i] = static_cast<uint8_t>(h->tables.shuffled[i]);
			h->shuf8[N + (i/nS)*nS + size_t(h->tables.shuffled[i])] = static_cast<uint8_t>(i % nS);
		}
	}
	h->tableKey = key;
	h->tablesUploaded = false;
	h->prepareMs = nowMs() - t0;
}
void joinTables(AqhHider* h)
{
	if(h->tablesJob.joinable())
	{
		h->tablesJob.join();
		h->stats.prepare_ms = h->prepareMs;
	}
}
int buildTables(AqhHider* h)
{
	joinTables(h);
	const AqhFrameParams& p = h->params;
	char key[512];
	std::snprintf(key, sizeof key, "%d %d|%d %d %d %d|%d %d|%a %a %p|%d %d|%d|%u %u|%d",
	              p.xres, p.yres, p.crop_xmin, p.crop_xmax, p.crop_ymin, p.crop_ymax, p.xsamples, p.ysamples,
	              p.filter_xwidth, p.filter_ywidth, (void*)p.filter_func, p.bucket_xsize, p.bucket_ysize, p.jitter,
	              p.rng_seed, p.rng_predraws, p.n_displays);
	if(h->tableKey == key && h->tablesUploaded) return AQH_OK;
	h->layout = replayLayout(p);
	h->tableKey.clear();
	try { h->tablesJob = std::thread(buildTablesJob, h, std::string(key)); }
	catch(...) { buildTablesJob(h, key); h->stats.prepare_ms = h->prepareMs; }       // no thread to be had: build in place
	return AQH_OK;
}

int uploadTables(AqhHider* h)
{
	joinTables(h);
	if(h->tablesUploaded) return AQH_OK;
	cudaStream_t st = h->stream;
	struct Up { DevBuf* b; const void* src; size_t bytes; };
	const Up ups[] = {
		{&h->dPosTab, h->tables.pos.data(), h->tables.pos.size()*4},
		{&h->dVal1d, h->tables.val1d.data(), h->tables.val1d.size()*4},
		{&h->dShuf, h->shuf8.data(), h->shuf8.size()},
		{&h->dPat, h->patPlanes.data(), h->patPlanes.size()},
		{&h->dFilt, h->filterTab.data(), h->filterTab.size()*4},
		{&h->dDofB, h->dofBounds.data(), h->dofBounds.size()*4},
		{&h->dDither, h->dither.data(), h->dither.size()*4},
	};
	for(const Up& u : ups)
	{
		CU(u.b->reserve(std::max<size_t>(u.bytes, 16)), "cudaMalloc(frame tables)");
		CU(cudaMemcpyAsync(u.b->p, u.src, u.bytes, cudaMemcpyHostToDevice, st), "cudaMemcpyAsync(frame tables)");
		h->stats.h2d_bytes += (int64_t)u.bytes;
	}
	CU(cudaStreamSynchronize(st), "cudaStreamSynchronize(frame tables)");
	h->tablesUploaded = true;
	return AQH_OK;
}

// Strips of pixel rows dealt round-robin to the ranks.
//   strip_rows > 0: fixed strip height (rounded down to a multiple of 16 rows, at least 16);
//   strip_rows <= 0 (default): balanced -- world*k strips of near-equal height with
//   k = max(1, rows / (64*world)), so that every rank owns the same number of strips (a fixed
//   64-row strip leaves 1080 rows as 17 strips: 3 on one of 8 ranks, 2 on the others).
} // namespace
namespace aqh {
static int stripRows(const AqhFrameParams& p)
{
	return std::max(16, (p.strip_rows/16)*16);
}
void computeStrips(const AqhFrameParams& p, int rank, std::vector<std::pair<int,int>>& strips)
{
	const int world = std::max(1, p.world_size);
	const int me = world > 1 ? rank : 0;
	strips.clear();
	if(p.strip_rows > 0)
	{
		const int strip = stripRows(p);
		int si = 0;
		for(int y0 = p.crop_ymin; y0 < p.crop_ymax; y0 += strip, ++si)
			if(si % world == me)
				strips.push_back(std::make_pair(y0, std::min(y0 + strip, p.crop_ymax)));
		return;
	}
	const int64_t rows = std::max(0, p.crop_ymax - p.crop_ymin);
	if(p.strip_rows == -2 && world > 1)
	{
		// explicit boundaries (aqh_balance_strips), clamped to the crop window and kept monotonic
		int y0 = std::min(std::max(p.strip_bounds[me], p.crop_ymin), p.crop_ymax);
		int y1 = std::min(std::max(p.strip_bounds[me + 1], p.crop_ymin), p.crop_ymax);
		if(me == 0) y0 = p.crop_ymin;
		if(me == world - 1) y1 = p.crop_ymax;
		if(y1 > y0) strips.push_back(std::make_pair(y0, y1));
		return;
	}
	if(p.strip_rows == -1 && world > 1)
	{
		// one contiguous strip per rank; inner boundaries on multiples of 16 rows where the frame is tall enough
		auto bound = [&](int r) -> int {
			if(r <= 0) return p.crop_ymin;
			if(r >= world) return p.crop_ymax;
			int y = p.crop_ymin + (int)(rows*r/world);
			if(rows >= 128*(int64_t)world) y = (y/16)*16; else if(rows >= 16*(int64_t)world) y = (y/4)*4;
			return std::min(std::max(y, p.crop_ymin), p.crop_ymax);
		};
		const int y0 = bound(me), y1 = bound(me + 1);
		if(y1 > y0) strips.push_back(std::make_pair(y0, y1));
		return;
	}
	const int64_t nstrips = (int64_t)world*std::max<int64_t>(1, rows/(64*(int64_t)world));
	for(int64_t si = me; si < nstrips; si += world)
	{
		const int y0 = p.crop_ymin + (int)(si*rows/nstrips), y1 = p.crop_ymin + (int)((si + 1)*rows/nstrips);
		if(y1 > y0) strips.push_back(std::make_pair(y0, y1));
	}
}
} // namespace aqh
namespace {

// Every rank also hides the `shift` rows of halo samples its filter footprint needs.
void buildTiling(AqhHider* h, bool mbdof)
{
	const AqhFrameParams& p = h->params;
	const ReplayLayout& L = h->layout;
	chooseTile(p, mbdof, h->tileW, h->tileH);
	h->ntx = (L.sw + h->tileW - 1)/h->tileW;
	h->nty = (L.sh + h->tileH - 1)/h->tileH;
	h->rowOwned.assign(p.yres, 0);
	computeStrips(p, p.rank, h->strips);
	{
		char key[160];
		std::snprintf(key, sizeof key, "%d %d %d %d %d %d", p.xres, p.yres, p.rank, p.world_size, p.strip_rows, p.n_displays);
		h->stripKey = key;
		for(const auto& st : h->strips) { std::snprintf(key, sizeof key, " %d-%d", st.first, st.second); h->stripKey += key; }
	}
	std::vector<uint8_t> rowNeeded(L.sh, 0);
	for(const auto& s : h->strips)
	{
		for(int y = s.first; y < s.second; ++y) h->rowOwned[y] = 1;
		for(int y = s.first - L.shiftY; y < s.second + L.shiftY; ++y) rowNeeded[y - L.sy0] = 1;
	}
	h->tileSlot.assign(size_t(h->ntx)*h->nty, -1);
	h->activeTiles.clear();
	for(int ty = 0; ty < h->nty; ++ty)
	{
		bool need = false;
		for(int r = ty*h->tileH; r < std::min((ty+1)*h->tileH, L.sh) && !need; ++r) need = rowNeeded[r] != 0;
		if(!need) continue;
		for(int tx = 0; tx < h->ntx; ++tx)
		{
			h->tileSlot[size_t(ty)*h->ntx + tx] = (int32_t)h->activeTiles.size();
			h->activeTiles.push_back(uint32_t(ty*h->ntx + tx));
		}
	}
}

// Everything about one grid that can be wrong, checked BEFORE any state is touched: a rejected grid or block
// leaves the frame exactly as it was (the caller may skip it and go on).
int checkGrid(AqhHider* h, int cu, int cv, int nkeys, uint32_t flags, const float* times, int csgNode, const float* radius, const float* Ng,
              int trimSet = 0, const float* trimUV = nullptr)
{
	if(trimSet != 0)
	{
		if(trimSet < 0 || trimSet + 1 >= (int)h->trimSetLoop.size()) return h->fail(AQH_ERR_BAD_PARAMS, "trimmed grid without a set of the table given to aqh_set_trim_loops");
		if(!trimUV) return h->fail(AQH_ERR_BAD_PARAMS, "trimmed grid without surface parameters (trim_uv)");
		if(flags & AQH_GRID_POINTS) return h->fail(AQH_ERR_BAD_PARAMS, "points are not trimmed");
	}
	if(flags & AQH_GRID_POINTS)
	{
		// CqMicroPolyGridPoints: cu + 1 points, no second dimension
		if(cu < 0 || cu > 65535 || cv != 0) return h->fail(AQH_ERR_BAD_PARAMS, "a points grid has cu + 1 points and cv = 0");
		if(!radius) return h->fail(AQH_ERR_BAD_PARAMS, "points grid without radii");
		if(nkeys != 1) return h->fail(AQH_ERR_UNSUPPORTED, "moving points (CqMicroPolygonMotionPoints) are not supported");
		if(flags & (AQH_GRID_TRIANGULAR | AQH_GRID_USES_CSG)) return h->fail(AQH_ERR_BAD_PARAMS, "points grids are neither triangular nor part of a solid");
	}
	else if(cu < 1 || cv < 1 || cu > 65535 || cv > 65535) return h->fail(AQH_ERR_BAD_PARAMS, "grid resolution out of range");
	if(nkeys < 1 || nkeys > 255) return h->fail(AQH_ERR_BAD_PARAMS, "grid key count out of range");
	if(flags & AQH_GRID_USES_CSG)
	{
		if(csgNode < 0 || csgNode >= (int)h->csgType.size())
			return h->fail(AQH_ERR_BAD_PARAMS, "CSG grid without a node of the tree given to aqh_set_csg_tree");
		if(h->csgType[csgNode] != AQH_CSG_PRIMITIVE) return h->fail(AQH_ERR_BAD_PARAMS, "a grid belongs to a primitive node of the CSG tree");
	}
	if((flags & AQH_GRID_CULL_BACKFACING) && (!Ng || !(flags & AQH_GRID_CAMERA_SPACE)))
		return h->fail(AQH_ERR_BAD_PARAMS, "backface culling needs camera-space P and the geometric normals Ng");
	if(nkeys > 1 && !times) return h->fail(AQH_ERR_BAD_PARAMS, "motion grid without key times");
	return AQH_OK;
}

// Sizes of the per-frame grid tables, to undo a partly appended block when an allocation fails half way.
struct GridTablesMark
{
	size_t nGrids, nKeyTimes, nRecs, nChunk;
	uint64_t recVb, recPb, recKo;
	bool anyMotionG, anyLodG, anyTriG, anyCamG, anyCSG, anyPoints;
	int maxKeysG;
};
GridTablesMark markGridTables(const AqhHider* h)
{
	return GridTablesMark{h->gcu.size(), h->gkeyTimes.size(), h->recs.n, h->chunk.n, h->recVb, h->recPb, h->recKo,
	                      h->anyMotionG, h->anyLodG, h->anyTriG, h->anyCamG, h->anyCSG, h->anyPoints, h->maxKeysG};
}
void rollbackGridTables(AqhHider* h, const GridTablesMark& m)
{
	h->gcu.resize(m.nGrids); h->gcv.resize(m.nGrids); h->gnkeys.resize(m.nGrids); h->gflags.resize(m.nGrids);
	h->glod.resize(2*m.nGrids); h->gkeyTimes.resize(m.nKeyTimes); h->gcsg.resize(m.nGrids); h->gtrim.resize(m.nGrids);
	h->recs.n = m.nRecs; h->chunk.n = m.nChunk;
	h->recVb = m.recVb; h->recPb = m.recPb; h->recKo = m.recKo;
	h->anyMotionG = m.anyMotionG; h->anyLodG = m.anyLodG; h->anyTriG = m.anyTriG; h->anyCamG = m.anyCamG;
	h->anyCSG = m.anyCSG; h->anyPoints = m.anyPoints; h->maxKeysG = m.maxKeysG;
}

int appendGridTables(AqhHider* h, int cu, int cv, int nkeys, uint32_t flags, const float* lod, const float* times, int csgNode, int trimSet = 0)
{
	if(h->recKo + (uint64_t)nkeys >= (1u << 24)) return h->fail(AQH_ERR_BAD_PARAMS, "too many motion keys in one frame");
	h->gcu.push_back(cu); h->gcv.push_back(cv); h->gnkeys.push_back(nkeys); h->gflags.push_back(flags);
	h->gcsg.push_back((flags & AQH_GRID_USES_CSG) ? csgNode : -1);
	h->gtrim.push_back(trimSet); h->anyTrim |= trimSet != 0;
	h->anyCSG |= (flags & AQH_GRID_USES_CSG) != 0; h->anyPoints |= (flags & AQH_GRID_POINTS) != 0;
	h->glod.push_back(lod ? lod[0] : -1.f); h->glod.push_back(lod ? lod[1] : -1.f);
	for(int k = 0; k < nkeys; ++k) h->gkeyTimes.push_back(nkeys > 1 ? times[k] : 0.f);
	// the grid's device record and the chunk index entries of the positions it covers
	GridRec r;
	const uint32_t nv = uint32_t(cu + 1)*uint32_t(cv + 1);
	r.vbase = (uint32_t)h->recVb; r.pbase = (uint32_t)h->recPb; r.nverts = nv;
	r.cu_cv = uint32_t(cu) | (uint32_t(cv) << 16);
	r.flags = flags;
	r.nkeys_koff = uint32_t(nkeys) | (uint32_t(h->recKo) << 8);
	r.lod0 = lod ? lod[0] : -1.f; r.lod1 = lod ? lod[1] : -1.f;
	const uint32_t g = (uint32_t)h->recs.n;
	if(!h->recs.push(r)) return h->fail(AQH_ERR_NO_MEMORY, "cudaHostAlloc(grid table)");
	const uint64_t np = uint64_t(nv)*uint64_t(nkeys);
	const uint64_t c0 = (h->recPb + 255)/256, c1 = (h->recPb + np - 1)/256;      // chunks whose first position lies in this grid
	if(c1 + 1 >= c0 + 1 && c1 >= c0)
	{
		if(!h->chunk.ensure(c1 + 3)) return h->fail(AQH_ERR_NO_MEMORY, "cudaHostAlloc(chunk table)");
		for(uint64_t c = c0; c <= c1; ++c) h->chunk.data()[c] = g;
		h->chunk.n = std::max<size_t>(h->chunk.n, c1 + 1);
	}
	h->anyMotionG |= nkeys > 1; h->anyLodG |= r.lod0 >= 0.f; h->maxKeysG = std::max(h->maxKeysG, nkeys);
	h->anyTriG |= (flags & AQH_GRID_TRIANGULAR) != 0; h->anyCamG |= (flags & AQH_GRID_CAMERA_SPACE) != 0;
	h->recVb += nv; h->recPb += np; h->recKo += (uint64_t)nkeys;
	return AQH_OK;
}

// Everything on the device for the frame.  download: copy the image(s) back to pinned host memory.
// AQH_TRACE=1: host wall-clock checkpoints of one frame on stderr
struct FrameTrace
{
	bool on; double t0, last;
	FrameTrace() : on(std::getenv("AQH_TRACE") != nullptr), t0(nowMs()), last(t0) {}
	void mark(const char* what) { if(!on) return; const double t = nowMs(); std::fprintf(stderr, "[aqh] %-22s +%8.3f ms  (%8.3f)\n", what, t - last, t - t0); last = t; }
};

int deliverBucketRows(AqhHider* h, const AqhCallbacks* cb, int rowEnd);

int renderFrame(AqhHider* h, bool download, bool zOnly = false, const AqhCallbacks* imagerCb = nullptr)
{
	FrameTrace tr;
	if(!h->inFrame) return h->fail(AQH_ERR_STATE, "no frame in progress");
	const AqhFrameParams& p = h->params;
	const ReplayLayout& L = h->layout;
	cudaStream_t st = h->stream;
	AqhFrameStats& S = h->stats;
	const int64_t nGrids = (int64_t)h->gcu.size();
	if(h->nPos >= (int64_t)0xfffffff0u) return h->fail(AQH_ERR_BAD_PARAMS, "more than 2^32 grid positions in one frame");
	if(nGrids > (int64_t)VINFO_GRID_MASK) return h->fail(AQH_ERR_BAD_PARAMS, "more than 2^28 grids in one frame");

	tr.mark("entry");
	const double tUp0 = nowMs();
	S.h2d_bytes = 0;
	int rc = AQH_OK;

	// ---- grid records and chunk index were built at submission (appendGridTables); close the chunk index with
	// two sentinels (k_project reads one entry past the chunk of a position)
	const bool anyMotion = h->anyMotionG, anyLod = h->anyLodG, anyTri = h->anyTriG, anyCam = h->anyCamG;
	const int64_t nChunks = (h->nPos + 255)/256;
	if(!h->chunk.ensure((size_t)nChunks + 2)) return h->fail(AQH_ERR_NO_MEMORY, "cudaHostAlloc(chunk table)");
	h->chunk.data()[nChunks] = h->chunk.data()[nChunks + 1] = (uint32_t)std::max<int64_t>(nGrids - 1, 0);
	const GridRec* recs = h->recs.data();
	const size_t nRecs = h->recs.n, nChunkEntries = (size_t)nChunks + 2;
	tr.mark("grid records");
	const bool mbdof = anyMotion || p.use_dof;
	buildTiling(h, mbdof);
	tr.mark("tiling");
	const int nActive = (int)h->activeTiles.size();

	// ---- incremental flushes: a flush (zOnly) and every later call of the frame only upload and project what was
	// submitted since the previous flush; projected positions (P4), packed shading (CO) and the per-sample occlusion keys
	// persist in HBM in between
	const bool incremental = zOnly || h->haveZ;
	const size_t pos0 = incremental ? (size_t)h->upPos : 0, vert0 = incremental ? (size_t)h->upVerts : 0;
	// ---- device allocations
	const size_t nPos = (size_t)h->nPos, nVerts = (size_t)h->nVerts;
	CU(h->dGrids.reserve(std::max<size_t>(nRecs, 1)*sizeof(GridRec)), "cudaMalloc(grid table)");
	CU(h->dChunk.reserve(nChunkEntries*4), "cudaMalloc(chunk table)");
	CU(h->dKeyTimes.reserve(std::max<size_t>(h->gkeyTimes.size(), 1)*4), "cudaMalloc(key times)");
	CU(h->dSplit.reserve(std::max<size_t>(h->gkeyTimes.size(), 1)*16), "cudaMalloc(split lines)");
	if(incremental)
	{
		CU(h->dP4.reserveKeep(std::max<size_t>(nPos, 1)*16 + 64, pos0*16, st), "cudaMalloc(P4)");
		CU(h->dCO.reserveKeep(std::max<size_t>(nVerts, 1)*32 + 64, vert0*32, st), "cudaMalloc(packed Ci/Oi)");
	}
	else
	{
		CU(h->dP4.reserve(std::max<size_t>(nPos, 1)*16 + 64), "cudaMalloc(P4)");
		CU(h->dCO.reserve(std::max<size_t>(nVerts, 1)*32 + 64), "cudaMalloc(packed Ci/Oi)");
	}
	CU(h->dTileSlot.reserve(std::max<size_t>(h->tileSlot.size(), 1)*4), "cudaMalloc(tile slots)");
	CU(h->dActive.reserve(std::max<size_t>(nActive, 1)*4), "cudaMalloc(active tiles)");
	CU(h->dBinCount.reserve(std::max<size_t>(nActive, 1)*4), "cudaMalloc(bin counts)");
	CU(h->dBinOffset.reserve((size_t(nActive) + 1)*4), "cudaMalloc(bin offsets)");
	CU(h->dTileFlags.reserve(std::max<size_t>(nActive, 1)*4), "cudaMalloc(tile flags)");
	CU(h->dMisc.reserve(256), "cudaMalloc(counters)");
	CU(h->dRowOwned.reserve(p.yres), "cudaMalloc(row ownership)");
	// resolved-sample planes [k][row][chunk][x][planeSC] (hider_device.h).  They hold the whole sample region when that
	// fits AqhFrameParams::plane_budget_mb; otherwise ringRows rows at a time, and the frame is hidden and filtered in
	// bands of tile rows (the reference's own memory bound is one bucket plus its overlap, bucketprocessor.cpp:259-358).
	const int nSamp = p.xsamples*p.ysamples;
	const int planeSC = std::min((nSamp + 3) & ~3, 64);
	const int planeChunks = (nSamp + planeSC - 1)/planeSC;
	const int planeW = ((L.sw + 31) & ~31) + 16;     // the filter stages spans of up to 32+14 pixels starting at multiples of 32
	const int ntaps = (2*L.shiftX+1)*(2*L.shiftY+1);
	const int nPlanes = 7 + h->aovFloats + 1;                                      // R G B Or Og Ob Z, the AOV floats, the mask plane
	const size_t planeRowBytes = size_t(planeW)*planeChunks*planeSC*nPlanes*4;
	int ringRows = L.sh, bandTileRows = h->nty;
	if(p.filter_mode == AQH_FILTER_REFERENCE_ORDER && !zOnly)
	{
		const size_t budget = size_t(p.plane_budget_mb > 0 ? p.plane_budget_mb : 4608) << 20;
		const size_t fit = budget/planeRowBytes;
		if(fit < size_t(L.sh))
		{
			bandTileRows = (int)std::max<int64_t>(1, (int64_t(fit) - 2*L.shiftY)/h->tileH);
			ringRows = std::min(L.sh, bandTileRows*h->tileH + 2*L.shiftY);
		}
	}
	const size_t planeStride = size_t(planeW)*ringRows*planeChunks*planeSC;
	if(zOnly) {}
	else if(p.filter_mode == AQH_FILTER_REFERENCE_ORDER)
	{
		CU(h->dPlanes.reserve(planeStride*nPlanes*4 + 256), "cudaMalloc(sample planes)");
	}
	else
		CU(h->dPartials.reserve(size_t(ntaps)*9*L.sw*L.sh*4), "cudaMalloc(tap partial sums)");
	const int nch = h->nChannels;
	CU(h->dChannels.reserve(size_t(p.xres)*p.yres*nch*4), "cudaMalloc(channel buffer)");
	DevDisplays disp{};
	disp.n = p.n_displays;
	for(int d = 0; d < p.n_displays; ++d)
	{
		const AqhDisplayDesc& dd = p.display[d];
		DevDisplay& o = disp.d[d];
		o.nChannels = dd.n_channels;
		for(int c = 0; c < dd.n_channels; ++c) o.channel[c] = dd.channel[c];
		o.type = dd.type ? dd.type : selectDataFormat(dd.quantize_one, dd.quantize_min, dd.quantize_max);
		o.entrySize = typeSize(o.type)*dd.n_channels;
		o.qZero = dd.quantize_zero; o.qOne = dd.quantize_one; o.qMin = dd.quantize_min; o.qMax = dd.quantize_max;
		o.qDither = dd.quantize_dither;
		CU(h->dDisplay[d].reserve(size_t(p.xres)*p.yres*o.entrySize), "cudaMalloc(display image)");
		o.out = h->dDisplay[d].as<unsigned char>();
		h->dispType[d] = o.type; h->dispEntry[d] = o.entrySize;
	}

	tr.mark("allocations");
	// ---- grids into HBM
	const float* dP = nullptr; const float* dCi = nullptr; const float* dOi = nullptr; const uint8_t* dCulled = nullptr;
	// Pipelined upload (large frames handed over in host memory): the bulk arrays go up in chunks of whole grids
	// on a second stream; each chunk is projected and bin-counted while the next one is on the bus.
	struct UploadChunk { const float* P; const float* Ci; const float* Oi; const uint8_t* culled; int64_t p0, p1, v0, v1; };
	std::vector<UploadChunk> plan;
	// AQH_PIPELINE_MIN_POS: smallest frame (grid positions) that takes this route (default 2^20; tests force 1; a huge value disables it)
	size_t pipeMin = size_t(1) << 20;
	if(const char* e = std::getenv("AQH_PIPELINE_MIN_POS")) pipeMin = (size_t)std::strtoull(e, nullptr, 10);
	bool pipelined = !incremental && h->copyStream != nullptr && nPos >= std::max<size_t>(pipeMin, 1);
	for(const Segment& s : h->segments)
		if(s.memorySpace != 0 || (h->anyCi && !s.Ci) || (h->anyOi && !s.Oi)) pipelined = false;
	const bool zeroCopy = !incremental && h->segments.size() == 1 && h->segments[0].memorySpace == 1;
	if(zeroCopy)
	{
		const Segment& s = h->segments[0];
		dP = s.P; dCi = s.Ci; dOi = s.Oi; dCulled = s.culled;
	}
	else if(nPos)
	{
		CU(h->dPraw.reserve(nPos*12), "cudaMalloc(P)");
		if(h->anyCi) CU(h->dCi.reserve(nVerts*12), "cudaMalloc(Ci)");
		if(h->anyOi) CU(h->dOi.reserve(nVerts*12), "cudaMalloc(Oi)");
		if(h->anyCulled) CU(h->dCulled.reserve(nVerts), "cudaMalloc(culled)");
		if(h->anyCulled && nVerts > vert0) CU(cudaMemsetAsync(h->dCulled.as<uint8_t>() + vert0, 0, nVerts - vert0, st), "cudaMemsetAsync(culled)");
		std::vector<float> ones;
		size_t po = 0, vo = 0, segIndex = 0;
		for(const Segment& s : h->segments)
		{
			if(pipelined) break;
			// the raw arrays are only read by k_project: segments an earlier flush projected need not travel again
			if(incremental && segIndex++ < h->upSegs) { po += s.nPos; vo += s.nVerts; continue; }
			const float* sP = s.P; const float* sCi = s.Ci; const float* sOi = s.Oi; const uint8_t* sCu = s.culled;
			if(s.staged)
			{
				sP = h->stP.as<float>() + (size_t)(uintptr_t)s.P;
				sCi = h->stCi.as<float>() + (size_t)(uintptr_t)s.Ci;
				sOi = h->stOi.as<float>() + (size_t)(uintptr_t)s.Oi;
				sCu = s.culled ? h->stCulled.as<uint8_t>() + ((size_t)(uintptr_t)s.culled - 1) : nullptr;
			}
			const cudaMemcpyKind kind = s.memorySpace ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
			CU(cudaMemcpyAsync(h->dPraw.as<float>() + po*3, sP, size_t(s.nPos)*12, kind, st), "cudaMemcpyAsync(P)");
			if(!s.memorySpace) S.h2d_bytes += s.nPos*12;
			for(int which = 0; which < 2; ++which)
			{
				const bool any = which ? h->anyOi : h->anyCi;
				if(!any) continue;
				const float* src = which ? sOi : sCi;
				float* dst = (which ? h->dOi.as<float>() : h->dCi.as<float>()) + vo*3;
				if(src)
				{
					CU(cudaMemcpyAsync(dst, src, size_t(s.nVerts)*12, kind, st), "cudaMemcpyAsync(Ci/Oi)");
					if(!s.memorySpace) S.h2d_bytes += s.nVerts*12;
				}
				else
				{
					// this run has no Ci/Oi but another one does: materialise the default 1.0
					ones.assign(size_t(s.nVerts)*3, 1.0f);
					CU(cudaMemcpyAsync(dst, ones.data(), ones.size()*4, cudaMemcpyHostToDevice, st), "cudaMemcpyAsync(default Ci/Oi)");
					CU(cudaStreamSynchronize(st), "cudaStreamSynchronize");
				}
			}
			if(h->anyCulled && sCu)
			{
				CU(cudaMemcpyAsync(h->dCulled.as<uint8_t>() + vo, sCu, size_t(s.nVerts), kind, st), "cudaMemcpyAsync(culled)");
				if(!s.memorySpace) S.h2d_bytes += s.nVerts;
			}
			po += s.nPos; vo += s.nVerts;
		}
		if(pipelined)
		{
			const int64_t target = std::max<int64_t>(int64_t(nPos)/12, pipeMin < 4096 ? 64 : (int64_t(1) << 19));     // ~12 chunks per frame
			int64_t g = 0, ppos = 0, vpos = 0;
			for(const Segment& sg : h->segments)
			{
				const float* sP = sg.P; const float* sCi = sg.Ci; const float* sOi = sg.Oi; const uint8_t* sCu = sg.culled;
				if(sg.staged)
				{
					sP = h->stP.as<float>() + (size_t)(uintptr_t)sg.P;
					sCi = h->stCi.as<float>() + (size_t)(uintptr_t)sg.Ci;
					sOi = h->stOi.as<float>() + (size_t)(uintptr_t)sg.Oi;
					sCu = sg.culled ? h->stCulled.as<uint8_t>() + ((size_t)(uintptr_t)sg.culled - 1) : nullptr;
				}
				const int64_t segP0 = ppos, segV0 = vpos, gEnd = g + sg.nGrids;
				while(g < gEnd)
				{
					UploadChunk c{};
					c.p0 = ppos; c.v0 = vpos;
					while(g < gEnd && ppos - c.p0 < target)
					{
						const int64_t nv = int64_t(h->gcu[g]+1)*(h->gcv[g]+1);
						ppos += nv*h->gnkeys[g]; vpos += nv; ++g;
					}
					c.p1 = ppos; c.v1 = vpos;
					c.P = sP + (c.p0 - segP0)*3;
					c.Ci = sCi ? sCi + (c.v0 - segV0)*3 : nullptr;
					c.Oi = sOi ? sOi + (c.v0 - segV0)*3 : nullptr;
					c.culled = sCu ? sCu + (c.v0 - segV0) : nullptr;
					plan.push_back(c);
				}
			}
		}
		dP = h->dPraw.as<float>();
		dCi = h->anyCi ? h->dCi.as<float>() : nullptr;
		dOi = h->anyOi ? h->dOi.as<float>() : nullptr;
		dCulled = h->anyCulled ? h->dCulled.as<uint8_t>() : nullptr;
	}
	// ---- the rarely used arrays: arbitrary output variables, normals for the backface cull, point radii, CSG tables.
	// Plain copies on the main stream (they are ahead of every kernel that reads them).
	const float* dAov = nullptr; const float* dNg = nullptr; const float* dNn = nullptr; const float* dRadius = nullptr; const float* dTrimUV = nullptr;
	{
		struct Extra { bool any; DevBuf* buf; size_t perVertex, perPos; const float* Segment::* member; std::vector<float>* staged; const float** out; const char* what; };
		const size_t A = (size_t)h->aovFloats;
		Extra extras[] = {
			{h->anyAov && A > 0, &h->dAov, A, 0, &Segment::aov, &h->hxAov, &dAov, "cudaMemcpyAsync(arbitrary output variables)"},
			{h->anyNg, &h->dNg, 3, 0, &Segment::Ng, &h->hxNg, &dNg, "cudaMemcpyAsync(Ng)"},
			{h->anyN, &h->dNn, 3, 0, &Segment::N, &h->hxN, &dNn, "cudaMemcpyAsync(N)"},
			{h->anyRadius, &h->dRadius, 0, 1, &Segment::radius, &h->hxRadius, &dRadius, "cudaMemcpyAsync(point radii)"},
			{h->anyTrimUV, &h->dTrimUV, 2, 0, &Segment::trimUV, &h->hxTrimUV, &dTrimUV, "cudaMemcpyAsync(surface parameters of trimmed grids)"},
		};
		for(Extra& x : extras)
		{
			if(!x.any) continue;
			if(zeroCopy && h->segments[0].*(x.member)) { *x.out = h->segments[0].*(x.member); continue; }
			const size_t total = (x.perVertex ? nVerts*x.perVertex : nPos*x.perPos);
			CU(x.buf->reserve(std::max<size_t>(total, 1)*4), "cudaMalloc(extra grid arrays)");
			CU(cudaMemsetAsync(x.buf->p, 0, std::max<size_t>(total, 1)*4, st), "cudaMemsetAsync");
			size_t po2 = 0, vo2 = 0;
			for(const Segment& sg : h->segments)
			{
				const size_t off = x.perVertex ? vo2*x.perVertex : po2*x.perPos;
				const size_t cnt = x.perVertex ? size_t(sg.nVerts)*x.perVertex : size_t(sg.nPos)*x.perPos;
				const float* src = sg.*(x.member);
				cudaMemcpyKind kind = sg.memorySpace ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
				if(sg.staged)
				{
					// grids handed over one by one: the frame-global host copy (it may be shorter than the frame when later grids lack the array)
					src = nullptr; kind = cudaMemcpyHostToDevice;
					if(x.staged->size() > off) { src = x.staged->data() + off; if(x.staged->size() < off + cnt) { x.staged->resize(off + cnt, 0.f); src = x.staged->data() + off; } }
				}
				if(src && cnt)
				{
					CU(cudaMemcpyAsync(x.buf->as<float>() + off, src, cnt*4, kind, st), x.what);
					if(kind == cudaMemcpyHostToDevice) S.h2d_bytes += (int64_t)cnt*4;
				}
				po2 += (size_t)sg.nPos; vo2 += (size_t)sg.nVerts;
			}
			CU(cudaStreamSynchronize(st), "cudaStreamSynchronize(extra grid arrays)");      // pageable host memory: keep it simple
			*x.out = x.buf->as<float>();
		}
	}
	const bool anyCullT = std::find_if(h->gflags.begin(), h->gflags.end(), [](uint32_t fl) { return (fl & AQH_GRID_CULL_TRANSPARENT) != 0; }) != h->gflags.end();
	if(anyCullT)
	{
		const size_t g0 = incremental ? (size_t)h->upGrids : 0;
		CU(h->dGridTail.reserveKeep(std::max<size_t>(nRecs, 1)*4, g0*4, st), "cudaMalloc(grid tails)");
		if(nRecs > g0) CU(cudaMemsetAsync(h->dGridTail.as<uint32_t>() + g0, 0, (nRecs - g0)*4, st), "cudaMemsetAsync");
	}
	const size_t nCsg = h->csgType.size(), nCsgOrder = h->csgOrder.size();
	if(h->anyCSG)
	{
		CU(h->dGridCsg.reserve(std::max<size_t>(nRecs, 1)*4), "cudaMalloc(CSG nodes of the grids)");
		CU(cudaMemcpyAsync(h->dGridCsg.p, h->gcsg.data(), nRecs*4, cudaMemcpyHostToDevice, st), "cudaMemcpyAsync(CSG nodes of the grids)");
		CU(h->dCsgTab.reserve((4*nCsg + nCsgOrder + 1)*4), "cudaMalloc(CSG tree)");
		int32_t* tab = h->dCsgTab.as<int32_t>();
		CU(cudaMemcpyAsync(tab, h->csgType.data(), nCsg*4, cudaMemcpyHostToDevice, st), "cudaMemcpyAsync(CSG tree)");
		CU(cudaMemcpyAsync(tab + nCsg, h->csgParent.data(), nCsg*4, cudaMemcpyHostToDevice, st), "cudaMemcpyAsync(CSG tree)");
		CU(cudaMemcpyAsync(tab + 2*nCsg, h->csgSlot.data(), nCsg*4, cudaMemcpyHostToDevice, st), "cudaMemcpyAsync(CSG tree)");
		CU(cudaMemcpyAsync(tab + 3*nCsg, h->csgKids.data(), nCsg*4, cudaMemcpyHostToDevice, st), "cudaMemcpyAsync(CSG tree)");
		if(nCsgOrder) CU(cudaMemcpyAsync(tab + 4*nCsg, h->csgOrder.data(), nCsgOrder*4, cudaMemcpyHostToDevice, st), "cudaMemcpyAsync(CSG tree)");
		CU(cudaStreamSynchronize(st), "cudaStreamSynchronize(CSG tree)");
	}
	const size_t nTrimSets = h->trimSetLoop.empty() ? 0 : h->trimSetLoop.size() - 1, nTrimLoops = h->trimLoopPoint.empty() ? 0 : h->trimLoopPoint.size() - 1;
	if(h->anyTrim)
	{
		CU(h->dGridTrim.reserve(std::max<size_t>(nRecs, 1)*4), "cudaMalloc(trim sets of the grids)");
		CU(cudaMemcpyAsync(h->dGridTrim.p, h->gtrim.data(), nRecs*4, cudaMemcpyHostToDevice, st), "cudaMemcpyAsync(trim sets of the grids)");
		const size_t words = h->trimSetLoop.size() + h->trimLoopPoint.size() + h->trimPoints.size();
		CU(h->dTrimTab.reserve(std::max<size_t>(words, 1)*4 + 16), "cudaMalloc(trim loops)");
		int32_t* tab = h->dTrimTab.as<int32_t>();
		CU(cudaMemcpyAsync(tab, h->trimSetLoop.data(), h->trimSetLoop.size()*4, cudaMemcpyHostToDevice, st), "cudaMemcpyAsync(trim loops)");
		CU(cudaMemcpyAsync(tab + h->trimSetLoop.size(), h->trimLoopPoint.data(), h->trimLoopPoint.size()*4, cudaMemcpyHostToDevice, st), "cudaMemcpyAsync(trim loops)");
		if(!h->trimPoints.empty())
			CU(cudaMemcpyAsync(tab + h->trimSetLoop.size() + h->trimLoopPoint.size(), h->trimPoints.data(), h->trimPoints.size()*4, cudaMemcpyHostToDevice, st), "cudaMemcpyAsync(trim loops)");
		CU(cudaStreamSynchronize(st), "cudaStreamSynchronize(trim loops)");
	}
	if(nRecs) CU(cudaMemcpyAsync(h->dGrids.p, recs, nRecs*sizeof(GridRec), cudaMemcpyHostToDevice, st), "cudaMemcpyAsync(grid table)");
	CU(cudaMemcpyAsync(h->dChunk.p, h->chunk.data(), nChunkEntries*4, cudaMemcpyHostToDevice, st), "cudaMemcpyAsync(chunk table)");
	if(!h->gkeyTimes.empty())
		CU(cudaMemcpyAsync(h->dKeyTimes.p, h->gkeyTimes.data(), h->gkeyTimes.size()*4, cudaMemcpyHostToDevice, st), "cudaMemcpyAsync(key times)");
	CU(cudaMemcpyAsync(h->dTileSlot.p, h->tileSlot.data(), h->tileSlot.size()*4, cudaMemcpyHostToDevice, st), "cudaMemcpyAsync(tile slots)");
	if(nActive)
		CU(cudaMemcpyAsync(h->dActive.p, h->activeTiles.data(), size_t(nActive)*4, cudaMemcpyHostToDevice, st), "cudaMemcpyAsync(active tiles)");
	CU(cudaMemcpyAsync(h->dRowOwned.p, h->rowOwned.data(), p.yres, cudaMemcpyHostToDevice, st), "cudaMemcpyAsync(row ownership)");
	S.h2d_bytes += (int64_t)(nRecs*sizeof(GridRec) + nChunkEntries*4 + h->gkeyTimes.size()*4 + h->tileSlot.size()*4 + size_t(nActive)*4 + p.yres);
	// No host wait here unless the upload is being timed on its own: the tables come from pinned or staged memory
	// and everything that follows is ordered behind the copies on the stream.
	if(tr.on || !pipelined) CU(cudaStreamSynchronize(st), "cudaStreamSynchronize(upload)");
	S.upload_ms = nowMs() - tUp0;
	tr.mark("upload done");

	// ---- frame descriptor
	DevFrame f{};
	f.xres = p.xres; f.yres = p.yres;
	f.cropX0 = p.crop_xmin; f.cropX1 = p.crop_xmax; f.cropY0 = p.crop_ymin; f.cropY1 = p.crop_ymax;
	f.xs = p.xsamples; f.ys = p.ysamples; f.n = f.xs*f.ys;
	f.shiftX = L.shiftX; f.shiftY = L.shiftY;
	f.xfwo2 = std::ceil(p.filter_xwidth)*0.5f; f.yfwo2 = std::ceil(p.filter_ywidth)*0.5f;
	f.clipNear = p.clip_near; f.clipFar = p.clip_far;
	f.shutterOpen = p.shutter_open; f.shutterClose = p.shutter_close;
	f.useDof = p.use_dof ? 1 : 0;
	f.dofMult = p.dof_multiplier; f.dofInvFocal = p.dof_one_over_focal_distance;
	f.dofScaleX = p.dof_scale_x; f.dofScaleY = p.dof_scale_y;
	for(int k = 0; k < 3; ++k) f.zthr[k] = p.zthreshold[k];
	f.depthFilter = p.depth_filter;
	{
		const bool dz = (p.display_mode & AQH_DMODE_Z) != 0;
		f.cullable = !(dz && (p.depth_filter == AQH_DEPTHFILTER_MAX || p.depth_filter == AQH_DEPTHFILTER_AVERAGE)) ? 1 : 0;
		f.midpointZ = (dz && p.depth_filter == AQH_DEPTHFILTER_MIDPOINT) ? 1 : 0;
	}
	f.expGain = p.exposure_gain; f.expGamma = p.exposure_gamma;
	f.jitter = p.jitter;
	std::memcpy(f.camToRaster, p.cam_to_raster, sizeof f.camToRaster);
	f.anyMotion = anyMotion; f.anyTransparent = 0; f.anyLod = anyLod; f.anyTriangular = anyTri; f.anyCamera = anyCam;
	f.sx0 = L.sx0; f.sy0 = L.sy0; f.sw = L.sw; f.sh = L.sh;
	f.tileW = h->tileW; f.tileH = h->tileH; f.ntx = h->ntx; f.nty = h->nty; f.nActiveTiles = nActive;
	f.Praw = dP; f.Ci = dCi; f.Oi = dOi; f.culled = dCulled;
	f.Ng = dNg; f.Nn = dNn; f.radius = dRadius; f.aov = dAov;
	f.aovFloats = dAov ? h->aovFloats : 0; f.nch = nch;
	f.gridTail = anyCullT ? h->dGridTail.as<uint32_t>() : nullptr;
	f.cullTransparentOk = (anyCullT && !(p.zthreshold[0] == 0.f && p.zthreshold[1] == 0.f && p.zthreshold[2] == 0.f)) ? 1 : 0;
	f.anyCSG = h->anyCSG ? 1 : 0; f.nCsgNodes = (int)nCsg; f.nCsgOrder = (int)nCsgOrder;
	f.gridCsg = h->anyCSG ? h->dGridCsg.as<int32_t>() : nullptr;
	if(h->anyCSG)
	{
		const int32_t* tab = h->dCsgTab.as<int32_t>();
		f.csgType = tab; f.csgParent = tab + nCsg; f.csgSlot = tab + 2*nCsg; f.csgKids = tab + 3*nCsg; f.csgOrder = tab + 4*nCsg;
	}
	f.anyTrim = (h->anyTrim && dTrimUV) ? 1 : 0;
	if(f.anyTrim)
	{
		const int32_t* tab = h->dTrimTab.as<int32_t>();
		f.gridTrim = h->dGridTrim.as<int32_t>();
		f.trimSetLoop = tab; f.trimLoopPoint = tab + h->trimSetLoop.size();
		f.trimPoints = reinterpret_cast<const float2*>(tab + h->trimSetLoop.size() + h->trimLoopPoint.size());
		f.trimUV = reinterpret_cast<const float2*>(dTrimUV);
	}
	(void)nTrimSets; (void)nTrimLoops;
	// the displays may show arbitrary output variables (filtered by later passes) and an imager may still change the
	// pixels: then the filter only writes the float channel buffer and k_finish exposes / quantises afterwards
	const bool imager = imagerCb && imagerCb->on_imager && download && !zOnly;
	f.deferDisplay = (h->aovFloats > 0 || imager) ? 1 : 0;
	f.deferExpose = imager ? 1 : 0;
	// Streamed delivery: with display callbacks on a single rank the finished rows of every band travel back on the copy
	// stream as soon as they are filtered, and their buckets are handed to the callbacks (reference order) while the
	// device hides the next bands.
	const bool streamed = download && !zOnly && imagerCb && (imagerCb->on_bucket || imagerCb->on_data || imagerCb->on_progress) &&
	                      std::max(1, p.world_size) == 1 && !f.deferDisplay && h->copyStream && p.filter_mode == AQH_FILTER_REFERENCE_ORDER;
	if(streamed) bandTileRows = std::min(bandTileRows, std::max(2, (h->nty + 5)/6));
	h->deliveredRows = -1; h->deliveredBuckets = 0;
	f.grids = h->dGrids.as<GridRec>(); f.chunkGrid = h->dChunk.as<uint32_t>();
	f.keyTimes = h->dKeyTimes.as<float>(); f.splitLines = h->dSplit.as<float4>();
	f.nPos = h->nPos; f.nVerts = h->nVerts; f.nGrids = (int)nGrids;
	f.P4 = h->dP4.as<float4>();
	f.CO = h->dCO.as<float4>();
	// (the frame tables -- posTab, val1d, shufTab, patPlanes, filterTab, dofBounds, dither -- are joined and uploaded below, after
	// the grids: projection and binning do not read them)
	f.tileSlot = h->dTileSlot.as<int32_t>(); f.activeTiles = h->dActive.as<uint32_t>();
	f.binCount = h->dBinCount.as<uint32_t>(); f.binOffset = h->dBinOffset.as<uint32_t>();
	f.tileFlags = h->dTileFlags.as<uint32_t>();
	f.errorFlags = h->dMisc.as<uint32_t>() + 1;
	f.counters = reinterpret_cast<unsigned long long*>(h->dMisc.as<unsigned char>() + 16);
	f.planes = h->dPlanes.as<float>(); f.maskPlane = reinterpret_cast<unsigned char*>(h->dPlanes.as<float>() + size_t(7 + h->aovFloats)*planeStride);
	{ const int mbits = 2*L.shiftX + 2*L.shiftY + 3; f.maskBytes = mbits <= 8 ? 1 : (mbits <= 16 ? 2 : 4); } f.planeStride = (int64_t)planeStride; f.planeW = planeW; f.planeSC = planeSC; f.planeChunks = planeChunks; f.ringRows = ringRows;
	f.filterMode = p.filter_mode; f.partials = h->dPartials.as<float>(); f.ntaps = ntaps;
	f.channels = h->dChannels.as<float>();
	f.occlImage = nullptr; f.zOnly = zOnly ? 1 : 0;
	f.zKeys = nullptr; f.zKeys2 = nullptr; f.flushedPos = 0;
	if(zOnly)
	{
		CU(h->dOccl.reserve(size_t(L.sw)*L.sh*4), "cudaMalloc(occlusion image)");
		f.occlImage = h->dOccl.as<float>();
	}
	if(incremental)
	{
		const size_t nKeys = size_t(L.sw)*L.sh*size_t(p.xsamples*p.ysamples);
		const bool mid = (p.display_mode & AQH_DMODE_Z) && p.depth_filter == AQH_DEPTHFILTER_MIDPOINT;
		if(!h->haveZ)
		{
			CU(h->dZKeys.reserve(nKeys*8), "cudaMalloc(occlusion keys)");
			CU(launchFillKeys(h->dZKeys.as<unsigned long long>(), nKeys, st), "k_fill_keys");
			if(mid)
			{
				CU(h->dZKeys2.reserve(nKeys*8), "cudaMalloc(occlusion keys)");
				CU(launchFillKeys(h->dZKeys2.as<unsigned long long>(), nKeys, st), "k_fill_keys");
			}
		}
		f.zKeys = h->dZKeys.as<unsigned long long>();
		f.zKeys2 = mid ? h->dZKeys2.as<unsigned long long>() : nullptr;
		f.flushedPos = zOnly ? 0 : h->flushedPos;
	}
	f.rowOwned = (std::max(1, p.world_size) > 1) ? h->dRowOwned.as<uint8_t>() : nullptr;
	f.plain = (!h->anyPoints && !h->anyLodG && !h->anyTriG && !h->anyTrim && h->maxKeysG <= 2 && !h->anyCSG && h->aovFloats == 0 &&
	           !zOnly && !incremental && !f.midpointZ) ? 1 : 0;
	if(const char* e = std::getenv("AQH_TUNE"))
	{
		int k = 0;
		for(const char* q = e; *q && k < 8; ++k) { f.tune[k] = (int)std::strtol(q, const_cast<char**>(&q), 10); if(*q == ',') ++q; }
	}

	// ---- device work
	S.gpu_launches = 0;
	CU(cudaEventRecord(h->ev[0], st), "cudaEventRecord");
	CU(cudaMemsetAsync(h->dMisc.p, 0, 256, st), "cudaMemsetAsync");
	CU(cudaMemsetAsync(h->dBinCount.p, 0, std::max<size_t>(nActive, 1)*4, st), "cudaMemsetAsync");
	CU(cudaMemsetAsync(h->dChannels.p, 0, size_t(p.xres)*p.yres*nch*4, st), "cudaMemsetAsync");
	for(int d = 0; d < p.n_displays; ++d)
		CU(cudaMemsetAsync(h->dDisplay[d].p, 0, size_t(p.xres)*p.yres*disp.d[d].entrySize, st), "cudaMemsetAsync");
	if(pipelined && !plan.empty())
	{
		while(h->chunkEv.size() < plan.size() + 1)
		{
			cudaEvent_t e; CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "cudaEventCreate");
			h->chunkEv.push_back(e);
		}
		// the copies must not overtake the memsets / table uploads queued above
		CU(cudaEventRecord(h->chunkEv[plan.size()], st), "cudaEventRecord");
		CU(cudaStreamWaitEvent(h->copyStream, h->chunkEv[plan.size()], 0), "cudaStreamWaitEvent");
		for(size_t c = 0; c < plan.size(); ++c)
		{
			const UploadChunk& u = plan[c];
			CU(cudaMemcpyAsync(h->dPraw.as<float>() + u.p0*3, u.P, size_t(u.p1 - u.p0)*12, cudaMemcpyHostToDevice, h->copyStream), "cudaMemcpyAsync(P)");
			S.h2d_bytes += (u.p1 - u.p0)*12;
			if(u.Ci) { CU(cudaMemcpyAsync(h->dCi.as<float>() + u.v0*3, u.Ci, size_t(u.v1 - u.v0)*12, cudaMemcpyHostToDevice, h->copyStream), "cudaMemcpyAsync(Ci)"); S.h2d_bytes += (u.v1 - u.v0)*12; }
			if(u.Oi) { CU(cudaMemcpyAsync(h->dOi.as<float>() + u.v0*3, u.Oi, size_t(u.v1 - u.v0)*12, cudaMemcpyHostToDevice, h->copyStream), "cudaMemcpyAsync(Oi)"); S.h2d_bytes += (u.v1 - u.v0)*12; }
			if(u.culled) { CU(cudaMemcpyAsync(h->dCulled.as<uint8_t>() + u.v0, u.culled, size_t(u.v1 - u.v0), cudaMemcpyHostToDevice, h->copyStream), "cudaMemcpyAsync(culled)"); S.h2d_bytes += (u.v1 - u.v0); }
			CU(cudaEventRecord(h->chunkEv[c], h->copyStream), "cudaEventRecord");
			CU(cudaStreamWaitEvent(st, h->chunkEv[c], 0), "cudaStreamWaitEvent");
			CU(launchProjectCount(f, u.p0, u.p1, st), "k_project / k_bin<count>"); S.gpu_launches += 2;
		}
		CU(launchSplitLines(f, st), "k_splitlines"); S.gpu_launches += 1;
	}
	else if(incremental)
	{
		// only what arrived since the last flush is projected; a flush also bins only that, the final frame bins every
		// position again (k_bin skips the opaque micropolygons the stored occlusion keys already contain)
		CU(launchProject(f, (int64_t)pos0, f.nPos, st), "k_project"); S.gpu_launches += nPos > pos0 ? 1 : 0;
		CU(launchSplitLines(f, st), "k_splitlines"); S.gpu_launches += nGrids ? 1 : 0;
		CU(launchBinCount(f, zOnly ? (int64_t)pos0 : 0, f.nPos, st), "k_bin<count>"); S.gpu_launches += 1;
	}
	else
	{
		CU(launchProjectCount(f, 0, f.nPos, st), "k_project / k_bin<count>"); S.gpu_launches += nPos ? 2 : 0;
		CU(launchSplitLines(f, st), "k_splitlines"); S.gpu_launches += nGrids ? 1 : 0;
	}
	CU(launchBinScan(f, st), "k_bin_scan"); S.gpu_launches += 1;
	// the fill pass needs the total entry count to size the list
	uint32_t totalEntries = 0, devFlags = 0;
	unsigned long long maxBin = 0;
	CU(cudaMemcpyAsync(&totalEntries, f.binOffset + nActive, 4, cudaMemcpyDeviceToHost, st), "cudaMemcpyAsync(bin total)");
	CU(cudaMemcpyAsync(&maxBin, f.counters + 3, 8, cudaMemcpyDeviceToHost, st), "cudaMemcpyAsync(longest bin)");
	CU(cudaMemcpyAsync(&devFlags, f.errorFlags, 4, cudaMemcpyDeviceToHost, st), "cudaMemcpyAsync(frame flags)");
	tr.mark("launched project+count");
	CU(cudaStreamSynchronize(st), "bin count");
	tr.mark("bin count sync");
	// The frame tables: built on a host thread since aqh_begin_frame (the replay of the renderer's random stream), i.e. while
	// the grids travelled and were projected; first needed by k_hide.  A cached set (same options as the last frame) costs nothing.
	rc = uploadTables(h);
	if(rc) return rc;
	f.posTab = h->dPosTab.as<float2>(); f.val1d = h->dVal1d.as<float>(); f.shufTab = h->dShuf.as<uint8_t>(); f.shufInvTab = f.shufTab + h->shuf8.size()/2;
	f.patPlanes = h->dPat.as<uint8_t>(); f.filterTab = h->dFilt.as<float>(); f.dofBounds = h->dDofB.as<float4>();
	f.dither = h->dDither.as<float>();
	tr.mark("frame tables");
	CU(h->dBinEntries.reserve(std::max<size_t>(totalEntries, 1)*8), "cudaMalloc(bin entries)");
	f.binEntries = h->dBinEntries.as<unsigned long long>();
	// the project kernel reports whether any vertex is non-opaque: only then does the hide kernel
	// carry deep-list heads in shared memory and a deep hit pool in HBM
	if(devFlags & 2u) h->sawTransparent = true;          // k_project only saw what was new in this call
	f.anyTransparent = (h->sawTransparent || !f.cullable || h->anyCSG) ? 1 : 0;
	f.binPartition = (f.useDof || f.anyMotion) ? 0 : 1;
	LaunchCfg cfg{};
	CU(hideKernelConfig(f, h->smCount, cfg), "hide kernel configuration (shared memory / occupancy)");
	if(f.anyTransparent)
	{
		// every sample of a tile owns AQH_DEEP_INLINE in-line hit slots; the overflow pool behind them holds the rest of
		// deep_hits_per_sample (an average over the tile's samples)
		const int per = p.deep_hits_per_sample > 0 ? p.deep_hits_per_sample : 16;
		const size_t tileSamples = size_t(f.tileW)*f.tileH*f.n;
		const size_t nsP = size_t(f.tileH)*f.ys*(size_t(f.tileW)*f.xs + 8);          // padded sample slots of a tile (hideStride)
		f.deepCapPerCta = (uint32_t)std::min<size_t>(size_t(std::max(per - AQH_DEEP_INLINE, 0))*tileSamples, (size_t(1) << 20) - 2);
		const size_t perCta = AQH_DEEP_INLINE*nsP + f.deepCapPerCta;
		CU(h->dDeepA.reserve(size_t(cfg.hideCtas)*perCta*8 + 16), "cudaMalloc(transparent hit pool)");
		CU(h->dDeepB.reserve(size_t(cfg.hideCtas)*perCta*16 + 16), "cudaMalloc(transparent hit pool)");
		CU(h->dDeepUV.reserve(size_t(cfg.hideCtas)*std::max<size_t>(f.deepCapPerCta, 1)*4), "cudaMalloc(transparent hit pool)");
		f.deepA = h->dDeepA.as<uint2>(); f.deepB = h->dDeepB.as<uint4>(); f.deepNext = h->dDeepUV.as<uint32_t>();
	}
	// bins are sorted front to back in runs of sortRun entries (a power of two covering the longest bin, capped)
	f.sortRun = 64;
	while(f.sortRun < (int)maxBin && f.sortRun < 8192) f.sortRun <<= 1;
	CU(launchBinFill(f, (incremental && zOnly) ? (int64_t)pos0 : 0, f.nPos, st), "k_bin<fill>"); S.gpu_launches += (nPos ? 1 : 0) + ((nActive && f.anyTransparent) ? 1 : 0) + ((nPos && nActive) ? 1 : 0);
	CU(cudaEventRecord(h->ev[1], st), "cudaEventRecord");
	if(zOnly && !h->haveZ)
	{
		// every pixel of the sample region starts uncovered (FLT_MAX); tiles this rank does not hide stay that way
		std::vector<float> inf(size_t(L.sw)*L.sh, FLT_MAX);
		CU(cudaMemcpyAsync(h->dOccl.p, inf.data(), inf.size()*4, cudaMemcpyHostToDevice, st), "cudaMemcpyAsync(occlusion image)");
		CU(cudaStreamSynchronize(st), "cudaStreamSynchronize");
	}
	// ---- bands: consecutive active tile rows, at most bandTileRows of them, hidden by one launch each; the output rows
	// whose whole filter footprint is then in the planes are filtered before the next band overwrites the ring
	struct Band { uint32_t slotBeg, slotEnd; int rowBeg, rowEnd; bool firstOfRun; };   // rows: sample rows relative to sy0
	std::vector<Band> bands;
	{
		int prevTy = -2;
		for(int i = 0; i < nActive; i += h->ntx)
		{
			const int ty = (int)(h->activeTiles[i]/(uint32_t)h->ntx);
			const bool contiguous = ty == prevTy + 1;
			if(bands.empty() || !contiguous || (bands.back().rowEnd - bands.back().rowBeg) >= bandTileRows*h->tileH)
				bands.push_back(Band{(uint32_t)i, (uint32_t)i, ty*h->tileH, ty*h->tileH, !contiguous});
			bands.back().slotEnd = (uint32_t)(i + h->ntx);
			bands.back().rowEnd = std::min((ty + 1)*h->tileH, L.sh);
			prevTy = ty;
		}
	}
	S.n_bands = (int64_t)bands.size();
	S.d2h_bytes = 0;
	std::vector<int> bandRowsDone(bands.size(), -1);         // streamed delivery: the crop rows below this are on their way to the host
	if(streamed)
	{
		const size_t chBytes = size_t(p.xres)*p.yres*nch*4;
		if(!h->hChannels.reserve(chBytes)) return h->fail(AQH_ERR_NO_MEMORY, "cudaHostAlloc(channel image)");
		for(int d = 0; d < p.n_displays; ++d)
			if(!h->hDisplay[d].reserve(size_t(p.xres)*p.yres*disp.d[d].entrySize)) return h->fail(AQH_ERR_NO_MEMORY, "cudaHostAlloc(display image)");
		// rows outside the crop window never travel: zero like the device images
		const int outside[2][2] = {{0, p.crop_ymin}, {p.crop_ymax, p.yres}};
		for(const auto& o : outside)
		{
			if(o[1] <= o[0]) continue;
			std::memset(h->hChannels.as<unsigned char>() + size_t(o[0])*p.xres*nch*4, 0, size_t(o[1] - o[0])*p.xres*nch*4);
			for(int d = 0; d < p.n_displays; ++d)
				std::memset(h->hDisplay[d].as<unsigned char>() + size_t(o[0])*p.xres*disp.d[d].entrySize, 0, size_t(o[1] - o[0])*p.xres*disp.d[d].entrySize);
		}
		while(h->readyEv.size() < bands.size())
		{
			cudaEvent_t e; CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "cudaEventCreate");
			h->readyEv.push_back(e);
		}
		h->hostStripKey.clear();
	}
	CU(h->dBandCursor.reserve(std::max<size_t>(bands.size(), 1)*4), "cudaMalloc(band cursors)");
	CU(cudaMemsetAsync(h->dBandCursor.p, 0, std::max<size_t>(bands.size(), 1)*4, st), "cudaMemsetAsync");
	std::unique_lock<std::mutex> filterTable(g_filterTableMutex[h->device & 63], std::defer_lock);
	// stage times: one event after every launch (hide and filter launches of several bands interleave)
	while(h->bandEv.size() < 2*bands.size() + 1)
	{
		cudaEvent_t e; CU(cudaEventCreate(&e), "cudaEventCreate");
		h->bandEv.push_back(e);
	}
	CU(cudaEventRecord(h->bandEv[0], st), "cudaEventRecord");
	std::vector<int> bandEvKind;                          // per recorded event after bandEv[0]: 0 = a hide launch ended, 1 = a filter launch ended
	bool tableUp = false;
	int filtNext = 0;                                     // next output row (relative to the crop window) to filter
	for(size_t b = 0; b < bands.size(); ++b)
	{
		const Band& bd = bands[b];
		CU(launchHide(f, cfg, bd.slotBeg, bd.slotEnd, h->dBandCursor.as<uint32_t>() + b, st), "k_hide"); S.gpu_launches += 1;
		bandEvKind.push_back(0);
		CU(cudaEventRecord(h->bandEv[bandEvKind.size()], st), "cudaEventRecord");
		if(zOnly) continue;
		if(b == 0)
		{
			filterTable.lock();      // released when this function returns (the frame is synchronised by then)
			tr.mark("launched hide");
		}
		// output row r (relative to the crop window) reads the sample rows r .. r + 2*shiftY
		if(bd.firstOfRun) filtNext = bd.rowBeg;
		int filtEnd = bd.rowEnd - 2*L.shiftY;
		if(p.filter_mode != AQH_FILTER_REFERENCE_ORDER)
		{
			// the tap partial sums cover the whole frame: one filter launch after the last band
			if(b + 1 < bands.size()) continue;
			filtNext = 0; filtEnd = L.sh - 2*L.shiftY;
		}
		const int y0 = p.crop_ymin + std::max(filtNext, 0), y1 = std::min(p.crop_ymin + filtEnd, p.crop_ymax);
		if(y1 > y0)
		{
			CU(launchFilter(f, disp, h->filterTab.data(), y0, y1, !tableUp, st), "k_filter"); S.gpu_launches += filterLaunchCount(f);
			tableUp = true;
			bandEvKind.push_back(1);
			CU(cudaEventRecord(h->bandEv[bandEvKind.size()], st), "cudaEventRecord");
			if(streamed)
			{
				// the band's finished rows go home on the copy stream, behind the filter launch that wrote them
				CU(cudaStreamWaitEvent(h->copyStream, h->bandEv[bandEvKind.size()], 0), "cudaStreamWaitEvent");
				const size_t chRow = size_t(p.xres)*nch*4;
				CU(cudaMemcpyAsync(h->hChannels.as<unsigned char>() + chRow*y0, h->dChannels.as<unsigned char>() + chRow*y0, chRow*size_t(y1 - y0),
				                   cudaMemcpyDeviceToHost, h->copyStream), "cudaMemcpyAsync(channels)");
				S.d2h_bytes += (int64_t)(chRow*size_t(y1 - y0));
				for(int d = 0; d < p.n_displays; ++d)
				{
					const size_t dRow = size_t(p.xres)*disp.d[d].entrySize;
					CU(cudaMemcpyAsync(h->hDisplay[d].as<unsigned char>() + dRow*y0, h->dDisplay[d].as<unsigned char>() + dRow*y0, dRow*size_t(y1 - y0),
					                   cudaMemcpyDeviceToHost, h->copyStream), "cudaMemcpyAsync(display)");
					S.d2h_bytes += (int64_t)(dRow*size_t(y1 - y0));
				}
				CU(cudaEventRecord(h->readyEv[b], h->copyStream), "cudaEventRecord");
				bandRowsDone[b] = y1;
			}
		}
		filtNext = std::max(filtNext, filtEnd);
	}
	if(f.deferDisplay && !zOnly)
	{
		if(imager)
		{
			// Imager shading (bucketprocessor.cpp:712-743) happens on the host, bucket by bucket in reference order, between
			// filtering and exposure: the float channel buffer makes a round trip
			const size_t chBytesAll = size_t(p.xres)*p.yres*nch*4;
			if(!h->hChannels.reserve(chBytesAll)) return h->fail(AQH_ERR_NO_MEMORY, "cudaHostAlloc(channel image)");
			CU(cudaMemcpyAsync(h->hChannels.p, h->dChannels.p, chBytesAll, cudaMemcpyDeviceToHost, st), "cudaMemcpyAsync(channels)");
			CU(cudaStreamSynchronize(st), "cudaStreamSynchronize(imager)");
			S.d2h_bytes += (int64_t)chBytesAll;
			for(int row = L.by0; row < L.by1; ++row)
				for(int col = L.bx0; col < L.bx1; ++col)
				{
					const int xPos = col*p.bucket_xsize, yPos = row*p.bucket_ysize;
					const int xSize = std::min(p.bucket_xsize, p.xres - xPos), ySize = std::min(p.bucket_ysize, p.yres - yPos);
					// the bucket's display region, cropped like CqBucketProcessor::DisplayRegion
					const int x0 = std::max(xPos, p.crop_xmin), x1 = std::min(xPos + xSize, p.crop_xmax);
					const int y0 = std::max(yPos, p.crop_ymin), y1 = std::min(yPos + ySize, p.crop_ymax);
					if(x1 <= x0 || y1 <= y0) continue;
					bool mine = std::max(1, p.world_size) == 1;
					for(int y = y0; y < y1 && !mine; ++y) mine = h->rowOwned[y] != 0;
					if(!mine) continue;
					float* ch = h->hChannels.as<float>() + (size_t(y0)*p.xres + x0)*nch;
					if(imagerCb->on_imager(imagerCb->user, x0, x1, y0, y1, ch, p.xres*nch, nch))
						return h->fail(AQH_ERR_CALLBACK, "on_imager callback failed");
				}
			CU(cudaMemcpyAsync(h->dChannels.p, h->hChannels.p, chBytesAll, cudaMemcpyHostToDevice, st), "cudaMemcpyAsync(channels)");
			S.h2d_bytes += (int64_t)chBytesAll;
		}
		CU(launchFinish(f, disp, f.deferExpose, st), "k_finish"); S.gpu_launches += 1;
	}
	CU(cudaEventRecord(h->ev[3], st), "cudaEventRecord");

	// ---- results
	tr.mark("launched filter");
	const double tDown0 = nowMs();
	if(streamed)
	{
		// hand over the buckets of every bucket row as soon as all of its crop rows have arrived
		h->deliveredRows = L.by0;
		for(size_t b = 0; b < bands.size(); ++b)
		{
			if(bandRowsDone[b] < 0) continue;
			CU(cudaEventSynchronize(h->readyEv[b]), "cudaEventSynchronize(band rows)");
			int rowEnd = h->deliveredRows;
			while(rowEnd < L.by1 && std::min(std::min((rowEnd + 1)*p.bucket_ysize, p.yres), p.crop_ymax) <= bandRowsDone[b]) ++rowEnd;
			h->haveHostImage = true;
			rc = deliverBucketRows(h, imagerCb, rowEnd);
			if(rc) return rc;
		}
		download = false;             // everything is on the host already
	}
	// (a copy into pageable memory: the call returns when the device has done everything queued before it)
	struct { uint32_t cursor, err; uint32_t pad[2]; unsigned long long ctr[18]; } misc;
	CU(cudaMemcpyAsync(&misc, h->dMisc.p, sizeof misc, cudaMemcpyDeviceToHost, st), "cudaMemcpyAsync(counters)");
	// With a communicator the strips are gathered on the device first (aqh_gather): rank 0 then downloads the whole
	// frame and the other ranks nothing.
	const bool gathered = download && h->comm && std::max(1, p.world_size) > 1;
	if(gathered)
	{
		h->rendered = true;
		rc = aqh_gather(h, 0);
		if(rc) return rc;
		if(p.rank != 0) download = false;
	}
	if(download)
	{
		// Only the pixel rows this rank owns travel back (all of them on a single rank); the rest of the host
		// images stays zero, like the device images.
		const size_t chRow = size_t(p.xres)*nch*4, chBytes = chRow*p.yres;
		const bool sharded = std::max(1, p.world_size) > 1 && !gathered;
		const bool firstCh = h->hChannels.cap < chBytes;
		if(!h->hChannels.reserve(chBytes)) return h->fail(AQH_ERR_NO_MEMORY, "cudaHostAlloc(channel image)");
		if(sharded && (firstCh || h->hostStripKey != h->stripKey)) std::memset(h->hChannels.p, 0, chBytes);
		std::vector<std::pair<int,int>> rows;
		if(sharded) rows = h->strips; else rows.push_back(std::make_pair(0, p.yres));
		for(const auto& r : rows)
		{
			CU(cudaMemcpyAsync(h->hChannels.as<unsigned char>() + chRow*r.first, h->dChannels.as<unsigned char>() + chRow*r.first,
			                   chRow*size_t(r.second - r.first), cudaMemcpyDeviceToHost, st), "cudaMemcpyAsync(channels)");
			S.d2h_bytes += (int64_t)(chRow*size_t(r.second - r.first));
		}
		for(int d = 0; d < p.n_displays; ++d)
		{
			const size_t dRow = size_t(p.xres)*disp.d[d].entrySize, b = dRow*p.yres;
			const bool firstD = h->hDisplay[d].cap < b;
			if(!h->hDisplay[d].reserve(b)) return h->fail(AQH_ERR_NO_MEMORY, "cudaHostAlloc(display image)");
			if(sharded && (firstD || h->hostStripKey != h->stripKey)) std::memset(h->hDisplay[d].p, 0, b);
			for(const auto& r : rows)
			{
				CU(cudaMemcpyAsync(h->hDisplay[d].as<unsigned char>() + dRow*r.first, h->dDisplay[d].as<unsigned char>() + dRow*r.first,
				                   dRow*size_t(r.second - r.first), cudaMemcpyDeviceToHost, st), "cudaMemcpyAsync(display)");
				S.d2h_bytes += (int64_t)(dRow*size_t(r.second - r.first));
			}
		}
		h->hostStripKey = h->stripKey;
	}
	if(zOnly)
	{
		const size_t ob = size_t(L.sw)*L.sh*4;
		if(!h->hOccl.reserve(ob)) return h->fail(AQH_ERR_NO_MEMORY, "cudaHostAlloc(occlusion image)");
		CU(cudaMemcpyAsync(h->hOccl.p, h->dOccl.p, ob, cudaMemcpyDeviceToHost, st), "cudaMemcpyAsync(occlusion image)");
		S.d2h_bytes += (int64_t)ob;
	}
	CU(cudaStreamSynchronize(st), "cudaStreamSynchronize(frame)");
	S.download_ms = (download || streamed) ? nowMs() - tDown0 : 0.0;
	tr.mark("frame sync");
	h->haveHostImage = download || streamed;
	if(zOnly)
	{
		h->haveOccl = true;
		// what this flush hid is in the occlusion keys now; later grids open a new staged run
		h->haveZ = true; h->flushedPos = h->nPos; h->nFlushes += 1;
		h->upPos = h->nPos; h->upVerts = h->nVerts; h->upGrids = nGrids; h->upSegs = h->segments.size();
		if(!h->segments.empty()) h->segments.back().closed = true;
	}
	else if(incremental)
	{
		h->upPos = h->nPos; h->upVerts = h->nVerts; h->upGrids = nGrids; h->upSegs = h->segments.size();
		if(!h->segments.empty()) h->segments.back().closed = true;
	}
	float ms = 0;
	cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]); S.project_bust_ms = ms;
	S.render_mpgs_ms = 0; S.filter_ms = 0; S.display_ms = 0;
	for(size_t i = 0; i < bandEvKind.size(); ++i)
	{
		cudaEventElapsedTime(&ms, h->bandEv[i], h->bandEv[i + 1]);
		(bandEvKind[i] ? S.filter_ms : S.render_mpgs_ms) += ms;
	}
	cudaEventElapsedTime(&ms, h->ev[0], h->ev[3]); S.device_total_ms = ms;
	S.gather_ms = 0;
	if(h->gatherPending) { cudaEventElapsedTime(&ms, h->ev[4], h->ev[5]); S.gather_ms = ms; h->gatherPending = false; }
	S.n_grids = nGrids; S.n_vertices = h->nVerts;
	S.n_micropolygons = (int64_t)misc.ctr[0]; S.n_bin_entries = (int64_t)misc.ctr[1]; S.n_deep_hits = (int64_t)misc.ctr[2];
	S.n_samples = int64_t(L.sw)*L.sh*f.n;
	{
		const DevBuf* all[] = {&h->dPraw, &h->dCi, &h->dOi, &h->dCulled, &h->dP4, &h->dCO, &h->dGrids, &h->dChunk, &h->dKeyTimes, &h->dSplit,
		                       &h->dPosTab, &h->dVal1d, &h->dShuf, &h->dPat, &h->dFilt, &h->dDofB, &h->dDither, &h->dTileSlot, &h->dActive,
		                       &h->dBinCount, &h->dBinOffset, &h->dBinEntries, &h->dMisc, &h->dTileFlags, &h->dPlanes, &h->dPartials,
		                       &h->dDeepA, &h->dDeepB, &h->dDeepUV, &h->dChannels, &h->dRowOwned, &h->dOccl, &h->dBandCursor,
		                       &h->dAov, &h->dNg, &h->dNn, &h->dRadius, &h->dGridTail, &h->dGridCsg, &h->dCsgTab, &h->dZKeys, &h->dZKeys2,
		                       &h->dTrimUV, &h->dGridTrim, &h->dTrimTab};
		S.device_bytes = 0;
		for(const DevBuf* b : all) S.device_bytes += (int64_t)b->cap;
		for(int d = 0; d < AQH_MAX_DISPLAYS; ++d) S.device_bytes += (int64_t)h->dDisplay[d].cap;
	}
	h->rendered = true;
	tr.mark("stats");
#ifdef AQH_PHASE_TIMING
	{
		// development build: warp-cycles worked / waited per phase of k_hide (hider_kernels.cu, PHASE_BARRIER)
		static const char* names[7] = {"prepare", "opaque pass", "deep pass", "resolve", "tile fetch", "refresh", "occlusion image"};
		double tot = 0;
		for(int k = 0; k < 14; ++k) tot += (double)misc.ctr[4 + k];
		for(int k = 0; k < 7 && tot > 0; ++k)
			std::fprintf(stderr, "[aqh phase] %-16s work %6.2f %%  wait %6.2f %%\n", names[k], 100.0*misc.ctr[4 + 2*k]/tot, 100.0*misc.ctr[5 + 2*k]/tot);
	}
#endif
	if(misc.err & 1u)
		return h->fail(AQH_ERR_DEEP_OVERFLOW, "transparent hit pool exhausted: raise AqhFrameParams::deep_hits_per_sample");
	if(misc.err & 8u)
		return h->fail(AQH_ERR_DEEP_OVERFLOW, "a sample of a CSG frame holds more than 48 hits");
	return AQH_OK;
}

// Hand the bucket rows [h->deliveredRows, rowEnd) to the display callbacks, buckets in the reference's row-major order
// (imagebuffer.cpp:708-733, NextBucket :791-802), from the host images.
int deliverBucketRows(AqhHider* h, const AqhCallbacks* cb, int rowEnd)
{
	const AqhFrameParams& p = h->params;
	const ReplayLayout& L = h->layout;
	const int total = (L.bx1 - L.bx0)*(L.by1 - L.by0);
	std::vector<unsigned char>& bucketData = h->bucketScratch;
	for(int row = h->deliveredRows; row < rowEnd; ++row)
	{
		for(int col = L.bx0; col < L.bx1; ++col)
		{
			const int xPos = col*p.bucket_xsize, yPos = row*p.bucket_ysize;
			const int xSize = std::min(p.bucket_xsize, p.xres - xPos), ySize = std::min(p.bucket_ysize, p.yres - yPos);
			if(cb->on_bucket)
			{
				const float* ch = h->hChannels.as<float>() + (size_t(yPos)*p.xres + xPos)*h->nChannels;
				if(cb->on_bucket(cb->user, xPos, xPos + xSize, yPos, yPos + ySize, ch, p.xres*h->nChannels, h->nChannels))
					return h->fail(AQH_ERR_CALLBACK, "on_bucket callback failed");
			}
			if(cb->on_data)
				for(int d = 0; d < p.n_displays; ++d)
				{
					const int es = h->dispEntry[d];
					if(p.display[d].flags & AQH_DISPLAY_SCANLINE_ORDER)
					{
						// PkDspyFlagsWantsScanLineOrder: CollapseBucketsToScanlines gathers the buckets of a row; once the bucket
						// at the right edge arrives SendToDisplay delivers the rows one at a time (ddmanager.cpp:1129-1175)
						if(xPos + xSize < p.xres) continue;
						for(int y = yPos; y < yPos + ySize; ++y)
							if(cb->on_data(cb->user, d, 0, p.xres, y, y + 1, es, h->hDisplay[d].as<unsigned char>() + size_t(y)*p.xres*es))
								return h->fail(AQH_ERR_CALLBACK, "on_data callback failed");
						continue;
					}
					bucketData.resize(size_t(xSize)*ySize*es);
					for(int y = 0; y < ySize; ++y)
						std::memcpy(&bucketData[size_t(y)*xSize*es],
						            h->hDisplay[d].as<unsigned char>() + (size_t(yPos + y)*p.xres + xPos)*es, size_t(xSize)*es);
					if(cb->on_data(cb->user, d, xPos, xPos + xSize, yPos, yPos + ySize, es, bucketData.data()))
						return h->fail(AQH_ERR_CALLBACK, "on_data callback failed");
				}
			++h->deliveredBuckets;
			if(cb->on_progress) cb->on_progress(cb->user, (100.0f*h->deliveredBuckets)/static_cast<float>(total));
		}
		h->deliveredRows = row + 1;
	}
	if(rowEnd >= L.by1 && cb->on_progress) cb->on_progress(cb->user, 100.0f);
	return AQH_OK;
}

} // namespace

// =====================================================================================
extern "C" {

int aqh_abi_version(void) { return AQH_ABI_VERSION; }

int aqh_create(AqhHider** out, int device)
{
	if(!out) return AQH_ERR_BAD_PARAMS;
	*out = nullptr;
	int count = 0;
	if(cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count)
		return AQH_ERR_NO_DEVICE;
	if(cudaSetDevice(device) != cudaSuccess) return AQH_ERR_NO_DEVICE;
	cudaDeviceProp prop;
	if(cudaGetDeviceProperties(&prop, device) != cudaSuccess) return AQH_ERR_NO_DEVICE;
	if(!kernelsArchOk()) return AQH_ERR_NO_DEVICE;   // the library only carries sm_100a code
	AqhHider* h = new AqhHider;
	h->device = device;
	h->smCount = prop.multiProcessorCount;
	if(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { delete h; return AQH_ERR_CUDA; }
	h->ownStream = true;
	for(int i = 0; i < 8; ++i) cudaEventCreate(&h->ev[i]);
	if(cudaStreamCreateWithFlags(&h->copyStream, cudaStreamNonBlocking) != cudaSuccess) h->copyStream = nullptr;
	*out = h;
	return AQH_OK;
}

int aqh_destroy(AqhHider* h)
{
	if(!h) return AQH_OK;
	cudaSetDevice(h->device);
	joinTables(h);
	cudaStreamSynchronize(h->stream);
	aqh_comm_destroy(h);
	DevBuf* bufs[] = {&h->dPraw, &h->dCi, &h->dOi, &h->dCulled, &h->dP4, &h->dCO, &h->dGrids, &h->dChunk, &h->dKeyTimes, &h->dSplit,
	                  &h->dPosTab, &h->dVal1d, &h->dShuf, &h->dPat, &h->dFilt, &h->dDofB, &h->dDither, &h->dTileSlot, &h->dActive,
	                  &h->dBinCount, &h->dBinOffset, &h->dBinEntries, &h->dMisc, &h->dTileFlags, &h->dPlanes, &h->dMask, &h->dPartials,
	                  &h->dDeepA, &h->dDeepB, &h->dDeepUV, &h->dChannels, &h->dRowOwned, &h->dBandCursor,
	                  &h->dAov, &h->dNg, &h->dNn, &h->dRadius, &h->dGridTail, &h->dGridCsg, &h->dCsgTab, &h->dZKeys, &h->dZKeys2,
		                       &h->dTrimUV, &h->dGridTrim, &h->dTrimTab};
	for(DevBuf* b : bufs) b->release();
	for(int d = 0; d < AQH_MAX_DISPLAYS; ++d) { h->dDisplay[d].release(); h->hDisplay[d].release(); }
	h->dOccl.release(); h->hOccl.release(); h->recs.release(); h->chunk.release();
	h->hChannels.release(); h->stP.release(); h->stCi.release(); h->stOi.release(); h->stCulled.release();
	for(int i = 0; i < 8; ++i) if(h->ev[i]) cudaEventDestroy(h->ev[i]);
	for(cudaEvent_t e : h->chunkEv) cudaEventDestroy(e);
	for(cudaEvent_t e : h->readyEv) cudaEventDestroy(e);
	for(cudaEvent_t e : h->bandEv) cudaEventDestroy(e);
	if(h->copyStream) cudaStreamDestroy(h->copyStream);
	if(h->ownStream && h->stream) cudaStreamDestroy(h->stream);
	delete h;
	return AQH_OK;
}

const char* aqh_last_error(const AqhHider* h) { return h ? h->lastError.c_str() : "null hider"; }

int aqh_set_stream(AqhHider* h, void* cuda_stream)
{
	if(!h) return AQH_ERR_BAD_PARAMS;
	if(h->inFrame) return h->fail(AQH_ERR_STATE, "cannot change stream inside a frame");
	if(h->ownStream && h->stream) { cudaStreamSynchronize(h->stream); cudaStreamDestroy(h->stream); }
	h->stream = static_cast<cudaStream_t>(cuda_stream);
	h->ownStream = false;
	return AQH_OK;
}

int aqh_frame_params_default(AqhFrameParams* p)
{
	if(!p) return AQH_ERR_BAD_PARAMS;
	std::memset(p, 0, sizeof *p);
	p->abi_version = AQH_ABI_VERSION;
	// CqOptions defaults, libs/core/options.cpp:273-305
	p->xres = 640; p->yres = 480;
	p->crop_xmin = 0; p->crop_xmax = 640; p->crop_ymin = 0; p->crop_ymax = 480;
	p->xsamples = 2; p->ysamples = 2;
	p->filter_xwidth = 2.f; p->filter_ywidth = 2.f;
	p->filter_func = aqh_gaussian_filter;
	p->bucket_xsize = 16; p->bucket_ysize = 16;
	p->clip_near = FLT_EPSILON; p->clip_far = FLT_MAX;
	p->shutter_open = 0.f; p->shutter_close = 0.f;
	p->use_dof = 0;
	p->depth_filter = AQH_DEPTHFILTER_MIN;
	p->zthreshold[0] = p->zthreshold[1] = p->zthreshold[2] = 1.f;
	p->display_mode = AQH_DMODE_RGB | AQH_DMODE_A;
	p->exposure_gain = 1.f; p->exposure_gamma = 1.f;
	p->jitter = 1;
	for(int i = 0; i < 4; ++i) p->cam_to_raster[i*4+i] = 1.f;
	p->rng_seed = 545;      // CqRandom().Reseed(545), ri.cpp:660
	p->rng_predraws = 0;
	p->n_displays = 0;
	p->rank = 0; p->world_size = 1;
	p->filter_mode = AQH_FILTER_REFERENCE_ORDER;
	return AQH_OK;
}

int aqh_frame_params_set_dof(AqhFrameParams* p, float fstop, float focallength, float focaldistance, float scale_x, float scale_y)
{
	if(!p) return AQH_ERR_BAD_PARAMS;
	// CqRenderer::SetDepthOfFieldData, renderer.h:368-377
	p->use_dof = (fstop < FLT_MAX) ? 1 : 0;
	if(fstop < FLT_MAX)
	{
		float lensDiameter = focallength / fstop;
		p->dof_multiplier = static_cast<float>(0.5 * lensDiameter * focaldistance / (focaldistance + lensDiameter));
		p->dof_one_over_focal_distance = static_cast<float>(1.0 / focaldistance);
	}
	p->dof_scale_x = scale_x; p->dof_scale_y = scale_y;
	return AQH_OK;
}

int aqh_display_from_mode(AqhDisplayDesc* d, const char* mode, int driver_order, float one, float mn, float mx, float dither)
{
	if(!d || !mode) return AQH_ERR_BAD_PARAMS;
	std::memset(d, 0, sizeof *d);
	const bool rgb = std::strstr(mode, "rgb") != nullptr;
	const bool a = std::strchr(mode, 'a') != nullptr;
	const bool z = std::strchr(mode, 'z') != nullptr;
	if(!rgb && !a && !z) return AQH_ERR_BAD_PARAMS;
	int n = 0;
	if(!driver_order && a) d->channel[n++] = AQH_CH_ALPHA;     // core order a,r,g,b,z (ddmanager.cpp:455-480)
	if(rgb) { d->channel[n++] = AQH_CH_CI_R; d->channel[n++] = AQH_CH_CI_G; d->channel[n++] = AQH_CH_CI_B; }
	if(driver_order && a) d->channel[n++] = AQH_CH_ALPHA;      // file driver order r,g,b,a (display.cpp:454-490)
	if(z) d->channel[n++] = AQH_CH_Z;
	d->n_channels = n;
	d->type = 0;
	d->quantize_zero = 0.f; d->quantize_one = one; d->quantize_min = mn; d->quantize_max = mx; d->quantize_dither = dither;
	return AQH_OK;
}

int aqh_begin_frame(AqhHider* h, const AqhFrameParams* p)
{
	if(!h || !p) return AQH_ERR_BAD_PARAMS;
	if(cudaSetDevice(h->device) != cudaSuccess) return h->fail(AQH_ERR_NO_DEVICE, "cudaSetDevice failed");
	int rc = validateParams(h, *p);
	if(rc) return rc;
	h->params = *p;
	if(h->params.world_size < 1) { h->params.world_size = 1; h->params.rank = 0; }
	if(!h->params.filter_func) h->params.filter_func = aqh_gaussian_filter;
	std::memset(&h->stats, 0, sizeof h->stats);
	rc = buildTables(h);          // starts the frame-table job; prepare_ms is reported when it is joined
	if(rc) return rc;
	resetFrameGrids(h);
	h->aovFloats = 0;
	for(int a = 0; a < h->params.n_aovs; ++a) h->aovFloats += h->params.aov[a].n_floats;
	h->nChannels = AQH_NUM_CHANNELS + h->aovFloats;
	if(h->aovFloats && h->params.filter_mode != AQH_FILTER_REFERENCE_ORDER)
		return h->fail(AQH_ERR_UNSUPPORTED, "arbitrary output variables need the reference-order filter mode");
	h->inFrame = true; h->rendered = false; h->haveHostImage = false; h->haveOccl = false;
	return AQH_OK;
}

int aqh_add_grid(AqhHider* h, const AqhGridDesc* g)
{
	if(!h || !g) return AQH_ERR_BAD_PARAMS;
	if(!h->inFrame) return h->fail(AQH_ERR_STATE, "aqh_add_grid outside aqh_begin_frame/aqh_end_frame");
	if(!g->P) return h->fail(AQH_ERR_BAD_PARAMS, "grid without P");
	// every check and every allocation comes before the first change of state
	int rc = checkGrid(h, g->cu, g->cv, g->nkeys, g->flags, g->key_times, g->csg_node, g->radius, g->Ng, g->trim_set, g->trim_uv);
	if(rc) return rc;
	for(int k = 0; k < g->nkeys; ++k)
		if(!g->P[k]) return h->fail(AQH_ERR_BAD_PARAMS, "grid key without P");
	const size_t nv = size_t(g->cu+1)*(g->cv+1), np = nv*g->nkeys;
	// copy into the pinned staging run (the caller keeps ownership of its arrays)
	if(!h->stP.reserve((h->stPUsed + np)*12, true, h->stPUsed*12) ||
	   !h->stCi.reserve((h->stVUsed + nv)*12, true, h->stVUsed*12) ||
	   !h->stOi.reserve((h->stVUsed + nv)*12, true, h->stVUsed*12) ||
	   !h->stCulled.reserve(h->stVUsed + nv, true, h->stVUsed))
		return h->fail(AQH_ERR_NO_MEMORY, "cudaHostAlloc(grid staging)");
	{
		const GridTablesMark mark = markGridTables(h);
		rc = appendGridTables(h, g->cu, g->cv, g->nkeys, g->flags, g->lod_bounds, g->key_times, g->csg_node, g->trim_set);
		if(rc) { rollbackGridTables(h, mark); return rc; }
	}
	// the rarely used arrays go to frame-global host copies indexed by the grid's vertex / position offsets
	{
		const size_t v0 = (size_t)h->nVerts, p0 = (size_t)h->nPos;
		const size_t A = (size_t)h->aovFloats;
		if(g->aov && A) { h->hxAov.resize((v0 + nv)*A, 0.f); std::memcpy(&h->hxAov[v0*A], g->aov, nv*A*4); h->anyAov = true; }
		if(g->Ng) { h->hxNg.resize((v0 + nv)*3, 0.f); std::memcpy(&h->hxNg[v0*3], g->Ng, nv*12); h->anyNg = true; }
		if(g->N) { h->hxN.resize((v0 + nv)*3, 0.f); std::memcpy(&h->hxN[v0*3], g->N, nv*12); h->anyN = true; }
		if(g->radius && (g->flags & AQH_GRID_POINTS)) { h->hxRadius.resize(p0 + np, 0.f); std::memcpy(&h->hxRadius[p0], g->radius, np*4); h->anyRadius = true; }
		if(g->trim_set && g->trim_uv) { h->hxTrimUV.resize((v0 + nv)*2, 0.f); std::memcpy(&h->hxTrimUV[v0*2], g->trim_uv, nv*8); h->anyTrimUV = true; }
	}
	for(int k = 0; k < g->nkeys; ++k)
		std::memcpy(h->stP.as<float>() + (h->stPUsed + size_t(k)*nv)*3, g->P[k], nv*12);
	float* ci = h->stCi.as<float>() + h->stVUsed*3;
	float* oi = h->stOi.as<float>() + h->stVUsed*3;
	if(g->Ci) std::memcpy(ci, g->Ci, nv*12); else std::fill(ci, ci + nv*3, 1.0f);
	if(g->Oi) std::memcpy(oi, g->Oi, nv*12); else std::fill(oi, oi + nv*3, 1.0f);
	uint8_t* cu8 = h->stCulled.as<uint8_t>() + h->stVUsed;
	if(g->culled) std::memcpy(cu8, g->culled, nv); else std::memset(cu8, 0, nv);
	// extend the trailing staged segment or open a new one
	if(h->segments.empty() || !h->segments.back().staged || h->segments.back().closed)
	{
		Segment s;
		s.firstGrid = (int64_t)h->gcu.size() - 1;
		s.staged = true; s.memorySpace = 0;
		// staged pointers are kept as OFFSETS (the pinned buffers may move when they grow)
		s.P = reinterpret_cast<const float*>(uintptr_t(h->stPUsed*3));
		s.Ci = reinterpret_cast<const float*>(uintptr_t(h->stVUsed*3));
		s.Oi = reinterpret_cast<const float*>(uintptr_t(h->stVUsed*3));
		s.culled = reinterpret_cast<const uint8_t*>(uintptr_t(h->stVUsed + 1));
		h->segments.push_back(s);
	}
	Segment& s = h->segments.back();
	s.nGrids += 1; s.nVerts += (int64_t)nv; s.nPos += (int64_t)np;
	h->stPUsed += np; h->stVUsed += nv;
	h->nVerts += (int64_t)nv; h->nPos += (int64_t)np;
	h->anyCi = h->anyOi = true;
	if(g->culled) h->anyCulled = true;
	return AQH_OK;
}

int aqh_add_grid_block(AqhHider* h, const AqhGridBlock* b)
{
	if(!h || !b) return AQH_ERR_BAD_PARAMS;
	if(!h->inFrame) return h->fail(AQH_ERR_STATE, "aqh_add_grid_block outside aqh_begin_frame/aqh_end_frame");
	if(b->n_grids < 0 || (b->n_grids > 0 && (!b->cu || !b->cv || !b->flags || !b->P)))
		return h->fail(AQH_ERR_BAD_PARAMS, "grid block tables missing");
	if(b->memory_space != 0 && b->memory_space != 1) return h->fail(AQH_ERR_BAD_PARAMS, "memory_space");
	Segment s;
	s.firstGrid = (int64_t)h->gcu.size();
	s.nGrids = b->n_grids;
	s.P = b->P; s.Ci = b->Ci; s.Oi = b->Oi; s.culled = b->culled;
	s.aov = h->aovFloats ? b->aov : nullptr; s.Ng = b->Ng; s.N = b->N; s.radius = b->radius; s.trimUV = b->trim_set ? b->trim_uv : nullptr;
	s.memorySpace = b->memory_space;
	size_t ko = 0;
	// a block is accepted or rejected as a whole: all of its grids are checked before the first one is appended
	for(int64_t g = 0; g < b->n_grids; ++g)
	{
		const int nk = b->nkeys ? b->nkeys[g] : 1;
		int rc = checkGrid(h, b->cu[g], b->cv[g], nk, b->flags[g], (b->key_times && nk > 1) ? b->key_times + ko : nullptr,
		                   b->csg_node ? b->csg_node[g] : -1, b->radius, b->Ng, b->trim_set ? b->trim_set[g] : 0, b->trim_uv);
		if(rc) return rc;
		ko += nk;
	}
	ko = 0;
	const GridTablesMark mark = markGridTables(h);
	for(int64_t g = 0; g < b->n_grids; ++g)
	{
		const int nk = b->nkeys ? b->nkeys[g] : 1;
		int rc = appendGridTables(h, b->cu[g], b->cv[g], nk, b->flags[g], b->lod_bounds ? b->lod_bounds + 2*g : nullptr,
		                          (b->key_times && nk > 1) ? b->key_times + ko : nullptr, b->csg_node ? b->csg_node[g] : -1, b->trim_set ? b->trim_set[g] : 0);
		if(rc) { rollbackGridTables(h, mark); return rc; }
		ko += nk;
		const int64_t nv = int64_t(b->cu[g]+1)*(b->cv[g]+1);
		s.nVerts += nv; s.nPos += nv*nk;
	}
	if(b->n_grids == 0) return AQH_OK;
	h->segments.push_back(s);
	h->nVerts += s.nVerts; h->nPos += s.nPos;
	if(b->Ci) h->anyCi = true;
	if(b->Oi) h->anyOi = true;
	if(b->culled) h->anyCulled = true;
	if(s.aov) h->anyAov = true;
	if(b->Ng) h->anyNg = true;
	if(b->N) h->anyN = true;
	if(b->radius) h->anyRadius = true;
	if(s.trimUV) h->anyTrimUV = true;
	return AQH_OK;
}

int aqh_render_device(AqhHider* h)
{
	if(!h) return AQH_ERR_BAD_PARAMS;
	if(cudaSetDevice(h->device) != cudaSuccess) return h->fail(AQH_ERR_NO_DEVICE, "cudaSetDevice failed");
	return renderFrame(h, false);
}

int aqh_flush(AqhHider* h)
{
	if(!h) return AQH_ERR_BAD_PARAMS;
	if(!h->inFrame) return h->fail(AQH_ERR_STATE, "aqh_flush outside aqh_begin_frame/aqh_end_frame");
	if(cudaSetDevice(h->device) != cudaSuccess) return h->fail(AQH_ERR_NO_DEVICE, "cudaSetDevice failed");
	return renderFrame(h, false, true);
}

int aqh_can_cull(const AqhHider* h, const float bound[6], int* culled)
{
	if(!h || !bound || !culled) return AQH_ERR_BAD_PARAMS;
	*culled = 0;
	if(!h->inFrame || !h->haveOccl) return AQH_OK;
	const AqhFrameParams& p = h->params;
	// RenderSurface never asks the tree in this mode (bucketprocessor.cpp:945-948)
	if((p.display_mode & AQH_DMODE_Z) && (p.depth_filter == AQH_DEPTHFILTER_MAX || p.depth_filter == AQH_DEPTHFILTER_AVERAGE))
		return AQH_OK;
	const float xmin = bound[0], ymin = bound[1], zmin = bound[2], xmax = bound[3], ymax = bound[4];
	if(!(xmin <= xmax) || !(ymin <= ymax)) return AQH_ERR_BAD_PARAMS;
	// The occlusion image covers the SAMPLE region [crop - shift, crop + shift): every pixel of it carries samples
	// (the halo pixels are filtered into the border pixels) and belongs to exactly one bucket's CqOcclusionTree
	// (setupTree, occlusion.cpp:54-104).  canCull crops the bound to the tree (occlusion.cpp:164-168) and compares
	// inclusively, so a bound that ends exactly on a pixel edge still touches the pixel beyond it.
	const ReplayLayout& L = h->layout;
	const float rx0 = (float)L.sx0, ry0 = (float)L.sy0, rx1 = (float)(L.sx0 + L.sw), ry1 = (float)(L.sy0 + L.sh);
	if(xmax < rx0 || ymax < ry0 || xmin > rx1 || ymin > ry1) { *culled = 1; return AQH_OK; }   // no sample can be reached
	const float fx0 = std::max(std::ceil(xmin) - 1.0f, rx0), fx1 = std::min(std::floor(xmax) + 1.0f, rx1);
	const float fy0 = std::max(std::ceil(ymin) - 1.0f, ry0), fy1 = std::min(std::floor(ymax) + 1.0f, ry1);
	const float* occl = h->hOccl.as<float>();
	for(int y = (int)fy0; y < (int)fy1; ++y)
		for(int x = (int)fx0; x < (int)fx1; ++x)
			if(!(occl[size_t(y - L.sy0)*L.sw + (x - L.sx0)] < zmin)) return AQH_OK;
	*culled = 1;
	return AQH_OK;
}

int aqh_end_frame(AqhHider* h, const AqhCallbacks* cb)
{
	if(!h) return AQH_ERR_BAD_PARAMS;
	FrameTrace tr;
	if(cudaSetDevice(h->device) != cudaSuccess) return h->fail(AQH_ERR_NO_DEVICE, "cudaSetDevice failed");
	tr.mark("end_frame: set device");
	int rc = renderFrame(h, true, false, cb);
	tr.mark("end_frame: rendered");
	h->inFrame = false;
	if(rc) return rc;
	if(!cb || (!cb->on_bucket && !cb->on_data && !cb->on_progress)) return AQH_OK;
	if(!h->haveHostImage) return AQH_OK;          // a rank whose strips were gathered to rank 0: the displays live there
	if(h->deliveredRows >= h->layout.by1) return AQH_OK;        // renderFrame handed everything over while the device was busy
	if(h->deliveredRows < 0) h->deliveredRows = h->layout.by0;
	return deliverBucketRows(h, cb, h->layout.by1);    // whatever renderFrame has not handed over while the device was still busy
}

int aqh_set_csg_tree(AqhHider* h, int n_nodes, const int32_t* type, const int32_t* parent)
{
	if(!h) return AQH_ERR_BAD_PARAMS;
	if(!h->inFrame) return h->fail(AQH_ERR_STATE, "aqh_set_csg_tree outside aqh_begin_frame/aqh_end_frame");
	if(h->anyCSG) return h->fail(AQH_ERR_STATE, "the CSG tree must be set before the first CSG grid");
	if(n_nodes < 0 || (n_nodes > 0 && (!type || !parent)) || n_nodes > 65536) return h->fail(AQH_ERR_BAD_PARAMS, "CSG tree");
	std::vector<int32_t> ty(type, type + n_nodes), pa(parent, parent + n_nodes), slot(n_nodes, -1), kids(n_nodes, 0), order;
	for(int i = 0; i < n_nodes; ++i)
	{
		if(ty[i] < AQH_CSG_PRIMITIVE || ty[i] > AQH_CSG_DIFFERENCE || pa[i] >= n_nodes || pa[i] == i || pa[i] < -1)
			return h->fail(AQH_ERR_BAD_PARAMS, "CSG tree: node type or parent out of range");
		if(pa[i] >= 0)
		{
			if(ty[pa[i]] == AQH_CSG_PRIMITIVE) return h->fail(AQH_ERR_BAD_PARAMS, "CSG tree: a primitive cannot have children");
			slot[i] = kids[pa[i]]++;                   // children are ordered by node index (creation order)
			if(kids[pa[i]] > 32) return h->fail(AQH_ERR_UNSUPPORTED, "CSG tree: more than 32 children of one node");
		}
	}
	// children-before-parents order of the non-primitive nodes: the order CqCSGTreeNode::ProcessSampleList recurses in
	// (depth first, children in order, the node itself last); also rejects cycles
	std::vector<std::vector<int32_t> > children(n_nodes);
	for(int i = 0; i < n_nodes; ++i) if(pa[i] >= 0) children[pa[i]].push_back(i);
	std::vector<char> seen(n_nodes, 0);
	for(int root = 0; root < n_nodes; ++root)
	{
		if(pa[root] >= 0) continue;
		std::vector<std::pair<int32_t, size_t> > stack;
		stack.push_back(std::make_pair((int32_t)root, size_t(0)));
		seen[root] = 1;
		while(!stack.empty())
		{
			const int32_t n = stack.back().first;
			if(stack.back().second < children[n].size())
			{
				const int32_t c = children[n][stack.back().second++];
				seen[c] = 1;
				if(ty[c] != AQH_CSG_PRIMITIVE) stack.push_back(std::make_pair(c, size_t(0)));
			}
			else
			{
				if(ty[n] != AQH_CSG_PRIMITIVE) order.push_back(n);
				stack.pop_back();
			}
		}
	}
	for(int i = 0; i < n_nodes; ++i) if(!seen[i]) return h->fail(AQH_ERR_BAD_PARAMS, "CSG tree: cycle");
	h->csgType.swap(ty); h->csgParent.swap(pa); h->csgSlot.swap(slot); h->csgKids.swap(kids); h->csgOrder.swap(order);
	return AQH_OK;
}

int aqh_set_trim_loops(AqhHider* h, int n_sets, const int32_t* set_first_loop, const int32_t* loop_first_point, const float* points)
{
	if(!h) return AQH_ERR_BAD_PARAMS;
	if(!h->inFrame) return h->fail(AQH_ERR_STATE, "aqh_set_trim_loops outside aqh_begin_frame/aqh_end_frame");
	if(h->anyTrim) return h->fail(AQH_ERR_STATE, "the trim loops must be set before the first trimmed grid");
	if(n_sets < 0 || n_sets > (1 << 20) || (n_sets > 0 && (!set_first_loop || !loop_first_point))) return h->fail(AQH_ERR_BAD_PARAMS, "trim loops");
	h->trimSetLoop.clear(); h->trimLoopPoint.clear(); h->trimPoints.clear();
	if(n_sets == 0) return AQH_OK;
	if(set_first_loop[0] != 0) return h->fail(AQH_ERR_BAD_PARAMS, "trim loops: set_first_loop[0] must be 0");
	for(int s = 0; s < n_sets; ++s)
		if(set_first_loop[s+1] < set_first_loop[s]) return h->fail(AQH_ERR_BAD_PARAMS, "trim loops: set_first_loop must not decrease");
	const int nLoops = set_first_loop[n_sets];
	if(loop_first_point[0] != 0) return h->fail(AQH_ERR_BAD_PARAMS, "trim loops: loop_first_point[0] must be 0");
	for(int l = 0; l < nLoops; ++l)
		if(loop_first_point[l+1] < loop_first_point[l]) return h->fail(AQH_ERR_BAD_PARAMS, "trim loops: loop_first_point must not decrease");
	const int nPoints = loop_first_point[nLoops];
	if(nPoints > 0 && !points) return h->fail(AQH_ERR_BAD_PARAMS, "trim loops: points missing");
	// entry 0 of the per-grid trim_set means "untrimmed": set s is addressed as s + 1, so the device tables carry a leading 0
	h->trimSetLoop.assign(set_first_loop, set_first_loop + n_sets + 1);
	h->trimSetLoop.insert(h->trimSetLoop.begin(), 0);
	h->trimLoopPoint.assign(loop_first_point, loop_first_point + nLoops + 1);
	h->trimPoints.assign(points, points + size_t(nPoints)*2);
	return AQH_OK;
}

int aqh_clear_caches(AqhHider* h)
{
	if(!h) return AQH_ERR_BAD_PARAMS;
	if(h->inFrame) return h->fail(AQH_ERR_STATE, "aqh_clear_caches inside a frame");
	joinTables(h);
	h->tableKey.clear(); h->tablesUploaded = false;
	return AQH_OK;
}

int aqh_capture_on_bucket(void* user, int xmin, int xmax1, int ymin, int ymax1, const float* channels, int row_stride_floats, int pixel_stride_floats)
{
	AqhCapture* c = static_cast<AqhCapture*>(user);
	if(!c) return 1;
	c->buckets += 1;
	if(!c->channels) return 0;
	if(pixel_stride_floats != c->n_channels || xmin < 0 || ymin < 0 || xmax1 > c->xres || ymax1 > c->yres) return 1;
	const size_t rowBytes = size_t(xmax1 - xmin)*pixel_stride_floats*4;
	for(int y = ymin; y < ymax1; ++y)
		std::memcpy(c->channels + (size_t(y)*c->xres + xmin)*c->n_channels, channels + size_t(y - ymin)*row_stride_floats, rowBytes);
	c->bytes += (int64_t)rowBytes*(ymax1 - ymin);
	return 0;
}

int aqh_capture_on_data(void* user, int display, int xmin, int xmax1, int ymin, int ymax1, int entrysize, const unsigned char* data)
{
	AqhCapture* c = static_cast<AqhCapture*>(user);
	if(!c || display < 0 || display >= AQH_MAX_DISPLAYS) return 1;
	if(!c->display[display]) return 0;
	if(xmin < 0 || ymin < 0 || xmax1 > c->xres || ymax1 > c->yres) return 1;
	const size_t rowBytes = size_t(xmax1 - xmin)*entrysize;
	for(int y = ymin; y < ymax1; ++y)
		std::memcpy(c->display[display] + (size_t(y)*c->xres + xmin)*entrysize, data + size_t(y - ymin)*rowBytes, rowBytes);
	c->bytes += (int64_t)rowBytes*(ymax1 - ymin);
	return 0;
}

int aqh_channel_count(const AqhHider* h, int* n)
{
	if(!h || !n) return AQH_ERR_BAD_PARAMS;
	*n = h->nChannels;
	return AQH_OK;
}

int aqh_frame_stats(const AqhHider* h, AqhFrameStats* out)
{
	if(!h || !out) return AQH_ERR_BAD_PARAMS;
	*out = h->stats;
	return AQH_OK;
}

int aqh_image_channels(const AqhHider* h, const float** data, int* width, int* height)
{
	if(!h || !data) return AQH_ERR_BAD_PARAMS;
	if(!h->haveHostImage) return AQH_ERR_STATE;
	*data = h->hChannels.as<float>();
	if(width) *width = h->params.xres;
	if(height) *height = h->params.yres;
	return AQH_OK;
}

int aqh_image_display(const AqhHider* h, int display, const unsigned char** data, int* entrysize, int* type)
{
	if(!h || !data || display < 0 || display >= h->params.n_displays) return AQH_ERR_BAD_PARAMS;
	if(!h->haveHostImage) return AQH_ERR_STATE;
	*data = h->hDisplay[display].as<unsigned char>();
	if(entrysize) *entrysize = h->dispEntry[display];
	if(type) *type = h->dispType[display];
	return AQH_OK;
}

int aqh_device_channels(const AqhHider* h, void** dev_ptr, size_t* bytes)
{
	if(!h || !dev_ptr) return AQH_ERR_BAD_PARAMS;
	if(!h->rendered) return AQH_ERR_STATE;
	*dev_ptr = h->dChannels.p;
	if(bytes) *bytes = size_t(h->params.xres)*h->params.yres*size_t(h->nChannels)*4;
	return AQH_OK;
}

int aqh_device_display(const AqhHider* h, int display, void** dev_ptr, size_t* bytes)
{
	if(!h || !dev_ptr || display < 0 || display >= h->params.n_displays) return AQH_ERR_BAD_PARAMS;
	if(!h->rendered) return AQH_ERR_STATE;
	*dev_ptr = h->dDisplay[display].p;
	if(bytes) *bytes = size_t(h->params.xres)*h->params.yres*h->dispEntry[display];
	return AQH_OK;
}

int aqh_strip_layout(const AqhFrameParams* p, int rank, int* n_strips, int* y0, int* y1, int capacity)
{
	if(!p || !n_strips || p->crop_ymax <= p->crop_ymin) return AQH_ERR_BAD_PARAMS;
	const int world = std::max(1, p->world_size);
	if(rank < 0 || rank >= world) return AQH_ERR_BAD_PARAMS;
	std::vector<std::pair<int,int>> strips;
	computeStrips(*p, rank, strips);
	*n_strips = (int)strips.size();
	for(int i = 0; i < (int)strips.size() && i < capacity; ++i)
	{
		if(y0) y0[i] = strips[i].first;
		if(y1) y1[i] = strips[i].second;
	}
	return AQH_OK;
}

int aqh_num_strips(const AqhHider* h, int* n)
{
	if(!h || !n) return AQH_ERR_BAD_PARAMS;
	*n = (int)h->strips.size();
	return AQH_OK;
}

int aqh_strip(const AqhHider* h, int i, int* y0, int* y1)
{
	if(!h || i < 0 || i >= (int)h->strips.size()) return AQH_ERR_BAD_PARAMS;
	if(y0) *y0 = h->strips[i].first;
	if(y1) *y1 = h->strips[i].second;
	return AQH_OK;
}

// ---- host-side leaves ----------------------------------------------------------------
struct AqhRandom { Random r; };
AqhRandom* aqh_random_create(uint32_t seed) { AqhRandom* r = new AqhRandom; r->r.reseed(seed); return r; }
void aqh_random_destroy(AqhRandom* r) { delete r; }
void aqh_random_reseed(AqhRandom* r, uint32_t seed) { r->r.reseed(seed); }
uint32_t aqh_random_uint(AqhRandom* r) { return r->r.nextUint(); }
float aqh_random_float(AqhRandom* r) { return r->r.nextFloat(); }
uint32_t aqh_random_int(AqhRandom* r, uint32_t range) { return r->r.nextInt(range); }

int aqh_sampler_tables(AqhRandom* r, int xs, int ys, int jitter, float* pos, float* v1d, int32_t* shuffled, int* ncache)
{
	if(!r || xs < 1 || ys < 1 || !pos || !v1d || !shuffled) return AQH_ERR_BAD_PARAMS;
	SamplerTables t;
	if(jitter) buildJitterTables(r->r, xs, ys, t); else buildGridTables(xs, ys, t);
	std::memcpy(pos, t.pos.data(), t.pos.size()*4);
	std::memcpy(v1d, t.val1d.data(), t.val1d.size()*4);
	std::memcpy(shuffled, t.shuffled.data(), t.shuffled.size()*4);
	if(ncache) *ncache = t.ncache;
	return AQH_OK;
}

int aqh_replay_frame_rng(const AqhFrameParams* p, uint8_t* planes, float* dither, int* sx0, int* sy0, int* sw, int* sh)
{
	if(!p || p->xsamples < 1 || p->ysamples < 1 || p->bucket_xsize < 1 || p->bucket_ysize < 1) return AQH_ERR_BAD_PARAMS;
	ReplayLayout L = replayLayout(*p);
	if(sx0) *sx0 = L.sx0; if(sy0) *sy0 = L.sy0; if(sw) *sw = L.sw; if(sh) *sh = L.sh;
	if(!planes) return AQH_OK;
	Random rng(p->rng_seed);
	rng.discard(p->rng_predraws);
	SamplerTables jit;
	buildJitterTables(rng, p->xsamples, p->ysamples, jit);
	replayFrame(*p, L, rng, p->jitter != 0, planes, dither);
	return AQH_OK;
}

int aqh_filter_table(const AqhFrameParams* p, float* table, int* n_entries)
{
	if(!p) return AQH_ERR_BAD_PARAMS;
	std::vector<float> t;
	buildFilterTable(*p, t);
	if(table) std::memcpy(table, t.data(), t.size()*4);
	if(n_entries) *n_entries = (int)t.size();
	return AQH_OK;
}

} // extern "C"
