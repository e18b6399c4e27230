// hider_shard.cpp -- one frame over the GPUs of a box: which rank needs which grid, strips of equal
// work, and the single collective of the path -- the gather of the finished strips to one rank over NCCL.
//
// The sharding helpers are host code without any device dependency (a front end calls them while it still holds the
// grids in host memory).  NCCL is bound at run time from libnccl.so.2, so the library loads on machines without it.
#include "hider_internal.h"

#include <cmath>
#include <dlfcn.h>
#include <mutex>
#include <thread>

using namespace aqh;

namespace {

struct RowRange { int lo, hi; };

// Raster row range a grid can contribute to: union of its keys, grown by the largest circle of confusion of its depth
// range and by the filter half-width + 1 (imagebuffer.cpp:519-554 does this per micropolygon; per grid it is a superset).
RowRange gridRowRange(const AqhFrameParams& p, const float* P, int64_t npos, bool cameraSpace)
{
	float ymin = INFINITY, ymax = -INFINITY, zmin = INFINITY, zmax = -INFINITY;
	const float* m = p.cam_to_raster;
	for(int64_t i = 0; i < npos; ++i)
	{
		float x = P[3*i], y = P[3*i+1], z = P[3*i+2];
		if(cameraSpace)
		{
			// the projection of k_project (CqMatrix::operator*, include/aqsis/math/matrix.h:717-750)
			float h = (m[3]*x + m[7]*y + m[11]*z + m[15]);
			float ry = (m[1]*x + m[5]*y + m[9]*z + m[13]);
			if(h != 1.f) ry = ry*(1.f/h);
			y = ry;
		}
		ymin = std::min(ymin, y); ymax = std::max(ymax, y);
		zmin = std::min(zmin, z); zmax = std::max(zmax, z);
	}
	double pad = std::floor(p.filter_ywidth/2.0) + 1.0;
	if(p.use_dof)
	{
		// |1/z - 1/fd| is convex in 1/z: over the depth range its maximum is at an end point
		auto coc = [&](double z) { return (double)p.dof_multiplier*std::fabs(1.0/z - (double)p.dof_one_over_focal_distance)*(double)p.dof_scale_y; };
		pad += std::max(coc(zmin), coc(zmax))*1.001 + 1e-3;
	}
	RowRange r;
	const double lo = std::floor((double)ymin - pad), hi = std::ceil((double)ymax + pad);
	r.lo = (int)std::max(-1.0e9, std::min(1.0e9, lo));
	r.hi = (int)std::max(-1.0e9, std::min(1.0e9, hi));
	if(!(ymin <= ymax)) { r.lo = 1; r.hi = 0; }      // no finite vertex: touches nothing
	return r;
}

// Runs fn(g, rowRange) over the grids of a host block on all host threads.
template<class Fn> int forEachGridRange(const AqhFrameParams& p, const AqhGridBlock& b, Fn fn)
{
	if(b.memory_space != 0) return AQH_ERR_BAD_PARAMS;
	if(b.n_grids <= 0) return AQH_OK;
	if(!b.cu || !b.cv || !b.P) return AQH_ERR_BAD_PARAMS;
	std::vector<int64_t> pstart((size_t)b.n_grids + 1, 0);
	for(int64_t g = 0; g < b.n_grids; ++g)
	{
		const int64_t nv = int64_t(b.cu[g] + 1)*(b.cv[g] + 1), nk = b.nkeys ? b.nkeys[g] : 1;
		pstart[g + 1] = pstart[g] + nv*nk;
	}
	unsigned nt = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
	if(b.n_grids < 256) nt = 1;
	std::vector<std::thread> pool;
	auto work = [&](int64_t g0, int64_t g1) {
		for(int64_t g = g0; g < g1; ++g)
		{
			const bool cam = b.flags && (b.flags[g] & AQH_GRID_CAMERA_SPACE);
			fn(g, gridRowRange(p, b.P + 3*pstart[g], pstart[g + 1] - pstart[g], cam));
		}
	};
	for(unsigned t = 1; t < nt; ++t) pool.emplace_back(work, b.n_grids*t/nt, b.n_grids*(t + 1)/nt);
	work(0, b.n_grids/nt);
	for(auto& th : pool) th.join();
	return AQH_OK;
}

// ---- NCCL, bound at run time ------------------------------------------------------------------------------------------
typedef struct { char internal[128]; } NcclUniqueId;
typedef void* NcclComm;
struct NcclApi
{
	int (*GetUniqueId)(NcclUniqueId*);
	int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int);
	int (*CommDestroy)(NcclComm);
	int (*GroupStart)();
	int (*GroupEnd)();
	int (*Send)(const void*, size_t, int /*ncclDataType_t*/, int, NcclComm, cudaStream_t);
	int (*Recv)(void*, size_t, int, int, NcclComm, cudaStream_t);
	const char* (*GetErrorString)(int);
	bool ok = false;
};
NcclApi& nccl()
{
	static NcclApi api;
	static std::once_flag once;
	std::call_once(once, [] {
		// an already loaded libnccl (e.g. the one a host framework brought) is found first by its soname
		void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
		if(!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
		if(!lib) return;
		bool all = true;
		auto sym = [&](const char* n) { void* s = dlsym(lib, n); if(!s) all = false; return s; };
		api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
		api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
		api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
		api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
		api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
		api.Send = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
		api.Recv = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
		api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
		api.ok = all;
	});
	return api;
}
int ncclFail(AqhHider* h, int rc, const char* what)
{
	h->lastError = std::string(what) + ": " + (nccl().GetErrorString ? nccl().GetErrorString(rc) : "NCCL error");
	return AQH_ERR_CUDA;
}
const int kNcclInt8 = 0;   // ncclInt8 / ncclChar

} // namespace

extern "C" {

int aqh_grid_rank_masks(const AqhFrameParams* p, const AqhGridBlock* b, uint64_t* rank_mask)
{
	if(!p || !b || (b->n_grids > 0 && !rank_mask)) return AQH_ERR_BAD_PARAMS;
	const int world = std::max(1, p->world_size);
	if(world > AQH_MAX_RANKS) return AQH_ERR_BAD_PARAMS;
	std::vector<std::vector<std::pair<int,int>>> strips(world);
	for(int r = 0; r < world; ++r) computeStrips(*p, r, strips[r]);
	return forEachGridRange(*p, *b, [&](int64_t g, RowRange rr) {
		uint64_t m = 0;
		for(int r = 0; r < world; ++r)
			for(const auto& s : strips[r])
				if(rr.hi >= s.first && rr.lo < s.second) { m |= uint64_t(1) << r; break; }
		rank_mask[g] = m;
	});
}

int aqh_grid_row_cost(const AqhFrameParams* p, const AqhGridBlock* b, double* row_cost)
{
	if(!p || !b || !row_cost || p->yres <= 0) return AQH_ERR_BAD_PARAMS;
	std::vector<RowRange> ranges((size_t)std::max<int64_t>(b->n_grids, 0));
	int rc = forEachGridRange(*p, *b, [&](int64_t g, RowRange rr) { ranges[(size_t)g] = rr; });
	if(rc) return rc;
	for(int64_t g = 0; g < b->n_grids; ++g)
	{
		// the micropolygons of the grid, spread evenly over the rows the grid itself covers (padding excluded as far as known)
		const int lo = std::max(ranges[g].lo, 0), hi = std::min(ranges[g].hi, p->yres - 1);
		if(hi < lo) continue;
		const double w = double(b->cu[g])*double(std::max(b->cv[g], 1))/double(hi - lo + 1);
		for(int y = lo; y <= hi; ++y) row_cost[y] += w;
	}
	return AQH_OK;
}

int aqh_balance_strips(AqhFrameParams* p, const double* row_cost)
{
	if(!p || !row_cost) return AQH_ERR_BAD_PARAMS;
	const int world = std::max(1, p->world_size);
	if(world > AQH_MAX_RANKS || p->crop_ymax <= p->crop_ymin) return AQH_ERR_BAD_PARAMS;
	double total = 0;
	for(int y = p->crop_ymin; y < p->crop_ymax; ++y) total += row_cost[y] + 1e-9;
	p->strip_bounds[0] = p->crop_ymin;
	double run = 0;
	int y = p->crop_ymin;
	for(int r = 1; r < world; ++r)
	{
		const double want = total*r/world;
		while(y < p->crop_ymax && run + row_cost[y] + 1e-9 <= want) { run += row_cost[y] + 1e-9; ++y; }
		// the hide kernel works in tiles of 4..16 rows: keep the cuts on multiples of 16 rows where the strips are tall
		const int snap = (p->crop_ymax - p->crop_ymin) >= 128*world ? 16 : 4;
		int cut = ((y + snap/2)/snap)*snap;
		cut = std::min(std::max(cut, p->strip_bounds[r - 1]), p->crop_ymax);
		p->strip_bounds[r] = cut;
	}
	p->strip_bounds[world] = p->crop_ymax;
	p->strip_rows = -2;
	return AQH_OK;
}

int aqh_comm_unique_id(void* id128)
{
	if(!id128) return AQH_ERR_BAD_PARAMS;
	if(!nccl().ok) return AQH_ERR_UNSUPPORTED;
	NcclUniqueId id;
	if(nccl().GetUniqueId(&id) != 0) return AQH_ERR_CUDA;
	std::memcpy(id128, &id, sizeof id);
	return AQH_OK;
}

int aqh_comm_init(AqhHider* h, const void* id128, int rank, int world_size)
{
	if(!h || !id128 || world_size < 1 || world_size > AQH_MAX_RANKS || rank < 0 || rank >= world_size) return AQH_ERR_BAD_PARAMS;
	if(!nccl().ok) return h->fail(AQH_ERR_UNSUPPORTED, "libnccl.so.2 could not be loaded");
	if(cudaSetDevice(h->device) != cudaSuccess) return h->fail(AQH_ERR_NO_DEVICE, "cudaSetDevice failed");
	if(h->comm) { nccl().CommDestroy(h->comm); h->comm = nullptr; }
	NcclUniqueId id;
	std::memcpy(&id, id128, sizeof id);
	NcclComm c = nullptr;
	const int rc = nccl().CommInitRank(&c, world_size, id, rank);
	if(rc != 0) return ncclFail(h, rc, "ncclCommInitRank");
	h->comm = c; h->commRank = rank; h->commWorld = world_size;
	return AQH_OK;
}

int aqh_comm_destroy(AqhHider* h)
{
	if(!h) return AQH_ERR_BAD_PARAMS;
	if(h->comm && nccl().ok) { cudaSetDevice(h->device); cudaStreamSynchronize(h->stream); nccl().CommDestroy(h->comm); }
	h->comm = nullptr; h->commRank = 0; h->commWorld = 1;
	return AQH_OK;
}

// Every strip is a contiguous range of rows of the full-size device images, so a rank's strips travel straight from
// its images into the same rows of root's images: no packing kernel, no staging buffer.
int aqh_gather(AqhHider* h, int root)
{
	if(!h) return AQH_ERR_BAD_PARAMS;
	if(!h->rendered) return h->fail(AQH_ERR_STATE, "aqh_gather before a frame was rendered");
	const AqhFrameParams& p = h->params;
	const int world = std::max(1, p.world_size);
	if(world == 1) return AQH_OK;
	if(!h->comm) return h->fail(AQH_ERR_STATE, "aqh_gather without a communicator (aqh_comm_init)");
	if(h->commWorld != world || h->commRank != p.rank || root < 0 || root >= world)
		return h->fail(AQH_ERR_BAD_PARAMS, "communicator and frame disagree on rank / world size");
	if(cudaSetDevice(h->device) != cudaSuccess) return h->fail(AQH_ERR_NO_DEVICE, "cudaSetDevice failed");
	NcclApi& N = nccl();
	cudaStream_t st = h->stream;
	cudaEventRecord(h->ev[4], st);
	int rc = N.GroupStart();
	if(rc != 0) return ncclFail(h, rc, "ncclGroupStart");
	const size_t chRow = size_t(p.xres)*size_t(h->nChannels)*4;
	std::vector<std::pair<int,int>> strips;
	for(int r = 0; r < world && rc == 0; ++r)
	{
		if(r == root || (p.rank != root && p.rank != r)) continue;
		computeStrips(p, r, strips);
		for(const auto& s : strips)
		{
			const size_t rows = size_t(s.second - s.first);
			for(int img = -1; img < p.n_displays && rc == 0; ++img)
			{
				const size_t rowBytes = img < 0 ? chRow : size_t(p.xres)*size_t(h->dispEntry[img]);
				unsigned char* base = (img < 0 ? h->dChannels.as<unsigned char>() : h->dDisplay[img].as<unsigned char>()) + rowBytes*size_t(s.first);
				if(p.rank == root) rc = N.Recv(base, rowBytes*rows, kNcclInt8, r, h->comm, st);
				else rc = N.Send(base, rowBytes*rows, kNcclInt8, root, h->comm, st);
			}
		}
	}
	const int rcEnd = N.GroupEnd();
	if(rc != 0) return ncclFail(h, rc, "ncclSend/ncclRecv");
	if(rcEnd != 0) return ncclFail(h, rcEnd, "ncclGroupEnd");
	cudaEventRecord(h->ev[5], st);
	h->gatherPending = true;
	return AQH_OK;
}

} // extern "C"
