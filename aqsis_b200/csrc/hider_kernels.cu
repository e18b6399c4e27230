// hider_kernels.cu -- sm_100a kernels of the REYES hider + pixel filter.
//
// Compile with -fmad=false: every float operation below must round exactly like the
// reference's x86-64 build, which has no fused multiply-add (SURVEY.md section 7,
// "bit-stable parity without FMA").  Divisions and square roots are the IEEE ones
// (nvcc defaults -prec-div=true -prec-sqrt=true; never --use_fast_math).
//
// Kernels (DESIGN.md has the data layout and the roofline of each):
//   k_project    Project_points: raw P -> raster x,y + camera z, vertex info bits, split lines
//                (CqMicroPolyGrid::Split, libs/core/micropolygon.cpp:684-749)
//   k_bin_*      Bust_grids + AddMPG: micropolygon bound -> tile lists
//                (micropolygon.cpp:770-856, imagebuffer.cpp:514-591)
//   k_hide       Prepare_bucket + Render_MPGs + Combine_samples for one tile per CTA
//                (bucketprocessor.cpp:95-206, 1067-1569; imagepixel.cpp:144-359)
//   k_filter     Filter_samples + ExposeBucket + quantise
//                (bucketprocessor.cpp:584-707, 766-806; ddmanager.cpp:1022-1118)
#include "hider_device.h"

#include <vector>
#include <cstring>
#include <algorithm>
#include <cmath>
#include <cfloat>

namespace aqh {

// ------------------------------------------------------------------------------------
// scalar helpers (include/aqsis/math/math.h:47-125)
__device__ __forceinline__ float minA(float a, float b) { return (a < b) ? a : b; }
__device__ __forceinline__ float maxA(float a, float b) { return (a < b) ? b : a; }
__device__ __forceinline__ int lceilF(float x) { int i = (int)x; return i + ((x > 0.f && x != (float)i) ? 1 : 0); }
__device__ __forceinline__ int lfloorF(float x) { int i = (int)x; return i - ((x < 0.f && x != (float)i) ? 1 : 0); }
__device__ __forceinline__ int floorI(float x) { return (int)floorf(x); }

// order-preserving float -> uint map for the packed (depth, submission order) keys
__device__ __forceinline__ uint32_t depthKey(float d)
{
	uint32_t b = __float_as_uint(d + 0.0f);   // -0 -> +0
	return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float keyDepth(uint32_t k)
{
	uint32_t b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
	return __uint_as_float(b);
}
#define KEY_EMPTY_HI 0xff7fffffu            /* depthKey(FLT_MAX): occlZ of a cleared sample */
#define KEY_EMPTY ((((unsigned long long)KEY_EMPTY_HI) << 32) | 0xffffffffull)

struct B2 { float mnx, mny, mnz, mxx, mxy, mxz; };   // CqBound

__device__ __forceinline__ B2 boundOf4(const float4& a, const float4& b, const float4& c, const float4& d)
{
	B2 r;
	r.mnx = minA(a.x, minA(b.x, minA(c.x, d.x))); r.mny = minA(a.y, minA(b.y, minA(c.y, d.y)));
	r.mnz = minA(a.z, minA(b.z, minA(c.z, d.z)));
	r.mxx = maxA(a.x, maxA(b.x, maxA(c.x, d.x))); r.mxy = maxA(a.y, maxA(b.y, maxA(c.y, d.y)));
	r.mxz = maxA(a.z, maxA(b.z, maxA(c.z, d.z)));
	return r;
}
__device__ __forceinline__ void encapsulate(B2& a, const B2& b)
{
	a.mxx = maxA(a.mxx, b.mxx); a.mxy = maxA(a.mxy, b.mxy); a.mxz = maxA(a.mxz, b.mxz);
	a.mnx = minA(a.mnx, b.mnx); a.mny = minA(a.mny, b.mny); a.mnz = minA(a.mnz, b.mnz);
}

// CqRenderer::GetCircleOfConfusion, renderer.h:401-406
__device__ __forceinline__ float2 cocAt(const DevFrame& f, float depth)
{
	float c = f.dofMult * fabsf(1.0f / depth - f.dofInvFocal);
	return make_float2(f.dofScaleX * c, f.dofScaleY * c);
}
// CqRenderer::MinCoCForBound, renderer.cpp:1602-1617
__device__ __forceinline__ float minCoCForBound(const DevFrame& f, float z1, float z2)
{
	float focalDist = 1.0f / f.dofInvFocal;
	if((z1 - focalDist)*(z2 - focalDist) < 0.f)
		return 0.f;
	float minBlur = minA(fabsf(1.0f/z1 - f.dofInvFocal), fabsf(1.0f/z2 - f.dofInvFocal));
	return f.dofMult * minA(f.dofScaleX, f.dofScaleY) * minBlur;
}

__device__ __forceinline__ uint32_t infoOf(const float4& v) { return __float_as_uint(v.w); }

// Does the micropolygon at position p use the sample's occluding slot?  isOpaque (all shading corners Oi >= 1: four
// with smooth shading, the first otherwise -- points are always constant, micropolygon.cpp:1482,1518-1521) or MatteAlpha,
// and cullable: not part of a CSG solid and the frame keeps no full hit lists for the depth filter
// (bucketprocessor.cpp:1074-1079, 1490-1491).
__device__ __forceinline__ bool mpCullable(const DevFrame& f, const GridRec& g)
{
	return f.cullable && !(g.flags & AQH_GRID_USES_CSG);
}
__device__ __forceinline__ bool mpOpaqueSlot(const DevFrame& f, const GridRec& g, uint32_t p, uint32_t info)
{
	bool opaque = (info & VINFO_OPAQUE) != 0;
	if((g.flags & AQH_GRID_SMOOTH) && !(g.flags & AQH_GRID_POINTS))
	{
		const uint32_t cu = g.cu_cv & 0xffffu;
		opaque = opaque && (infoOf(f.P4[p+1]) & VINFO_OPAQUE) && (infoOf(f.P4[p+cu+1]) & VINFO_OPAQUE) && (infoOf(f.P4[p+cu+2]) & VINFO_OPAQUE);
	}
	return (opaque || (g.flags & AQH_GRID_MATTE_ALPHA)) && mpCullable(f, g);
}

// ------------------------------------------------------------------------------------
// Trim curves: CqTrimLoopArray::TrimPoint / LineIntersects over the tessellated loops of one surface
// (geometry/trimcurve.cpp:145-242).  Rare (trimmed NURBS patches only): out of line.
__device__ __noinline__ bool trimPointSet(const int32_t* setLoop, const int32_t* loopPoint, const float2* pts, int set, float x, float y)
{
	const int l0 = setLoop[set], l1 = setLoop[set + 1];
	if(l1 == l0) return false;                       // no trim loops at all
	int cCrosses = 0;
	for(int l = l0; l < l1; ++l)
	{
		const int p0 = loopPoint[l], size = loopPoint[l + 1] - p0;
		bool oddNodes = false;
		for(int i = 0, j = size - 1; i < size; j = i++)
		{
			const float2 a = pts[p0 + i], b = pts[p0 + j];
			// does the segment span the point in y, and does its crossing lie on the low-x side of the point?
			if(((a.y < y) && (b.y >= y)) || ((b.y < y) && (a.y >= y)))
				if(a.x + (y - a.y) / (b.y - a.y) * (b.x - a.x) < x) oddNodes = !oddNodes;
		}
		cCrosses += oddNodes ? 1 : 0;
	}
	return !(cCrosses & 1);
}
__device__ __noinline__ bool trimLineSet(const int32_t* setLoop, const int32_t* loopPoint, const float2* pts, int set, float2 v1, float2 v2)
{
	const int l0 = setLoop[set], l1 = setLoop[set + 1];
	const float x1 = v1.x, y1 = v1.y, x2 = v2.x, y2 = v2.y;
	for(int l = l0; l < l1; ++l)
	{
		const int p0 = loopPoint[l], size = loopPoint[l + 1] - p0;
		for(int i = 0, j = size - 1; i < size; j = i++)
		{
			const float x3 = pts[p0 + i].x, y3 = pts[p0 + i].y, x4 = pts[p0 + j].x, y4 = pts[p0 + j].y;
			const float d = (x2-x1)*(y4-y3) - (y2-y1)*(x4-x3);
			if(d == 0.0f) continue;
			const float r = ((y1-y3)*(x4-x3) - (x1-x3)*(y4-y3)) / d;
			const float s = ((y1-y3)*(x2-x1) - (x1-x3)*(y2-y1)) / d;
			if((r >= 0.0f) && (s >= 0.0f) && (r <= 1.0f) && (s <= 1.0f)) return true;
		}
	}
	return false;
}
// BilinearEvaluate (libs/core/bilinear.h:190-224), one component
__device__ __forceinline__ float bilinearEvaluate(float A, float B, float C, float D, float s, float t)
{
	float AB, CD;
	if(s <= 0.0f) { AB = A; CD = C; }
	else if(s >= 1.0f) { AB = B; CD = D; }
	else { AB = (B - A)*s + A; CD = (D - C)*s + C; }
	if(t <= 0.0f) return AB;
	if(t >= 1.0f) return CD;
	return (CD - AB)*t + AB;
}
// The per-hit test of a micropolygon a trim curve crosses (CqMicroPolygon::Sample, micropolygon.cpp:1594-1628): the hit's
// surface parameters, interpolated over the four corners, lie in the trimmed-away region -- and the sense is "inside".
__device__ __noinline__ bool trimRejectHit(const int32_t* gridTrim, const int32_t* setLoop, const int32_t* loopPoint, const float2* pts,
                                           const float2* trimUV, uint32_t gi, uint32_t gflags, size_t v0, uint32_t cu, float u, float v)
{
	if(gflags & AQH_GRID_TRIM_OUTSIDE) return false;
	const int set = gridTrim[gi];
	if(set == 0) return false;      // bCanBeTrimmed (a NURBS surface without loops can be trimmed too: TrimPoint is false everywhere)
	const float2 A = trimUV[v0], B = trimUV[v0 + 1], C = trimUV[v0 + cu + 1], D = trimUV[v0 + cu + 2];
	const float rx = bilinearEvaluate(A.x, B.x, C.x, D.x, u, v), ry = bilinearEvaluate(A.y, B.y, C.y, D.y, u, v);
	return trimPointSet(setLoop, loopPoint, pts, set, rx, ry);
}

// ------------------------------------------------------------------------------------
// k_project: one thread per position.
__global__ void __launch_bounds__(256) k_project(const __grid_constant__ DevFrame f, int64_t pA, int64_t pB)
{
	const int64_t i = pA + (int64_t)blockIdx.x * 256 + threadIdx.x;
	if(i >= pB) return;
	// locate the grid: bounded binary search between the grids of this and the next 256-position chunk
	uint32_t lo = f.chunkGrid[i >> 8], hi = f.chunkGrid[(i >> 8) + 1];
	while(lo < hi)
	{
		uint32_t mid = (lo + hi + 1) >> 1;
		if((int64_t)f.grids[mid].pbase <= i) lo = mid; else hi = mid - 1;
	}
	const GridRec g = f.grids[lo];
	const uint32_t local = (uint32_t)(i - g.pbase);
	const uint32_t key = local / g.nverts;
	const uint32_t v = local - key*g.nverts;
	float x = f.Praw[3*i], y = f.Praw[3*i+1], z = f.Praw[3*i+2];
	if(g.flags & AQH_GRID_CAMERA_SPACE)
	{
		// CqMatrix::operator*(CqVector3D), include/aqsis/math/matrix.h:717-750
		const float* m = f.camToRaster;
		float h = (m[3]*x + m[7]*y + m[11]*z + m[15]);
		float rx = (m[0]*x + m[4]*y + m[8]*z + m[12]);
		float ry = (m[1]*x + m[5]*y + m[9]*z + m[13]);
		if(h != 1.f)
		{
			float invh = 1.f/h;
			rx = rx*invh; ry = ry*invh;
		}
		x = rx; y = ry;   // z keeps the camera depth (micropolygon.cpp:727-729)
	}
	uint32_t info = lo & VINFO_GRID_MASK;
	if(key == 0)
	{
		const uint32_t cu = g.cu_cv & 0xffffu, cv = g.cu_cv >> 16;
		const uint32_t iv = v / (cu + 1), iu = v - iv*(cu + 1);
		const uint32_t vid = g.vbase + v;
		float ci0 = 1.0f, ci1 = 1.0f, ci2 = 1.0f, oi0 = 1.0f, oi1 = 1.0f, oi2 = 1.0f;
		if(f.Ci) { ci0 = f.Ci[3*(size_t)vid]; ci1 = f.Ci[3*(size_t)vid+1]; ci2 = f.Ci[3*(size_t)vid+2]; }
		if(f.Oi) { oi0 = f.Oi[3*(size_t)vid]; oi1 = f.Oi[3*(size_t)vid+1]; oi2 = f.Oi[3*(size_t)vid+2]; }
		// packed copy for the hit shading (two 16-byte loads per corner instead of six scalar ones)
		f.CO[2*(size_t)vid] = make_float4(ci0, ci1, ci2, oi0);
		f.CO[2*(size_t)vid+1] = make_float4(oi1, oi2, 0.f, 0.f);
		const bool opaque = (oi0 >= 1.0f) && (oi1 >= 1.0f) && (oi2 >= 1.0f);
		// every vertex of a points grid is a micropolygon and CqMicroPolyGridPoints::Split ignores the culled flags
		bool valid = (g.flags & AQH_GRID_POINTS) ? true : ((iu < cu) && (iv < cv) && !(f.culled && f.culled[vid]));
		if((g.flags & AQH_GRID_CULL_BACKFACING) && f.Ng && !(g.flags & (AQH_GRID_USES_CSG | AQH_GRID_POINTS)))
		{
			// Sides 1 (micropolygon.cpp:431-474): ((s * Ng) . P) >= 0 with P in camera space; s turns Ng to the side of a user normal
			const float px_ = f.Praw[3*i], py_ = f.Praw[3*i+1], pz_ = f.Praw[3*i+2];
			const float gx = f.Ng[3*(size_t)vid], gy = f.Ng[3*(size_t)vid+1], gz = f.Ng[3*(size_t)vid+2];
			float sgn = 1.0f;
			if(f.Nn)
			{
				const float nx = f.Nn[3*(size_t)vid], ny = f.Nn[3*(size_t)vid+1], nz = f.Nn[3*(size_t)vid+2];
				sgn = ((nx*gx + ny*gy + nz*gz) < 0.0f) ? -1.0f : 1.0f;
			}
			if((((sgn*gx)*px_) + ((sgn*gy)*py_) + ((sgn*gz)*pz_)) >= 0.f) valid = false;
		}
		if((g.flags & AQH_GRID_CULL_TRANSPARENT) && f.Oi && f.cullTransparentOk && !(g.flags & AQH_GRID_POINTS))
		{
			// fully transparent micropolygons are dropped from the last shading point down to the first one that is not
			// black (micropolygon.cpp:493-522): remember the last non-black vertex, k_bin drops what lies behind it
			if(!(oi0 == 0.f && oi1 == 0.f && oi2 == 0.f)) atomicMax(&f.gridTail[lo], v + 1u);
		}
		if(valid && f.anyTrim && !(g.flags & AQH_GRID_POINTS))
		{
			// micropolygon.cpp:784-835: a micropolygon whose four corners are all trimmed away and whose edges no trim curve
			// crosses is dropped; one with any corner trimmed away has its hits tested (MarkTrimmed)
			const int set = f.gridTrim[lo];
			if(set != 0)
			{
				const float2 A = f.trimUV[vid], B = f.trimUV[vid + 1], C = f.trimUV[vid + cu + 2], D = f.trimUV[vid + cu + 1];
				const bool outside = (g.flags & AQH_GRID_TRIM_OUTSIDE) != 0;
				const bool tA = trimPointSet(f.trimSetLoop, f.trimLoopPoint, f.trimPoints, set, A.x, A.y) != outside;
				const bool tB = trimPointSet(f.trimSetLoop, f.trimLoopPoint, f.trimPoints, set, B.x, B.y) != outside;
				const bool tC = trimPointSet(f.trimSetLoop, f.trimLoopPoint, f.trimPoints, set, C.x, C.y) != outside;
				const bool tD = trimPointSet(f.trimSetLoop, f.trimLoopPoint, f.trimPoints, set, D.x, D.y) != outside;
				if(tA && tB && tC && tD)
				{
					if(!trimLineSet(f.trimSetLoop, f.trimLoopPoint, f.trimPoints, set, A, B) &&
					   !trimLineSet(f.trimSetLoop, f.trimLoopPoint, f.trimPoints, set, B, C) &&
					   !trimLineSet(f.trimSetLoop, f.trimLoopPoint, f.trimPoints, set, C, D) &&
					   !trimLineSet(f.trimSetLoop, f.trimLoopPoint, f.trimPoints, set, D, A)) valid = false;
				}
				if(valid && (tA || tB || tC || tD)) info |= VINFO_MP_TRIMMED;
			}
		}
		if(opaque) info |= VINFO_OPAQUE;
		if(valid) info |= VINFO_MP_VALID;
		// frame-level "there is transparency" flag: read before the atomic so that only the first
		// few non-opaque vertices pay for one
		if(!opaque && !(*(volatile uint32_t*)f.errorFlags & 2u)) atomicOr(f.errorFlags, 2u);
	}
	f.P4[i] = make_float4(x, y, z, __uint_as_float(info));
}

// Triangle split line per (grid, key): micropolygon.cpp:733-749.  One thread per grid.
__global__ void __launch_bounds__(128) k_splitlines(const __grid_constant__ DevFrame f)
{
	const int gi = blockIdx.x*128 + threadIdx.x;
	if(gi >= f.nGrids) return;
	const GridRec g = f.grids[gi];
	const uint32_t cu = g.cu_cv & 0xffffu, cv = g.cu_cv >> 16;
	const uint32_t nkeys = g.nkeys_koff & 0xffu, koff = g.nkeys_koff >> 8;
	for(uint32_t k = 0; k < nkeys; ++k)
	{
		const float4* P = f.P4 + g.pbase + (size_t)k*g.nverts;
		float4 v0 = P[0], v1 = P[cu], v2 = P[cv*(cu+1)];
		float4 sl;
		if(((v1.x - v0.x)*(v2.y - v0.y) - (v1.y - v0.y)*(v2.x - v0.x)) >= 0.f)
			sl = make_float4(v1.x, v1.y, v2.x, v2.y);
		else
			sl = make_float4(v2.x, v2.y, v1.x, v1.y);
		f.splitLines[koff + k] = sl;
	}
}

// ------------------------------------------------------------------------------------
// Binning.  The tile range of a micropolygon only has to be a SUPERSET of the tiles whose
// samples it can touch (the hide kernel repeats the reference's exact tests), so the bound
// is padded a little to stay conservative under the lerp rounding of the motion sub-bounds.
struct TileRange { int tx0, tx1, ty0, ty1; };   // inclusive

__device__ __forceinline__ bool mpTileRange(const DevFrame& f, int64_t p, const float4& a, TileRange& tr, uint32_t& zminKey)
{
	const uint32_t info = infoOf(a);
	if(!(info & VINFO_MP_VALID)) return false;
	const GridRec g = f.grids[info & VINFO_GRID_MASK];
	if((uint64_t)(p - g.pbase) >= g.nverts) return false;            // only key 0 starts micropolygons
	// incremental flushes: a flush (z only) hides just the micropolygons that can occlude; the final frame has the
	// opaque ones of flushed grids in the stored occlusion keys already
	if(f.zOnly && f.zKeys) { if(!mpOpaqueSlot(f, g, (uint32_t)p, info)) return false; }
	else if(p < f.flushedPos) { if(mpOpaqueSlot(f, g, (uint32_t)p, info)) return false; }
	const uint32_t cu = g.cu_cv & 0xffffu;
	const uint32_t nkeys = g.nkeys_koff & 0xffu;
	if((g.flags & AQH_GRID_CULL_TRANSPARENT) && f.Oi && f.cullTransparentOk && !(g.flags & AQH_GRID_POINTS))
		if((uint32_t)(p - g.pbase) >= f.gridTail[info & VINFO_GRID_MASK]) return false;      // the trailing run of Oi == 0 vertices
	B2 B;
	if(g.flags & AQH_GRID_POINTS)
	{
		// CqMicroPolygonPoints::Initialise, geometry/points.h:350-359: position +- radius, flat in z
		const float r = f.radius[p];
		B.mnx = a.x - r; B.mny = a.y - r; B.mnz = a.z - 0.f; B.mxx = a.x + r; B.mxy = a.y + r; B.mxz = a.z + 0.f;
	}
	else
	{
		B = boundOf4(a, f.P4[p+1], f.P4[p+cu+1], f.P4[p+cu+2]);
		for(uint32_t k = 1; k < nkeys; ++k)
		{
			const float4* Pk = f.P4 + p + (size_t)k*g.nverts;
			B2 kb = boundOf4(Pk[0], Pk[1], Pk[cu+1], Pk[cu+2]);
			encapsulate(B, kb);
		}
	}
	zminKey = depthKey(B.mnz);
	if(f.useDof)
	{
		// imagebuffer.cpp:519-531
		float2 c1 = cocAt(f, B.mnz), c2 = cocAt(f, B.mxz);
		float mcx = maxA(c1.x, c2.x), mcy = maxA(c1.y, c2.y);
		B.mnx -= mcx; B.mny -= mcy; B.mxx += mcx; B.mxy += mcy;
	}
	const float pad = 1.0f/128.0f;
	float x0 = B.mnx - pad - fabsf(B.mnx)*1e-6f, x1 = B.mxx + pad + fabsf(B.mxx)*1e-6f;
	float y0 = B.mny - pad - fabsf(B.mny)*1e-6f, y1 = B.mxy + pad + fabsf(B.mxy)*1e-6f;
	// pixel range [floor(min), ceil(max)) clamped to the sample region, in float to survive huge bounds
	float fx0 = fmaxf(floorf(x0), (float)f.sx0), fx1 = fminf(ceilf(x1), (float)(f.sx0 + f.sw));
	float fy0 = fmaxf(floorf(y0), (float)f.sy0), fy1 = fminf(ceilf(y1), (float)(f.sy0 + f.sh));
	if(!(fx0 < fx1) || !(fy0 < fy1)) return false;
	tr.tx0 = ((int)fx0 - f.sx0) / f.tileW; tr.tx1 = ((int)fx1 - 1 - f.sx0) / f.tileW;
	tr.ty0 = ((int)fy0 - f.sy0) / f.tileH; tr.ty1 = ((int)fy1 - 1 - f.sy0) / f.tileH;
	return true;
}

template<bool FILL>
__global__ void __launch_bounds__(256) k_bin(const __grid_constant__ DevFrame f, int64_t pA, int64_t pB)
{
	const int64_t p = pA + (int64_t)blockIdx.x * 256 + threadIdx.x;
	TileRange tr;
	uint32_t zminKey = 0;
	bool live = p < pB;
	if(live)
	{
		const float4 a = f.P4[p];
		live = mpTileRange(f, p, a, tr, zminKey);
	}
	unsigned nent = 0;
	unsigned long long entryHi = 0;
	if(FILL && live)
	{
		// bit 63: the micropolygon goes to the transparent pass; bits 32-62: its depth key without the last bit (the order only
		// steers the culling)
		// (frames of the static kernel only: binPartition)
		if(f.binPartition)
		{
			const uint32_t info = infoOf(f.P4[p]);
			const bool deep = f.anyTransparent && !mpOpaqueSlot(f, f.grids[info & VINFO_GRID_MASK], (uint32_t)p, info);
			entryHi = ((unsigned long long)((zminKey >> 1) | (deep ? 0x80000000u : 0u))) << 32;
		}
		else entryHi = (unsigned long long)zminKey << 32;
	}
	if(live)
	for(int ty = tr.ty0; ty <= tr.ty1; ++ty)
		for(int tx = tr.tx0; tx <= tr.tx1; ++tx)
		{
			const int slot = f.tileSlot[ty*f.ntx + tx];
			if(slot < 0) continue;
			if(FILL)
			{
				uint32_t at = atomicAdd(&f.binCount[slot], 1u);
				// (does not use the opaque slot, nearest depth of the micropolygon, position index): sorted per tile by
				// k_bin_sort -- the opaque micropolygons first, each group front to back
				f.binEntries[f.binOffset[slot] + at] = entryHi | (uint32_t)p;
			}
			else
			{
				atomicAdd(&f.binCount[slot], 1u);
				++nent;
			}
		}
	if(!FILL)
	{
		// statistics: one pair of global atomics per CTA (every warp of the frame hitting the same two
		// words serialises in L2)
		__shared__ unsigned s_nmp, s_tot;
		if(threadIdx.x == 0) { s_nmp = 0; s_tot = 0; }
		__syncthreads();
		unsigned nmp = __popc(__ballot_sync(0xffffffffu, nent != 0));
		unsigned tot = __reduce_add_sync(0xffffffffu, nent);
		if((threadIdx.x & 31) == 0 && tot) { atomicAdd(&s_nmp, nmp); atomicAdd(&s_tot, tot); }
		__syncthreads();
		if(threadIdx.x == 0 && s_tot)
		{
			atomicAdd(&f.counters[0], (unsigned long long)s_nmp);
			atomicAdd(&f.counters[1], (unsigned long long)s_tot);
		}
	}
}

// Exclusive scan of binCount -> binOffset (single CTA; the tile count is small), and reset
// of binCount for the fill pass.
__global__ void __launch_bounds__(1024) k_bin_scan(const __grid_constant__ DevFrame f)
{
	__shared__ uint32_t s_part[1024];
	const int n = f.nActiveTiles;
	const int per = (n + 1023) / 1024;
	const int beg = threadIdx.x * per, end = min(beg + per, n);
	uint32_t sum = 0, mx = 0;
	for(int i = beg; i < end; ++i) { sum += f.binCount[i]; mx = max(mx, f.binCount[i]); }
	mx = __reduce_max_sync(0xffffffffu, mx);
	if((threadIdx.x & 31) == 0 && mx) atomicMax(&f.counters[3], (unsigned long long)mx);   // longest bin: sizes the sort
	s_part[threadIdx.x] = sum;
	__syncthreads();
	// Hillis-Steele inclusive scan over the 1024 partials
	for(int d = 1; d < 1024; d <<= 1)
	{
		uint32_t v = (threadIdx.x >= d) ? s_part[threadIdx.x - d] : 0;
		__syncthreads();
		s_part[threadIdx.x] += v;
		__syncthreads();
	}
	uint32_t run = (threadIdx.x == 0) ? 0 : s_part[threadIdx.x - 1];
	for(int i = beg; i < end; ++i)
	{
		uint32_t c = f.binCount[i];
		f.binOffset[i] = run;
		run += c;
		f.binCount[i] = 0;
	}
	if(threadIdx.x == 1023) f.binOffset[n] = s_part[1023];
}

// Order every tile's bin front to back (ascending nearest depth, ties by submission order).
// The hide kernel's results do not depend on the order -- every hit competes through an
// order-independent (depth, submission) key -- but its SPEED does: with near micropolygons
// first the reference's own occlusion cull (Bound.zmin > occlZ, bucketprocessor.cpp:1179)
// rejects hidden ones before their edge tests.  This mirrors the reference, which renders a
// bucket's surfaces nearest first (CqBucket::closest_surface, bucket.h).
// One CTA per tile; bitonic sort of up to SORT_MAX entries in shared memory; longer bins are
// sorted in SORT_MAX-sized runs.
#define SORT_MAX 8192
__global__ void __launch_bounds__(256) k_bin_sort(const __grid_constant__ DevFrame f, int sortMax)
{
	extern __shared__ unsigned long long s_keys[];
	const int slot = blockIdx.x;
	const uint32_t beg = f.binOffset[slot], end = f.binOffset[slot+1];
	for(uint32_t c0 = beg; c0 < end; c0 += sortMax)
	{
		const int cnt = (int)min((uint32_t)sortMax, end - c0);
		if(cnt < 2) break;
		int n = 32;
		while(n < cnt) n <<= 1;
		for(int i = threadIdx.x; i < n; i += 256) s_keys[i] = (i < cnt) ? f.binEntries[c0 + i] : ~0ull;
		__syncthreads();
		for(int k = 2; k <= n; k <<= 1)
			for(int j = k >> 1; j > 0; j >>= 1)
			{
				for(int t = threadIdx.x; t < (n >> 1); t += 256)
				{
					const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));     // index with bit j clear
					const int l = i | j;
					const unsigned long long a = s_keys[i], b = s_keys[l];
					const bool up = (i & k) == 0;
					if((a > b) == up) { s_keys[i] = b; s_keys[l] = a; }
				}
				__syncthreads();
			}
		for(int i = threadIdx.x; i < cnt; i += 256) f.binEntries[c0 + i] = s_keys[i];
		__syncthreads();
	}
}

// ------------------------------------------------------------------------------------
// Micropolygon hit-test pieces.

// Vertex order of a micropolygon (what CqMicroPolygon::ComputeVertexOrder decides, micropolygon.cpp:1207-1284): which
// corner to drop when an edge of the loop  slot 0 -> 1 -> 3 -> 2 -> 0  has collapsed (squared length, z included, below
// 1e-8), and in which direction to walk the remaining corners so that the edge tests see a consistent winding.  Returns
// the packed code the reference keeps in m_IndexCode: two bits per corner slot, first corner in bits 0-1, plus
// DEGENERACY_MASK for triangles.
__device__ __forceinline__ float mag2d3(const float4& a, const float4& b)
{
	float dx = a.x-b.x, dy = a.y-b.y, dz = a.z-b.z;
	return dx*dx + dy*dy + dz*dz;
}
#define DEGENERACY_MASK 0x8000000
// corner i of four values held in registers (a dynamically indexed array would live in local memory)
__device__ __forceinline__ float pick4(const float v[4], int i) { return i == 0 ? v[0] : (i == 1 ? v[1] : (i == 2 ? v[2] : v[3])); }
__device__ __forceinline__ int computeVertexOrder(const float4 P[4])
{
	// (double)m < 1e-8 for a float m  <=>  m <= the largest float below 1e-8 (no double arithmetic needed)
	const float tiny = 9.99999993922529029e-9f;
	// corner sequences, one nibble per position (0xF = none): the full loop, or the loop without the corner that
	// coincides with its predecessor -- edges 0-1 and 1-3 both leave {0,3,2}, edge 3-2 leaves {0,1,2}, edge 2-0 leaves {0,1,3}
	uint32_t seq = 0x2310u;
	if(mag2d3(P[0], P[1]) <= tiny) seq = 0xF230u;
	else if(mag2d3(P[1], P[3]) <= tiny) seq = 0xF230u;
	else if(mag2d3(P[3], P[2]) <= tiny) seq = 0xF210u;
	else if(mag2d3(P[2], P[0]) <= tiny) seq = 0xF310u;
	const int s0 = seq & 3, s1 = (seq >> 4) & 3, s2 = (seq >> 8) & 3;
	const bool triangle = (seq >> 12) == 0xFu;
	const int s3 = (seq >> 12) & 3;
	const float X[4] = {P[0].x, P[1].x, P[2].x, P[3].x}, Y[4] = {P[0].y, P[1].y, P[2].y, P[3].y};
	// the first corner of every sequence is corner 0
	const float v0x = X[0], v0y = Y[0], v1x = pick4(X, s1), v1y = pick4(Y, s1), v2x = pick4(X, s2), v2y = pick4(Y, s2);
	const bool forward = ((v0x - v1x)*(v1y - v2y)) >= ((v0y - v1y)*(v1x - v2x));
	if(triangle)
		return (forward ? (s0 | (s1 << 2) | (s2 << 4)) : (s0 | (s2 << 2) | (s1 << 4))) | DEGENERACY_MASK;
	return forward ? (s0 | (s1 << 2) | (s2 << 4) | (s3 << 6)) : (s0 | (s3 << 2) | (s2 << 4) | (s1 << 6));
}

// The hit-test cache of one micropolygon at one instant: CqHitTestCache minus the DoF
// members (micropolygon.h:520-551), filled by cachePointInPolyTest (micropolygon.cpp:1346-1392).
struct HitCache
{
	float X[4], Y[4], XM[4], YM[4];
	float Ax, Ay, Ex, Ey, Fx, Fy, Gx, Gy;
	float z[4];
	int linear;
};

__device__ __forceinline__ void cachePointInPolyTest(HitCache& c, const float px[4], const float py[4], const float pz[4], int code)
{
	c.z[0] = pz[0]; c.z[1] = pz[1]; c.z[2] = pz[2]; c.z[3] = pz[3];
	// CqInvBilinear::setVertices(A,B,C,D), bilinear.h:260-275
	c.Ax = px[0]; c.Ay = py[0];
	c.Ex = px[1] - px[0]; c.Ey = py[1] - py[0];
	c.Fx = px[2] - px[0]; c.Fy = py[2] - py[0];
	c.Gx = -c.Ex - px[2] + px[3]; c.Gy = -c.Ey - py[2] + py[3];
	float patchSize = maxA(maxA(fabsf(c.Fx), fabsf(c.Fy)), maxA(fabsf(c.Ex), fabsf(c.Ey)));
	float irregularity = maxA(fabsf(c.Gx), fabsf(c.Gy));
	c.linear = ((double)irregularity < 1e-2*(double)patchSize) ? 1 : 0;
	const int i0 = (code >> 2) & 3, i1 = (code >> 4) & 3, i2 = (code >> 6) & 3, i3 = code & 3;
	const float qx[4] = {pick4(px, i0), pick4(px, i1), pick4(px, i2), pick4(px, i3)};
	const float qy[4] = {pick4(py, i0), pick4(py, i1), pick4(py, i2), pick4(py, i3)};
	int j = 3;
#pragma unroll
	for(int i = 0; i < 4; ++i)
	{
		c.YM[i] = qx[i] - qx[j];
		c.XM[i] = qy[i] - qy[j];
		c.X[i] = qx[j];
		c.Y[i] = qy[j];
		j = i;
	}
	if(code & DEGENERACY_MASK)
	{
#pragma unroll
		for(int i = 2; i < 4; ++i)
		{
			c.YM[i] = qx[3] - qx[1];
			c.XM[i] = qy[3] - qy[1];
			c.X[i] = qx[1];
			c.Y[i] = qy[1];
		}
	}
}

// The four edge tests of CqMicroPolygon::fContains (micropolygon.cpp:1306-1335): edges 0,1
// fail on <= 0, edges 2,3 on < 0.  The reference's m_LastFailedEdge only reorders the tests.
__device__ __forceinline__ bool edgeTests(const float* X, const float* Y, const float* XM, const float* YM, float x, float y)
{
	float e0 = ((y - Y[0]) * YM[0]) - ((x - X[0]) * XM[0]);
	float e1 = ((y - Y[1]) * YM[1]) - ((x - X[1]) * XM[1]);
	float e2 = ((y - Y[2]) * YM[2]) - ((x - X[2]) * XM[2]);
	float e3 = ((y - Y[3]) * YM[3]) - ((x - X[3]) * XM[3]);
	return !(e0 <= 0.f) && !(e1 <= 0.f) && !(e2 < 0.f) && !(e3 < 0.f);
}

// CqInvBilinear::operator(), bilinear.h:277-311
__device__ __forceinline__ float2 invBilinear(float Ax, float Ay, float Ex, float Ey, float Fx, float Fy,
                                              float Gx, float Gy, int linear, float Px, float Py)
{
	float u = 0.5f, v = 0.5f;
#pragma unroll
	for(int it = 0; it < 2; ++it)
	{
		if(it == 1 && linear) break;
		// bilinEval(uv) - P
		float bx = (Ax + Ex*u + Fx*v + Gx*u*v) - Px;
		float by = (Ay + Ey*u + Fy*v + Gy*u*v) - Py;
		float M1x = Ex + Gx*v, M1y = Ey + Gy*v;
		float M2x = Fx + Gx*u, M2y = Fy + Gy*u;
		float det = M1x*M2y - M1y*M2x;
		if(it == 0 || det != 0.f) det = 1.f/det;
		float sx = det * (bx*M2y - by*M2x);
		float sy = det * (-(bx*M1y - by*M1x));
		u -= sx; v -= sy;
	}
	return make_float2(u, v);
}
// bilerp, bilinear.h:229-236
__device__ __forceinline__ float bilerpZ(const float* z, float2 uv)
{
	float w0 = (1.f-uv.y)*(1.f-uv.x);
	float w1 = (1.f-uv.y)*uv.x;
	float w2 = uv.y*(1.f-uv.x);
	float w3 = uv.y*uv.x;
	return w0*z[0] + w1*z[1] + w2*z[2] + w3*z[3];
}

// CqMotionSpec::GetMotionObjectInterpolated on the split line (motion.h:176-228,
// micropolygon.h:182-188) + the side test of micropolygon.cpp:1630-1654.
__device__ __noinline__ bool triangleSplitRejectImpl(const float4* sl, const float* times, uint32_t nkeys,
                                                     float posx, float posy, float dofx, float dofy, float D, float time,
                                                     int useDof, float dofMult, float dofInvFocal, float dofScaleX, float dofScaleY)
{
	float4 s;
	if(nkeys == 1) s = sl[0];
	else if(time >= times[nkeys-1]) s = sl[nkeys-1];
	else if(time <= times[0]) s = sl[0];
	else
	{
		uint32_t i = 0;
		while(time >= times[i+1]) i += 1;
		float Fr = (time - times[i]) / (times[i+1] - times[i]);
		if(times[i] == time) s = sl[i];
		else
		{
			float4 a = sl[i], b = sl[i+1];
			s = make_float4(((1.0f - Fr)*a.x) + (Fr*b.x), ((1.0f - Fr)*a.y) + (Fr*b.y),
			                ((1.0f - Fr)*a.z) + (Fr*b.z), ((1.0f - Fr)*a.w) + (Fr*b.w));
		}
	}
	float Ax = s.x, Ay = s.y, Bx = s.z, By = s.w;
	float hx = posx, hy = posy;
	if(useDof)
	{
		// GetCircleOfConfusion(D), renderer.h:401-406
		float c = dofMult * fabsf(1.0f / D - dofInvFocal);
		hx += (dofScaleX * c)*dofx; hy += (dofScaleY * c)*dofy;
	}
	float v = (Ay - By)*hx + (Bx - Ax)*hy + (Ax*By - Bx*Ay);
	return v <= 0.f;
}
// Rare (polygon grids only): kept out of line so that the sampling loops stay small in the instruction cache.
__device__ __forceinline__ bool triangleSplitReject(const DevFrame& f, const GridRec& g, float2 pos, float2 dofOff, float D, float time)
{
	const uint32_t nkeys = g.nkeys_koff & 0xffu, koff = g.nkeys_koff >> 8;
	return triangleSplitRejectImpl(f.splitLines + koff, f.keyTimes + koff, nkeys, pos.x, pos.y, dofOff.x, dofOff.y, D, time,
	                               f.useDof, f.dofMult, f.dofInvFocal, f.dofScaleX, f.dofScaleY);
}

// CqMicroPolygon::InterpolateOutputs + CacheOutputInterpCoeffs*, micropolygon.cpp:1443-1529.
// v0: the micropolygon's first vertex in the Ci/Oi arrays; smooth: bilinear over the four corners, else the first corner.
__device__ __forceinline__ void shadeAt(const DevFrame& f, size_t v0, uint32_t cu, bool smooth, float2 uv, float* col, float* opa)
{
	if(smooth)
	{
		const float4 a0 = f.CO[2*v0], a1 = f.CO[2*v0+1];
		const float4 b0 = f.CO[2*(v0+1)], b1 = f.CO[2*(v0+1)+1];
		const float4 c0 = f.CO[2*(v0+cu+1)], c1 = f.CO[2*(v0+cu+1)+1];
		const float4 d0 = f.CO[2*(v0+cu+2)], d1 = f.CO[2*(v0+cu+2)+1];
		float w[4];
		w[0] = (1.f-uv.x)*(1.f-uv.y);
		w[1] = uv.x*(1.f-uv.y);
		w[2] = (1.f-uv.x)*uv.y;
		w[3] = uv.x*uv.y;
		col[0] = w[0]*a0.x + w[1]*b0.x + w[2]*c0.x + w[3]*d0.x;
		col[1] = w[0]*a0.y + w[1]*b0.y + w[2]*c0.y + w[3]*d0.y;
		col[2] = w[0]*a0.z + w[1]*b0.z + w[2]*c0.z + w[3]*d0.z;
		opa[0] = w[0]*a0.w + w[1]*b0.w + w[2]*c0.w + w[3]*d0.w;
		opa[1] = w[0]*a1.x + w[1]*b1.x + w[2]*c1.x + w[3]*d1.x;
		opa[2] = w[0]*a1.y + w[1]*b1.y + w[2]*c1.y + w[3]*d1.y;
	}
	else
	{
		const float4 a0 = f.CO[2*v0], a1 = f.CO[2*v0+1];
		col[0] = a0.x; col[1] = a0.y; col[2] = a0.z;
		opa[0] = a0.w; opa[1] = a1.x; opa[2] = a1.y;
	}
}
__device__ __forceinline__ void shadeHit(const DevFrame& f, const GridRec& g, uint32_t p, float2 uv, float* col, float* opa)
{
	shadeAt(f, (size_t)g.vbase + (p - g.pbase), g.cu_cv & 0xffffu, (g.flags & AQH_GRID_SMOOTH) && !(g.flags & AQH_GRID_POINTS), uv, col, opa);
}
// What the resolve needs of a micropolygon's grid besides the position index, packed for the transparent hit records:
// first vertex in the Ci/Oi arrays, and cu | smooth shading << 16 | matte << 17.
#define HITF_SMOOTH (1u << 16)
#define HITF_MATTE (1u << 17)
__device__ __forceinline__ uint2 hitShadeInfo(const GridRec& g, uint32_t p)
{
	uint32_t fl = g.cu_cv & 0xffffu;
	if((g.flags & AQH_GRID_SMOOTH) && !(g.flags & AQH_GRID_POINTS)) fl |= HITF_SMOOTH;
	if(g.flags & AQH_GRID_MATTE) fl |= HITF_MATTE;
	return make_uint2(g.vbase + (p - g.pbase), fl);
}

// ------------------------------------------------------------------------------------
// Shared-memory layout of the hide kernel (dynamic).  The samples of a tile live in a padded
// SUB-SAMPLE GRID: row gy = py*ys + sy, column gx = px*xs + sx, idx = gy*stride + gx with
// stride = tileW*xs + SMEM_PAD, so that the candidate rectangle of a micropolygon is a plain 2-D
// window (no per-candidate table lookups) and consecutive lanes touch consecutive banks.
//   u64 keys[nsP] | f32 posx[nsP] | f32 posy[nsP] | [f32 time[nsP]] | [float2 dof[nsP]] | [u32 head[nsP]] |
//   StaticRec recs[nwarps*RECS_PER_WARP] | u32 pixZ[tileW*tileH] | u16 subOfs[n] | u8 shufPat[tileW*tileH]
#define SMEM_PAD 8
#define RECS_PER_WARP 8
struct StaticRec   // 40 words
{
	float X[4], Y[4], XM[4], YM[4];
	float bminx, bminy, bmaxx, bmaxy;
	float Ax, Ay, Ex, Ey, Fx, Fy, Gx, Gy;
	float z[4];
	uint32_t zminKey;
	uint32_t p;
	uint32_t flags;     // grid id | REC_*
	uint32_t rect;      // gx0 | gx1<<8 | gy0<<16 | gy1<<24  (tile sub-sample coordinates, <= 255)
	uint32_t v0, shade; // hitShadeInfo(): what a transparent hit record carries for the resolve
	uint32_t pad[2];
};
enum : uint32_t { REC_LINEAR = 1u << 28, REC_RARE = 1u << 29 /* a disc, or lod, triangular or trimmed: sampled by the out-of-line copy */, REC_POINT = 1u << 30 /* a disc: centre Ax,Ay radius Ex */,
                  REC_TRIM = 1u << 31 /* a trim curve crosses the micropolygon */ };

struct TileCtx
{
	int tileX0, tileY0;        // global pixel of the tile origin
	int rx0, ry0, rx1, ry1;    // tile ∩ sample region (global pixels)
};

struct MovScratch;
struct HideSmem
{
	MovScratch* mov;           // per-warp scratch of the motion blur / depth of field path (MBDOF kernel only)
	unsigned long long* keys;  // per sample (occlusion depth key << 32 | position index of the hit that set it)
	float* posx;
	float* posy;
	float* time;
	float2* dof;
	uint32_t* head;
	StaticRec* recs;
	uint16_t* subOfs;          // sample index i -> (i / xs)*stride + i % xs
	uint8_t* shufPat;
	uint32_t* pixZ;            // per pixel: upper bound of the depth keys of its samples (hierarchical z)
	uint32_t* tileZ;           // max of pixZ over the tile: nothing behind it can be visible in the tile
	uint32_t* dirty;           // set by every successful opaque store since the last refresh
	int stride, nsP;
};

__host__ __device__ __forceinline__ int hideStride(const DevFrame& f) { return f.tileW*f.xs + SMEM_PAD; }

__device__ __forceinline__ HideSmem carveSmem(const DevFrame& f, unsigned char* base, int nrecs, size_t movBytes)
{
	HideSmem s;
	s.stride = hideStride(f);
	s.nsP = f.tileH*f.ys*s.stride;
	const size_t ns = (size_t)s.nsP;
	size_t o = 0;
	// f.midpointZ: a second array right behind the first -- keys[idx] = occlusion key (second nearest opaque
	// depth), keys[nsP + idx] = nearest opaque hit (bucketprocessor.cpp:1502-1529); otherwise they are the same
	s.keys = (unsigned long long*)(base + o); o += ns*8*(f.midpointZ ? 2 : 1);
	s.posx = (float*)(base + o); o += ns*4;
	s.posy = (float*)(base + o); o += ns*4;
	s.dof = 0; s.time = 0; s.head = 0;
	if(f.useDof) { s.dof = (float2*)(base + o); o += ns*8; }
	if(f.anyMotion) { s.time = (float*)(base + o); o += ns*4; }
	if(f.anyTransparent) { s.head = (uint32_t*)(base + o); o += ns*4; }
	o = (o + 15) & ~(size_t)15;
	s.recs = (StaticRec*)(base + o); o += (size_t)nrecs*sizeof(StaticRec);
	s.mov = (MovScratch*)(base + o); o += movBytes;
	s.pixZ = (uint32_t*)(base + o); o += (size_t)f.tileW*f.tileH*4;
	s.subOfs = (uint16_t*)(base + o); o += (size_t)f.n*2;
	s.shufPat = (uint8_t*)(base + o);
	return s;
}

static size_t hideSmemBytes(const DevFrame& f, int nrecs, size_t movBytes)
{
	const size_t ns = (size_t)f.tileH*f.ys*hideStride(f);
	size_t o = ns*16 + (f.midpointZ ? ns*8 : 0);
	if(f.useDof) o += ns*8;
	if(f.anyMotion) o += ns*4;
	if(f.anyTransparent) o += ns*4;
	o = (o + 15) & ~(size_t)15;
	o += (size_t)nrecs*sizeof(StaticRec) + movBytes;
	o += (size_t)f.tileW*f.tileH*4;
	o += (size_t)f.n*2 + (size_t)f.tileW*f.tileH;
	return (o + 15) & ~(size_t)15;
}

// sample index of (pixel-local x, y, i) in the padded sub-sample grid
__device__ __forceinline__ int sampleIdx(const DevFrame& f, const HideSmem& s, int plx, int ply, int i)
{
	return (ply*f.ys)*s.stride + plx*f.xs + s.subOfs[i];
}

// Deposit a hit: opaque hits race for the per-sample (depth, order) minimum; transparent ones
// are appended to the CTA's deep pool if they are in front of the final opaque depth.
// StoreSample, bucketprocessor.cpp:1471-1569.
template<bool PL = false>
__device__ __forceinline__ void storeOpaque(const DevFrame& f, const HideSmem& s, unsigned long long* key, float D, uint32_t p)
{
	if(!(D < FLT_MAX)) return;                       // occlZ(=FLT_MAX) <= D
	unsigned long long nk = ((unsigned long long)depthKey(D) << 32) | p;
	if(nk < *key)                                    // cheap pre-check; keys only ever decrease
	{
		if(!PL && f.midpointZ)
		{
			// midpoint depth filter with a z display (bucketprocessor.cpp:1502-1529): the sample keeps the
			// nearest hit and occlZ = the second nearest depth.  Order independent: the value that loses
			// the race for "nearest" is offered to "second nearest".
			unsigned long long* nearKey = key + s.nsP;
			const unsigned long long old = atomicMin(nearKey, nk);
			atomicMin(key, old > nk ? old : nk);
		}
		else
			atomicMin(key, nk);
		*(volatile uint32_t*)s.dirty = 1u;
	}
}

// Transparent hits of a tile (per persistent CTA, in HBM).  Every sample owns DEEP_INLINE in-line slots, stored RANK-MAJOR
// (slot j of sample idx at [j*nsP + idx]): the lanes of a warp work on neighbouring samples, so both the stores of
// the sampling loop and the loads of the resolve are runs of consecutive records instead of one 32-byte sector
// per lane.  Hits beyond the in-line slots go to an overflow pool chained per sample.  The per-sample word in shared
// memory (HideSmem::head) holds the hit count in bits 0-11 and the overflow chain head + 1 in bits 12-31.
// A hit is two records: A = (depth bits, position index of the micropolygon) -- what orders the list -- and
// B = (u, v, first Ci/Oi vertex, cu | shading flags) -- what shades it without another look at the grid tables.
#define DEEP_INLINE AQH_DEEP_INLINE
#define DEEP_COUNT_MASK 0xfffu
#define DEEP_COUNT_MAX 4000u
struct DeepCtx
{
	uint2* A; uint4* B;    // DEEP_INLINE * nsP in-line records, then ovCap overflow records
	uint32_t* ovNext;      // overflow chain: next slot + 1, 0 = end
	uint32_t ovCap;
	uint32_t* ovCount;     // lives in shared memory
	uint32_t ovBase;       // DEEP_INLINE * nsP
};
#define DEEP_NIL 0xffffffffu
// Handles of the entries of one sample: 0 .. DEEP_INLINE-1 = in-line slot, DEEP_INLINE + s = overflow slot s.
__device__ __forceinline__ uint32_t deepFirst(uint32_t word) { return (word & DEEP_COUNT_MASK) ? 0u : DEEP_NIL; }
__device__ __forceinline__ uint32_t deepNext(const DeepCtx& dc, uint32_t word, uint32_t h)
{
	if(h < DEEP_INLINE)
	{
		const uint32_t cnt = word & DEEP_COUNT_MASK;
		if(h + 1u < DEEP_INLINE && h + 1u < cnt) return h + 1u;
		const uint32_t ov = word >> 12;
		return ov ? DEEP_INLINE + ov - 1u : DEEP_NIL;
	}
	const uint32_t n = dc.ovNext[h - DEEP_INLINE];
	return n ? DEEP_INLINE + n - 1u : DEEP_NIL;
}
// index of an entry's records
__device__ __forceinline__ size_t deepAt(const DeepCtx& dc, int nsP, int idx, uint32_t h)
{
	return (h < DEEP_INLINE) ? (size_t)h*nsP + idx : (size_t)dc.ovBase + (h - DEEP_INLINE);
}
__device__ __forceinline__ void storeDeep(const DevFrame& f, const DeepCtx& dc, const HideSmem& s, int idx,
                                          float D, uint32_t p, float2 uv, uint32_t v0, uint32_t shade, bool cullable = true)
{
	const uint32_t occl = (uint32_t)(s.keys[idx] >> 32);
	if(cullable && !(depthKey(D) < occl)) return;     // isCullable && occlZ <= D
	uint32_t* word = &s.head[idx];
	const uint32_t rank = atomicAdd(word, 1u) & DEEP_COUNT_MASK;
	size_t at;
	if(rank < DEEP_INLINE) at = (size_t)rank*s.nsP + idx;
	else
	{
		// (a count about to leave its 12 bits, or a full pool: the frame reports AQH_ERR_DEEP_OVERFLOW)
		if(rank >= DEEP_COUNT_MAX) { atomicSub(word, 1u); atomicOr(f.errorFlags, 1u); return; }
		const uint32_t slot = atomicAdd(dc.ovCount, 1u);
		if(slot >= dc.ovCap) { atomicOr(f.errorFlags, 1u); return; }
		uint32_t cur = *(volatile uint32_t*)word;
		for(;;)
		{
			const uint32_t prev = atomicCAS(word, cur, (cur & DEEP_COUNT_MASK) | ((slot + 1u) << 12));
			if(prev == cur) break;
			cur = prev;
		}
		dc.ovNext[slot] = cur >> 12;
		at = (size_t)dc.ovBase + slot;
	}
	dc.A[at] = make_uint2(__float_as_uint(D), p);
	dc.B[at] = make_uint4(__float_as_uint(uv.x), __float_as_uint(uv.y), v0, shade);
}

// sample level of detail = lods[i] of the pixel's lod pattern (imagepixel.cpp:356)
__device__ __noinline__ float sampleLodImpl(const uint8_t* lodPlane, const float* val1d, int idx, int stride, int xs, int ys, int n,
                                            int px0, int py0, int sw)
{
	const int gy = idx / stride, gx = idx - gy*stride;
	const int plx = gx / xs, ply = gy / ys;
	const int i = (gy - ply*ys)*xs + (gx - plx*xs);
	const int pat = lodPlane[(size_t)(py0 + ply)*sw + (px0 + plx)];
	return val1d[(size_t)pat*n + i];
}
// Rare (level-of-detail ranges only): out of line.
__device__ __forceinline__ float sampleLod(const DevFrame& f, const TileCtx& t, const HideSmem& s, int idx)
{
	return sampleLodImpl(f.patPlanes + (size_t)4*f.sw*f.sh, f.val1d, idx, s.stride, f.xs, f.ys, f.n,
	                     t.tileX0 - f.sx0, t.tileY0 - f.sy0, f.sw);
}

// ---- hierarchical z.  pixZ[pixel] >= the occlusion depth key of every sample of the pixel (the keys
// only ever decrease, so a stale value is merely conservative).  A micropolygon whose nearest depth
// lies behind pixZ of every pixel it can touch would fail the reference's per-sample cull
// "Bound.zmin > occlZ" (bucketprocessor.cpp:1179, :1393) at each of its candidates, so it is
// dropped before its set-up: same image, far fewer instructions for hidden geometry.
__device__ __forceinline__ void refreshPixZ(const DevFrame& f, const TileCtx& t, const HideSmem& s, int lane)
{
	const int tw = t.rx1 - t.rx0, th = t.ry1 - t.ry0;
	const uint32_t* hi = reinterpret_cast<const uint32_t*>(s.keys) + 1;
	uint32_t tileMax = 0;
	for(int ly = 0; ly < th; ++ly)
		for(int lx = 0; lx < tw; ++lx)
		{
			const int base = (ly*f.ys)*s.stride + lx*f.xs;
			uint32_t m = 0;
			for(int i = lane; i < f.n; i += 32) m = max(m, hi[2*(base + s.subOfs[i])]);
			m = __reduce_max_sync(0xffffffffu, m);
			if(lane == 0) s.pixZ[ly*f.tileW + lx] = m;
			tileMax = max(tileMax, m);
		}
	if(lane == 0) *(volatile uint32_t*)s.tileZ = tileMax;
}
// largest pixZ over the pixel rectangle [sX,eX) x [sY,eY) (global pixel coordinates inside the tile)
__device__ __forceinline__ uint32_t pixZMax(const DevFrame& f, const TileCtx& t, const uint32_t* pixZ, int sX, int eX, int sY, int eY)
{
	uint32_t m = 0;
	for(int y = sY; y < eY; ++y)
		for(int x = sX; x < eX; ++x)
			m = max(m, pixZ[(y - t.tileY0)*f.tileW + (x - t.tileX0)]);
	return m;
}

// ---- static micropolygons, no depth of field: RenderMPG_Static (bucketprocessor.cpp:1097-1218)
// (inlined: a call would force the caller's TileCtx into local memory)
__device__ __forceinline__ void setupStaticRec(const DevFrame& f, const TileCtx& t, const uint32_t* pixZ, uint32_t p, bool wantOpaque, StaticRec& r)
{
	r.rect = 0;
	r.p = p;
	const float4 a = f.P4[p];
	const uint32_t info = infoOf(a);
	const uint32_t gi = info & VINFO_GRID_MASK;
	const GridRec g = f.grids[gi];
	const uint32_t cu = g.cu_cv & 0xffffu;
	const bool isPoint = (g.flags & AQH_GRID_POINTS) != 0;
	if(mpOpaqueSlot(f, g, p, info) != wantOpaque) return;
	const bool cullableMP = mpCullable(f, g);
	float4 P[4];
	P[0] = a;
	B2 B;
	float pointR = 0.f;
	if(isPoint)
	{
		pointR = f.radius[p];
		P[1] = a; P[2] = a; P[3] = a;
		B.mnx = a.x - pointR; B.mny = a.y - pointR; B.mnz = a.z - 0.f; B.mxx = a.x + pointR; B.mxy = a.y + pointR; B.mxz = a.z + 0.f;
	}
	else
	{
		P[1] = f.P4[p+1]; P[2] = f.P4[p+cu+1]; P[3] = f.P4[p+cu+2];
		B = boundOf4(P[0], P[1], P[3], P[2]);
	}
	const float bminx = B.mnx, bmaxx = B.mxx, bminy = B.mny, bmaxy = B.mxy;
	// pixel range clamped to (tile ∩ sample region); compare in float first so that huge
	// bounds cannot overflow the int conversions.
	if(!(bmaxx >= (float)t.rx0) || !(bmaxy >= (float)t.ry0) || !(bminx < (float)t.rx1) || !(bminy < (float)t.ry1)) return;
	int eX = (bmaxx >= (float)t.rx1) ? t.rx1 : min(lceilF(bmaxx), t.rx1);
	int eY = (bmaxy >= (float)t.ry1) ? t.ry1 : min(lceilF(bmaxy), t.ry1);
	int sX = (bminx < (float)t.rx0) ? t.rx0 : max(floorI(bminx), t.rx0);
	int sY = (bminy < (float)t.ry0) ? t.ry0 : max(floorI(bminy), t.ry0);
	if(sX >= eX || sY >= eY) return;
	if(cullableMP && depthKey(B.mnz) > pixZMax(f, t, pixZ, sX, eX, sY, eY)) return;      // hidden behind every sample it could touch
	const int xs = f.xs, ys = f.ys;
	int im = (bminx < (float)sX) ? 0 : floorI((bminx - (float)sX) * (float)xs);
	int in = (bminy < (float)sY) ? 0 : floorI((bminy - (float)sY) * (float)ys);
	int em = (bmaxx > (float)eX) ? xs : lceilF((bmaxx - (float)(eX - 1)) * (float)xs);
	int en = (bmaxy > (float)eY) ? ys : lceilF((bmaxy - (float)(eY - 1)) * (float)ys);
	im = max(im, 0); in = max(in, 0); em = min(em, xs); en = min(en, ys);
	int gx0 = (sX - t.tileX0)*xs + im, gx1 = (eX - 1 - t.tileX0)*xs + em;
	int gy0 = (sY - t.tileY0)*ys + in, gy1 = (eY - 1 - t.tileY0)*ys + en;
	if(gx1 <= gx0 || gy1 <= gy0) return;
	if(isPoint)
	{
		r.bminx = bminx; r.bminy = bminy; r.bmaxx = bmaxx; r.bmaxy = bmaxy;
		r.Ax = a.x; r.Ay = a.y; r.Ex = pointR;
		r.z[0] = a.z;
		r.zminKey = cullableMP ? depthKey(B.mnz) : 0u;
		r.flags = gi | REC_POINT | REC_RARE;
		{ const uint2 si = hitShadeInfo(g, p); r.v0 = si.x; r.shade = si.y; }
		r.rect = (uint32_t)gx0 | ((uint32_t)gx1 << 8) | ((uint32_t)gy0 << 16) | ((uint32_t)gy1 << 24);
		return;
	}
	const int code = computeVertexOrder(P);
	HitCache c;
	const float px[4] = {P[0].x, P[1].x, P[2].x, P[3].x};
	const float py[4] = {P[0].y, P[1].y, P[2].y, P[3].y};
	const float pz[4] = {P[0].z, P[1].z, P[2].z, P[3].z};
	cachePointInPolyTest(c, px, py, pz, code);
#pragma unroll
	for(int i = 0; i < 4; ++i) { r.X[i] = c.X[i]; r.Y[i] = c.Y[i]; r.XM[i] = c.XM[i]; r.YM[i] = c.YM[i]; r.z[i] = c.z[i]; }
	r.bminx = bminx; r.bminy = bminy; r.bmaxx = bmaxx; r.bmaxy = bmaxy;
	r.Ax = c.Ax; r.Ay = c.Ay; r.Ex = c.Ex; r.Ey = c.Ey; r.Fx = c.Fx; r.Fy = c.Fy; r.Gx = c.Gx; r.Gy = c.Gy;
	r.zminKey = cullableMP ? depthKey(B.mnz) : 0u;     // a non-cullable (CSG) micropolygon is never rejected by depth
	uint32_t fl = gi;
	if(c.linear) fl |= REC_LINEAR;
	if((g.flags & AQH_GRID_TRIANGULAR) || g.lod0 >= 0.0f) fl |= REC_RARE;
	if(info & VINFO_MP_TRIMMED) fl |= REC_RARE | REC_TRIM;
	r.flags = fl;
	{ const uint2 si = hitShadeInfo(g, p); r.v0 = si.x; r.shade = si.y; }
	r.rect = (uint32_t)gx0 | ((uint32_t)gx1 << 8) | ((uint32_t)gy0 << 16) | ((uint32_t)gy1 << 24);
}

// The motion blur / depth of field kernel meets static micropolygons rarely (frames without depth of field only): one
// out-of-line copy, its tile passed by value.
__device__ __noinline__ void setupStaticRecCall(const DevFrame& f, TileCtx t, const uint32_t* pixZ, uint32_t p, bool wantOpaque, StaticRec* r)
{
	setupStaticRec(f, t, pixZ, p, wantOpaque, *r);
}

// The sampling loop of one record.  RARE = false is the hot instantiation: plain quadrilaterals, nothing but the bound
// test, the occlusion cull, the edge tests, the inverse bilinear map and the store.  Everything uncommon -- discs of
// RiPoints, level-of-detail windows, trim curves, the split line of triangular grids -- lives in the RARE = true copy,
// which exists once per kernel, out of line (sampleStaticRecRare): that code inside the hot loop cost 5-10 % of the
// whole frame in registers and instruction fetch even when it never ran (profiles/README.md, A/B r2w-r2y).
template<bool OPAQUE, bool RARE, bool PL = false>
__device__ __forceinline__ void sampleStaticLoop(const DevFrame& f, const TileCtx& t, const HideSmem& s, const DeepCtx& dc,
                                                 const StaticRec& r, int lane)
{
	const uint32_t rect = r.rect;
	if(rect == 0) return;
	const int gx0 = rect & 0xff, gx1 = (rect >> 8) & 0xff, gy0 = (rect >> 16) & 0xff, gy1 = rect >> 24;
	const int W = gx1 - gx0, total = W*(gy1 - gy0);
	// lanes walk the window in row-major order, 32 candidates per step
	const int dy = 32 / W, dx = 32 - dy*W;
	int cy = lane / W, cx = lane - cy*W;
	const float bminx = r.bminx, bminy = r.bminy, bmaxx = r.bmaxx, bmaxy = r.bmaxy;
	const uint32_t zminKey = r.zminKey;
	const int stride = s.stride;
	for(int k = lane; k < total; k += 32)
	{
		const int idx = (gy0 + cy)*stride + gx0 + cx;
		cx += dx; cy += dy;
		if(cx >= W) { cx -= W; cy += 1; }
		const float x = s.posx[idx], y = s.posy[idx];
		// Bound.Contains2D, bound.h:144-151
		if((x < bminx || x > bmaxx) || (y < bminy || y > bmaxy)) continue;
		// occlusion cull against the current opaque depth (bucketprocessor.cpp:1179)
		const uint32_t occl = (uint32_t)(s.keys[idx] >> 32);
		if(zminKey > occl) continue;
		float2 uv; float D;
		if(RARE && (r.flags & REC_POINT))
		{
			// CqMicroPolygonPoints::Sample, geometry/points.cpp:653-664
			const float dx = r.Ax - x, dy = r.Ay - y;
			if(!((dx*dx + dy*dy) < r.Ex*r.Ex)) continue;
			uv = make_float2(0.f, 0.f);
			D = r.z[0];
		}
		else
		{
			if(!edgeTests(r.X, r.Y, r.XM, r.YM, x, y)) continue;
			uv = invBilinear(r.Ax, r.Ay, r.Ex, r.Ey, r.Fx, r.Fy, r.Gx, r.Gy, (r.flags & REC_LINEAR) ? 1 : 0, x, y);
			D = bilerpZ(r.z, uv);
		}
		if(RARE)
		{
			const GridRec g = f.grids[r.flags & VINFO_GRID_MASK];
			if(g.lod0 >= 0.0f)
			{
				const float lod = sampleLod(f, t, s, idx);
				if(g.lod0 > lod || lod >= g.lod1) continue;
			}
			if(r.flags & REC_TRIM)
				if(trimRejectHit(f.gridTrim, f.trimSetLoop, f.trimLoopPoint, f.trimPoints, f.trimUV, r.flags & VINFO_GRID_MASK, g.flags,
				                 r.v0, g.cu_cv & 0xffffu, uv.x, uv.y)) continue;
			if(g.flags & AQH_GRID_TRIANGULAR)
				if(triangleSplitReject(f, g, make_float2(x, y), make_float2(0.f, 0.f), D, 0.0f)) continue;
		}
		if(OPAQUE)
			storeOpaque<PL>(f, s, &s.keys[idx], D, r.p);
		else
			storeDeep(f, dc, s, idx, D, r.p, uv, r.v0, r.shade, zminKey != 0u);
	}
}
// (the structs travel by value: a reference would pin the caller's copies in local memory)
__device__ __noinline__ void sampleStaticRecRare(const DevFrame& f, TileCtx t, HideSmem s, DeepCtx dc, const StaticRec* r, int lane, bool opaque)
{
	if(opaque) sampleStaticLoop<true, true>(f, t, s, dc, *r, lane);
	else sampleStaticLoop<false, true>(f, t, s, dc, *r, lane);
}
// PL: a PLAIN frame (DevFrame::plain) has no uncommon records
template<bool OPAQUE, bool AGG, bool PL = false>
__device__ __forceinline__ void sampleStaticRec(const DevFrame& f, const TileCtx& t, const HideSmem& s, const DeepCtx& dc,
                                                const StaticRec& r, int lane)
{
	if(r.rect == 0) return;                                                       // rejected at set-up (its flags are stale)
	if(!PL && (r.flags & REC_RARE)) sampleStaticRecRare(f, t, s, dc, &r, lane, OPAQUE);      // warp-uniform
	else sampleStaticLoop<OPAQUE, false, PL>(f, t, s, dc, r, lane);
}

// ---- motion blur and/or depth of field: RenderMPG_MBOrDof (bucketprocessor.cpp:1221-1469).
// One warp per micropolygon.  The key vertices and key bounds of the micropolygon are staged once in the
// warp's shared-memory scratch.  The reference's loops -- time sub-bounds, then lens cells (DoF) or
// (pixel, sample index) pairs (MB) -- are walked by the lanes with only the reference's CHEAP gates
// (time window, shifted-bound test, occlusion cull) evaluated in place; survivors are pushed on the warp's
// candidate queue, and the expensive part (vertex interpolation at the sample time, per-vertex circle of
// confusion, edge set-up, edge tests, inverse bilinear) is run on full batches of 32 queued candidates,
// one per lane, so that the heavy code exists once and runs with (almost) all lanes active.
struct MovingMP
{
	uint32_t p, nverts, cu, nkeys;
	const float* times;
	int code;
	GridRec g;
};

// PLAIN: the instantiations of k_hide for frames WITHOUT discs, level-of-detail ranges, trim curves, triangular grids, more than two
// motion keys, CSG solids, arbitrary output variables, an incremental flush / occlusion-only pass or the midpoint depth filter
// (DevFrame::plain, decided on the host from the frame options and the grids submitted; AQH_TUNE=0,0,0,0,1 forces the general
// kernels, tests/test_features_gpu.py compares the two).  The general motion kernel is bound by instruction fetch; without those
// branches the hot code is only a few per cent shorter, but it stays in the instruction cache: config 3 426 ms (general,
// free-running warps) -> 238 ms.  The static kernel loses its out-of-line calls (rare records, CSG resolve) and with them a
// 960-byte stack frame: 10 968 -> 5 760 SASS instructions, config 2 hide 10.3 -> 9.3 ms, config 4 -7.5 % (profiles/README.md).
#define MB_RARE(x) (!PLAIN && (x))
#define PLAIN_K2 PLAIN        /* at most two motion keys */
#define MOV_KMAX 4          /* keys staged in shared memory; further keys are read from HBM */
#define MOV_QCAP 1056       /* 1024 pushes per enumeration round + the < 32 left by the previous one */
struct MovScratch           // per warp, 16-byte aligned
{
	float4 kv[MOV_KMAX][4];         // key vertices in natural order (x, y, z, info)
	float4 kb[MOV_KMAX][2];         // key bounds: (mnx, mny, mnz, mxx), (mxy, mxz, -, -)
	uint32_t qcount, pad[3];
	uint16_t q[MOV_QCAP];           // candidate sample indices
};

// keys past MOV_KMAX (and the resolve stage) read HBM: out of line, the staged case is the hot one
__device__ __noinline__ float4 movVertGlobal(const float4* Pk, uint32_t cu, int i)
{
	return Pk[(uint32_t)(i & 1) + (uint32_t)(i >> 1)*(cu + 1)];
}
__device__ __noinline__ B2 keyBoundGlobal(const float4* Pk, uint32_t cu)
{
	return boundOf4(Pk[0], Pk[1], Pk[cu+1], Pk[cu+2]);
}
template<bool PLAIN>
__device__ __forceinline__ float4 movVert(const DevFrame& f, const MovingMP& m, const MovScratch* ws, uint32_t k, int i)
{
	if(PLAIN) return ws->kv[k][i];
	if(ws && k < MOV_KMAX) return ws->kv[k][i];
	return movVertGlobal(f.P4 + m.p + (size_t)k*m.nverts, m.cu, i);
}
template<bool PLAIN>
__device__ __forceinline__ B2 keyBound(const DevFrame& f, const MovingMP& m, const MovScratch* ws, uint32_t k)
{
	if(PLAIN || (ws && k < MOV_KMAX))
	{
		const float4 a = ws->kb[k][0], b = ws->kb[k][1];
		B2 r; r.mnx = a.x; r.mny = a.y; r.mnz = a.z; r.mxx = a.w; r.mxy = b.x; r.mxz = b.y;
		return r;
	}
	return keyBoundGlobal(f.P4 + m.p + (size_t)k*m.nverts, m.cu);
}

// Vertices of the micropolygon for one sample: CqMicroPolygon(Motion)::Sample,
// micropolygon.cpp:1561-1589 and 1768-1873.  Returns false when the tight-bound test fails.
template<bool PLAIN>
__device__ __forceinline__ bool samplePoints(const DevFrame& f, const MovingMP& m, const MovScratch* ws, bool moving, const B2& mpBound,
                                             float2 cocMin, float2 cocMax, float2 pos, float2 dofOff, float time,
                                             float px[4], float py[4], float pz[4], bool doBoundTest)
{
	B2 tight;
	float Fraction = 0.0f;
	bool Exact = true;
	uint32_t iIndex = 0;
	if(moving)
	{
		const float* times = m.times;
		if(time > times[0])
		{
			if(time >= times[PLAIN_K2 ? 1u : m.nkeys-1])
				iIndex = PLAIN_K2 ? 1u : m.nkeys - 1;
			else
			{
				if(!PLAIN_K2) while(time >= times[iIndex+1]) iIndex += 1;
				Fraction = (time - times[iIndex]) / (times[iIndex+1] - times[iIndex]);
				Exact = (times[iIndex] == time);
			}
		}
		if(doBoundTest)
		{
			if(Exact)
				tight = keyBound<PLAIN>(f, m, ws, iIndex);
			else
			{
				const B2 b1 = keyBound<PLAIN>(f, m, ws, iIndex), b2 = keyBound<PLAIN>(f, m, ws, iIndex+1);
				tight.mnx = (1.f-Fraction)*b1.mnx + Fraction*b2.mnx; tight.mny = (1.f-Fraction)*b1.mny + Fraction*b2.mny;
				tight.mnz = (1.f-Fraction)*b1.mnz + Fraction*b2.mnz;
				tight.mxx = (1.f-Fraction)*b1.mxx + Fraction*b2.mxx; tight.mxy = (1.f-Fraction)*b1.mxy + Fraction*b2.mxy;
				tight.mxz = (1.f-Fraction)*b1.mxz + Fraction*b2.mxz;
			}
		}
	}
	else
		tight = mpBound;
	if(doBoundTest)
	{
		if(f.useDof)
		{
			// dofSampleInBound, micropolygon.cpp:1531-1550
			float mnx = pos.x + cocMin.x*dofOff.x, mny = pos.y + cocMin.y*dofOff.y;
			float mxx = pos.x + cocMax.x*dofOff.x, mxy = pos.y + cocMax.y*dofOff.y;
			if(dofOff.x < 0.f) { float tt = mnx; mnx = mxx; mxx = tt; }
			if(dofOff.y < 0.f) { float tt = mny; mny = mxy; mxy = tt; }
			if(mnx > tight.mxx || mny > tight.mxy || mxx < tight.mnx || mxy < tight.mny) return false;
		}
		else
		{
			if((pos.x < tight.mnx || pos.x > tight.mxx) || (pos.y < tight.mny || pos.y > tight.mxy)) return false;
		}
	}
	if(Exact)
	{
#pragma unroll
		for(int i = 0; i < 4; ++i) { const float4 v = movVert<PLAIN>(f, m, ws, iIndex, i); px[i] = v.x; py[i] = v.y; pz[i] = v.z; }
	}
	else
	{
		const float F1 = 1.0f - Fraction;
#pragma unroll
		for(int i = 0; i < 4; ++i)
		{
			const float4 a = movVert<PLAIN>(f, m, ws, iIndex, i), b = movVert<PLAIN>(f, m, ws, iIndex+1, i);
			px[i] = (F1*a.x) + (Fraction*b.x);
			py[i] = (F1*a.y) + (Fraction*b.y);
			pz[i] = (F1*a.z) + (Fraction*b.z);
		}
	}
	if(f.useDof)
	{
#pragma unroll
		for(int i = 0; i < 4; ++i)
		{
			float2 cm = cocAt(f, pz[i]);
			px[i] = px[i] - cm.x*dofOff.x;
			py[i] = py[i] - cm.y*dofOff.y;
		}
	}
	return true;
}

// Per-micropolygon constants of the heavy part.
struct MovCtx
{
	MovingMP m;
	B2 mpBound;
	float2 cocMin, cocMax;
	bool moving, opaquePass, cullable;
	float pointR;            // > 0: a disc (CqMicroPolygonPoints), centre = the staged vertex
	bool isPoint;
	uint2 shadeInfo;         // hitShadeInfo() of the micropolygon, for its transparent hit records
	bool trimmed;            // a trim curve crosses the micropolygon
	uint32_t gridIndex;
};

// One queued candidate (micropolygon, sample): everything of CqMicroPolygon(Motion)::Sample after the gates
// that were evaluated at enumeration time.
template<bool PLAIN>
__device__ __forceinline__ void movCandidate(const DevFrame& f, const TileCtx& t, const HideSmem& s, const DeepCtx& dc,
                                             const MovCtx& c, const MovScratch* ws, int idx)
{
	const float2 pos = make_float2(s.posx[idx], s.posy[idx]);
	const float time = s.time ? s.time[idx] : f.shutterOpen;
	if(MB_RARE(c.m.g.lod0 >= 0.0f))
	{
		float lod = sampleLod(f, t, s, idx);
		if(c.m.g.lod0 > lod || lod >= c.m.g.lod1) return;
	}
	const float2 dofOff = s.dof ? s.dof[idx] : make_float2(0.f, 0.f);
	if(MB_RARE(c.isPoint))
	{
		// CqMicroPolygonPoints::Sample under depth of field (geometry/points.cpp:653-664): the SAMPLE is moved by its lens
		// offset times the point's circle of confusion
		const float4 P0 = ws->kv[0][0];
		const float2 cm = cocAt(f, P0.z);
		const float sx = pos.x + dofOff.x*cm.x, sy = pos.y + dofOff.y*cm.y;
		const float dx = P0.x - sx, dy = P0.y - sy;
		if(!((dx*dx + dy*dy) < c.pointR*c.pointR)) return;
		if(c.opaquePass) storeOpaque<PLAIN>(f, s, &s.keys[idx], P0.z, c.m.p);
		else storeDeep(f, dc, s, idx, P0.z, c.m.p, make_float2(0.f, 0.f), c.shadeInfo.x, c.shadeInfo.y, c.cullable);
		return;
	}
	float px[4], py[4], pz[4];
	if(!samplePoints<PLAIN>(f, c.m, ws, c.moving, c.mpBound, c.cocMin, c.cocMax, pos, dofOff, time, px, py, pz, true)) return;
	HitCache hc;
	cachePointInPolyTest(hc, px, py, pz, c.m.code);
	if(!edgeTests(hc.X, hc.Y, hc.XM, hc.YM, pos.x, pos.y)) return;
	const float2 uv = invBilinear(hc.Ax, hc.Ay, hc.Ex, hc.Ey, hc.Fx, hc.Fy, hc.Gx, hc.Gy, hc.linear, pos.x, pos.y);
	const float D = bilerpZ(hc.z, uv);
	// (CqMicroPolygonMotion::Sample leaves the per-hit trim test as a todo, micropolygon.cpp:1877-1884: only micropolygons
	// that do not move are tested)
	if(MB_RARE(c.trimmed && !c.moving))
		if(trimRejectHit(f.gridTrim, f.trimSetLoop, f.trimLoopPoint, f.trimPoints, f.trimUV, c.gridIndex, c.m.g.flags,
		                 c.shadeInfo.x, c.m.cu, uv.x, uv.y)) return;
	if(MB_RARE(c.m.g.flags & AQH_GRID_TRIANGULAR))
		if(triangleSplitReject(f, c.m.g, pos, dofOff, D, time)) return;
	if(c.opaquePass)
		storeOpaque<PLAIN>(f, s, &s.keys[idx], D, c.m.p);
	else
		storeDeep(f, dc, s, idx, D, c.m.p, uv, c.shadeInfo.x, c.shadeInfo.y, c.cullable);
}

// Queue a candidate that passed the cheap gates (called from divergent per-lane loops).
__device__ __forceinline__ void movPush(const DevFrame& f, MovScratch* ws, int idx)
{
	const uint32_t slot = atomicAdd(&ws->qcount, 1u);
	if(slot < MOV_QCAP) ws->q[slot] = (uint16_t)idx;
	else atomicOr(f.errorFlags, 4u);          // cannot happen: rounds are sized to the queue
}

// Run the heavy part on batches of up to 32 queued candidates while at least `need` are waiting.
// Called by the whole warp at a converged point.
template<bool PLAIN>
__device__ __forceinline__ void movDrain(const DevFrame& f, const TileCtx& t, const HideSmem& s, const DeepCtx& dc,
                                         const MovCtx& c, MovScratch* ws, int lane, uint32_t need)
{
	for(;;)
	{
		__syncwarp();
		uint32_t cnt = *(volatile uint32_t*)&ws->qcount;
		if(cnt > MOV_QCAP) cnt = MOV_QCAP;
		if(cnt < need) break;
		const uint32_t take = cnt < 32u ? cnt : 32u, base = cnt - take;
		__syncwarp();
		if(lane == 0) *(volatile uint32_t*)&ws->qcount = base;
		if((uint32_t)lane < take) movCandidate<PLAIN>(f, t, s, dc, c, ws, (int)ws->q[base + lane]);
	}
	__syncwarp();
}

// Returns false for a static micropolygon in a frame without depth of field: the reference
// renders those with RenderMPG_Static even when other grids move (bucketprocessor.cpp:1087-1090).
template<bool PLAIN>
__device__ __forceinline__ bool renderMBOrDof(const DevFrame& f, const TileCtx& t, const HideSmem& s, const DeepCtx& dc, MovScratch* ws,
                              uint32_t p, int lane, bool opaquePass)
{
	MovCtx c;
	MovingMP& m = c.m;
	m.p = p;
	const float4 a = f.P4[p];
	const uint32_t info = infoOf(a);
	m.g = f.grids[info & VINFO_GRID_MASK];
	m.cu = m.g.cu_cv & 0xffffu;
	m.nverts = m.g.nverts;
	m.nkeys = m.g.nkeys_koff & 0xffu;
	m.times = f.keyTimes + (m.g.nkeys_koff >> 8);
	const bool moving = m.nkeys > 1;
	if(!moving && !f.useDof) return false;
	c.moving = moving; c.opaquePass = opaquePass;
	c.isPoint = MB_RARE((m.g.flags & AQH_GRID_POINTS) != 0);
	c.pointR = c.isPoint ? f.radius[p] : 0.f;
	c.cullable = mpCullable(f, m.g);
	c.shadeInfo = hitShadeInfo(m.g, p);
	c.trimmed = (info & VINFO_MP_TRIMMED) != 0;
	c.gridIndex = info & VINFO_GRID_MASK;
	// ---- stage the key vertices and key bounds in the warp's scratch
	__syncwarp();
	{
		const uint32_t nk = PLAIN_K2 ? (moving ? 2u : 1u) : (m.nkeys < MOV_KMAX ? m.nkeys : MOV_KMAX);
		if((uint32_t)lane < 4u*nk)
		{
			const uint32_t k = (uint32_t)lane >> 2; const int i = lane & 3;
			// a point has one vertex: all four slots hold it
			ws->kv[k][i] = c.isPoint ? a : f.P4[p + (size_t)k*m.nverts + (uint32_t)(i & 1) + (uint32_t)(i >> 1)*(m.cu + 1)];
		}
		if(lane == 0) ws->qcount = 0;
		__syncwarp();
		if((uint32_t)lane < nk)
		{
			B2 b = boundOf4(ws->kv[lane][0], ws->kv[lane][1], ws->kv[lane][2], ws->kv[lane][3]);
			if(c.isPoint) { b.mnx = a.x - c.pointR; b.mny = a.y - c.pointR; b.mxx = a.x + c.pointR; b.mxy = a.y + c.pointR; b.mnz = a.z - 0.f; b.mxz = a.z + 0.f; }
			ws->kb[lane][0] = make_float4(b.mnx, b.mny, b.mnz, b.mxx);
			ws->kb[lane][1] = make_float4(b.mxy, b.mxz, 0.f, 0.f);
		}
		__syncwarp();
	}
	float4 P[4];
	P[0] = ws->kv[0][0]; P[1] = ws->kv[0][1]; P[2] = ws->kv[0][2]; P[3] = ws->kv[0][3];
	bool opaque = (info & VINFO_OPAQUE) != 0;
	if((m.g.flags & AQH_GRID_SMOOTH) && !c.isPoint)
		opaque = opaque && (infoOf(P[1]) & VINFO_OPAQUE) && (infoOf(P[2]) & VINFO_OPAQUE) && (infoOf(P[3]) & VINFO_OPAQUE);
	const bool opaqueSlot = (opaque || (m.g.flags & AQH_GRID_MATTE_ALPHA)) && c.cullable;
	if(opaqueSlot != opaquePass) return true;
	m.code = c.isPoint ? 0xE4 : computeVertexOrder(P);
	// m_Bound: union of the key bounds (AppendKey, micropolygon.cpp:1952-1967)
	B2 mpBound = keyBound<PLAIN>(f, m, ws, 0);
	if(PLAIN_K2) { if(moving) { B2 kb = keyBound<PLAIN>(f, m, ws, 1); encapsulate(mpBound, kb); } }
	else for(uint32_t k = 1; k < m.nkeys; ++k) { B2 kb = keyBound<PLAIN>(f, m, ws, k); encapsulate(mpBound, kb); }
	c.mpBound = mpBound;
	// CacheHitTestValues (static :1406-1423, moving :1916-1938)
	float2 cocMin = make_float2(0.f, 0.f), cocMax = make_float2(0.f, 0.f);
	if(f.useDof)
	{
		if(!moving)
		{
			float2 c0 = cocAt(f, P[0].z), c1 = cocAt(f, P[1].z), c2 = cocAt(f, P[2].z), c3 = cocAt(f, P[3].z);
			cocMin = make_float2(minA(minA(c0.x, c1.x), minA(c2.x, c3.x)), minA(minA(c0.y, c1.y), minA(c2.y, c3.y)));
			cocMax = make_float2(maxA(maxA(c0.x, c1.x), maxA(c2.x, c3.x)), maxA(maxA(c0.y, c1.y), maxA(c2.y, c3.y)));
		}
		else
		{
			float2 c1 = cocAt(f, mpBound.mnz), c2 = cocAt(f, mpBound.mxz);
			if(minCoCForBound(f, mpBound.mnz, mpBound.mxz) == 0.f) cocMin = make_float2(0.f, 0.f);
			else cocMin = make_float2(minA(c1.x, c2.x), minA(c1.y, c2.y));
			cocMax = make_float2(maxA(c1.x, c2.x), maxA(c1.y, c2.y));
		}
	}
	c.cocMin = cocMin; c.cocMax = cocMax;
	const int n = f.n;
	const float opentime = f.shutterOpen, closetime = f.shutterClose;
	float timePerSample = 0.f;
	bool fastShutter = false;
	if(moving)
	{
		const float tol = 10.f*FLT_EPSILON;
		float d = fabsf(closetime - opentime);
		fastShutter = (d <= tol*fabsf(closetime)) || (d <= tol*fabsf(opentime));   // isClose, math.h:174-182
		if(!fastShutter) timePerSample = (float)n / (closetime - opentime);
	}
	// BuildBoundList (micropolygon.cpp:1689-1756): the shutter is cut into `divisions` time ranges, each with the bound of
	// the micropolygon over that range.
	int divisions = 1;
	float dt = 0.f;
	B2 kb0 = mpBound;
	if(moving)
	{
		kb0 = keyBound<PLAIN>(f, m, ws, 0);
		float cx = kb0.mxx - kb0.mnx, cy = kb0.mxy - kb0.mny;
		float polyLen2 = (cy == 0.f) ? cx*cx : ((cx == 0.f) ? cy*cy : cx*cx + cy*cy);
		const float4 pl = movVert<PLAIN>(f, m, ws, PLAIN_K2 ? 1u : m.nkeys-1, 0);
		float mx = P[0].x - pl.x, my = P[0].y - pl.y;
		float moveDist2 = (my == 0.f) ? mx*mx : ((mx == 0.f) ? my*my : mx*mx + my*my);
		int polyLengthsMoved = max(1, lfloorF(sqrtf(moveDist2/polyLen2)));
		const int timeRanges = max(4, n);
		divisions = min(polyLengthsMoved, timeRanges);
		dt = (closetime - opentime) / (float)(unsigned)divisions;
	}
	// rows of tile pixels enumerated per round, so that one round queues at most 32 lanes x 32 pixels
	const int tileRows = t.ry1 - t.ry0;
	const int bandRows = (f.tileW >= 32) ? 1 : (32 / f.tileW);
	const int cellRounds = (n + 31) >> 5;
	// The divisions are set up 32 at a time, ONE PER LANE (the reference walks them one after the other; every quantity of
	// a division only depends on its index -- the running time is re-added from the group's start, the running bound is
	// the previous division's end bound): time window, bound, sample-index window, depth key, circle of confusion, and
	// whether anything of it can reach the tile.  The survivors are then enumerated one by one with all lanes:
	// rounds of at most 1024 candidates driven by ONE loop with ONE drain call, so that the heavy per-candidate code
	// exists once in the instruction stream.
	const int nGroups = (divisions + 31) >> 5;
	int group = -1, round = 0, nRounds = 0;
	uint32_t survivors = 0;
	float accBase = opentime + dt;           // the reference's timeAcc at the first division of the group
	B2 carryBound = kb0;                     // its runBound there: the end bound of the division before
	uint32_t carryKey = 0;                   // and its startKey: the last key the earlier divisions have passed
	// this lane's division of the current group
	B2 dBnd = mpBound;
	float dTime0 = 0.f, dTime1 = 0.f, dCocX = 0.f, dCocY = 0.f;
	int dIndexT0 = 0, dIndexT1 = 0;
	uint32_t dZminKey = 0;
	// the division being enumerated (warp-uniform)
	B2 Bnd = mpBound;
	float time0 = 0.f, time1 = 0.f;
	int indexT0 = 0, indexT1 = 0;
	uint32_t zminKey = 0;
	float maxCocX = 0.f, maxCocY = 0.f;
	int sX = 0, eX = 0, sY = 0, eY = 0, cnt = 1, total = 0;        // pixel window and index window of the flattened enumerations
	bool byIndex = false;                                          // depth of field: enumerate (pixel, sample index) instead of (lens cell, pixel)
	for(;;)
	{
		bool last = false;
		if(round >= nRounds)
		{
			for(;;)
			{
				if(survivors == 0u)
				{
					if(++group >= nGroups) { last = true; break; }
					const int d = (group << 5) + lane;
					bool ok = d < divisions;
					dBnd = mpBound;
					if(moving)
					{
						const float* times = m.times;
						// timeAcc of division d: opentime + dt, then + dt once per earlier division (the same additions in the same order)
						// (lanes past the last division carry no division: they need not add)
						float acc = accBase;
						const int nAdd = ok ? lane : 0;
						for(int k = 0; k < nAdd; ++k) acc = acc + dt;
						uint32_t endKey = 1;
						if(!PLAIN_K2) while(acc > times[endKey] && endKey < m.nkeys - 1) ++endKey;
						const uint32_t endKey_1 = endKey - 1;
						const B2 end0 = keyBound<PLAIN>(f, m, ws, endKey_1), end1 = keyBound<PLAIN>(f, m, ws, endKey);
						const float end0Time = times[endKey_1], end1Time = times[endKey];
						const float mix = (acc - end0Time) / (end1Time - end0Time);
						B2 mid = end0;
						mid.mnx += mix * (end1.mnx - end0.mnx); mid.mny += mix * (end1.mny - end0.mny); mid.mnz += mix * (end1.mnz - end0.mnz);
						mid.mxx += mix * (end1.mxx - end0.mxx); mid.mxy += mix * (end1.mxy - end0.mxy); mid.mxz += mix * (end1.mxz - end0.mxz);
						// the running bound the reference enters this division with: the end bound of the previous one
						B2 run;
						run.mnx = __shfl_up_sync(0xffffffffu, mid.mnx, 1); run.mny = __shfl_up_sync(0xffffffffu, mid.mny, 1); run.mnz = __shfl_up_sync(0xffffffffu, mid.mnz, 1);
						run.mxx = __shfl_up_sync(0xffffffffu, mid.mxx, 1); run.mxy = __shfl_up_sync(0xffffffffu, mid.mxy, 1); run.mxz = __shfl_up_sync(0xffffffffu, mid.mxz, 1);
						// ... and the last key it had passed by then (startKey): the end key - 1 of the previous division, 0 at the start
						uint32_t startKey = PLAIN_K2 ? 0u : __shfl_up_sync(0xffffffffu, endKey_1, 1);
						if(lane == 0) { run = carryBound; startKey = carryKey; }
						encapsulate(run, mid);
						if(!PLAIN) while(startKey < endKey_1) { startKey++; const B2 kb = keyBound<PLAIN>(f, m, ws, startKey); encapsulate(run, kb); }
						dBnd = run;
						dTime0 = acc - dt;
						const float nextAcc = acc + dt;
						dTime1 = (d != divisions - 1) ? (nextAcc - dt) : closetime;
						// hand the group's end state to the next group
						accBase = __shfl_sync(0xffffffffu, nextAcc, 31);
						carryBound.mnx = __shfl_sync(0xffffffffu, mid.mnx, 31); carryBound.mny = __shfl_sync(0xffffffffu, mid.mny, 31); carryBound.mnz = __shfl_sync(0xffffffffu, mid.mnz, 31);
						carryBound.mxx = __shfl_sync(0xffffffffu, mid.mxx, 31); carryBound.mxy = __shfl_sync(0xffffffffu, mid.mxy, 31); carryBound.mxz = __shfl_sync(0xffffffffu, mid.mxz, 31);
						if(!PLAIN_K2) carryKey = __shfl_sync(0xffffffffu, endKey_1, 31);
						if(dTime1 < opentime || dTime0 > closetime) ok = false;
						if(fastShutter) { dIndexT0 = 0; dIndexT1 = n; }
						else
						{
							dIndexT0 = max(0, lfloorF((dTime0 - opentime) * timePerSample));
							dIndexT1 = lceilF((dTime1 - opentime) * timePerSample);
						}
						if(dIndexT1 > n) dIndexT1 = n;
						if(dIndexT0 >= n) ok = false;
					}
					if(dBnd.mnz > f.clipFar || dBnd.mxz < f.clipNear) ok = false;
					dZminKey = c.cullable ? depthKey(dBnd.mnz) : 0u;      // a CSG micropolygon is never rejected by depth
					if(f.useDof)
					{
						float2 c1 = cocAt(f, dBnd.mnz), c2 = cocAt(f, dBnd.mxz);
						dCocX = maxA(c1.x, c2.x); dCocY = maxA(c1.y, c2.y);
					}
					// quick reject of the whole division: the union of the lens-cell boxes (dofBounds lie in [-1,1], so every cell
					// box is inside the bound grown by the largest circle of confusion) or the bound itself misses the tile
					if(!(dBnd.mxx + dCocX >= (float)t.rx0) || !(dBnd.mxy + dCocY >= (float)t.ry0) ||
					   !(dBnd.mnx - dCocX < (float)t.rx1) || !(dBnd.mny - dCocY < (float)t.ry1)) ok = false;
					// nothing behind the whole tile can be visible in it
					if(ok && dZminKey > *(volatile uint32_t*)s.tileZ) ok = false;
					survivors = __ballot_sync(0xffffffffu, ok);
					continue;
				}
				const int src = __ffs(survivors) - 1;
				survivors &= survivors - 1u;
				Bnd.mnx = __shfl_sync(0xffffffffu, dBnd.mnx, src); Bnd.mny = __shfl_sync(0xffffffffu, dBnd.mny, src); Bnd.mnz = __shfl_sync(0xffffffffu, dBnd.mnz, src);
				Bnd.mxx = __shfl_sync(0xffffffffu, dBnd.mxx, src); Bnd.mxy = __shfl_sync(0xffffffffu, dBnd.mxy, src); Bnd.mxz = __shfl_sync(0xffffffffu, dBnd.mxz, src);
				time0 = __shfl_sync(0xffffffffu, dTime0, src); time1 = __shfl_sync(0xffffffffu, dTime1, src);
				indexT0 = __shfl_sync(0xffffffffu, dIndexT0, src); indexT1 = __shfl_sync(0xffffffffu, dIndexT1, src);
				zminKey = __shfl_sync(0xffffffffu, dZminKey, src);
				maxCocX = __shfl_sync(0xffffffffu, dCocX, src); maxCocY = __shfl_sync(0xffffffffu, dCocY, src);
				// the pixel window of the flattened enumerations: the bound (grown by the circle of confusion) on the tile
				{
					const float bminx = Bnd.mnx - maxCocX, bmaxx = Bnd.mxx + maxCocX, bminy = Bnd.mny - maxCocY, bmaxy = Bnd.mxy + maxCocY;
					eX = (bmaxx >= (float)t.rx1) ? t.rx1 : min(lceilF(bmaxx), t.rx1);
					eY = (bmaxy >= (float)t.ry1) ? t.ry1 : min(lceilF(bmaxy), t.ry1);
					sX = (bminx < (float)t.rx0) ? t.rx0 : max(floorI(bminx), t.rx0);
					sY = (bminy < (float)t.ry0) ? t.ry0 : max(floorI(bminy), t.ry0);
				}
				if(sX >= eX || sY >= eY) continue;
				if(f.useDof)
				{
					// The reference walks the n lens cells and, for each, the pixels its shifted bound covers; of every such
					// pixel it tests the one sample that uses the cell -- and rejects it unless its time lies in the division.
					// Sample times are stratified by sample index, so for a moving micropolygon only the indices of the
					// division's window (widened by one) can pass that test: it is cheaper to walk (pixel, index) pairs and look
					// the cell up -- measured for every window length (config 3: 224 ms with windows up to n/4 only, 201 ms up
					// to n/2, 190 ms always).  Same candidates, same tests.
					const int w0 = max(0, indexT0 - 1), w1 = min(n, indexT1 + 1);
					byIndex = moving && !fastShutter && f.jitter && (w1 - w0)*(f.tune[5] ? f.tune[5] : 1) <= n;       // (without jitter every sample time is 0: no stratification)
					if(byIndex)
					{
						indexT0 = w0; cnt = w1 - w0;
						total = (eX - sX)*(eY - sY)*cnt;
						nRounds = (total + 1023) >> 10;
					}
					else nRounds = ((tileRows + bandRows - 1)/bandRows)*cellRounds;
				}
				else
				{
					// the reference's do-while visits at least one index per pixel
					byIndex = false;
					cnt = max(1, indexT1 - indexT0);
					total = (eX - sX)*(eY - sY)*cnt;
					nRounds = (total + 1023) >> 10;
				}
				round = 0;
				break;
			}
		}
		if(!last)
		{
			if(f.useDof && !byIndex)
			{
				const int band0 = (round / cellRounds)*bandRows, c0 = (round % cellRounds) << 5;
				const int by0 = t.ry0 + band0, by1 = min(t.ry1, by0 + bandRows);
				const int cell = c0 + lane;
				if(cell < n)
				{
					const float4 db = f.dofBounds[cell];
					const float bminx = Bnd.mnx - db.z*maxCocX, bmaxx = Bnd.mxx - db.x*maxCocX;
					const float bminy = Bnd.mny - db.w*maxCocY, bmaxy = Bnd.mxy - db.y*maxCocY;
					if((bmaxx >= (float)t.rx0) && (bmaxy >= (float)t.ry0) && (bminx < (float)t.rx1) && (bminy < (float)t.ry1))
					{
						const int ceX = (bmaxx >= (float)t.rx1) ? t.rx1 : min(lceilF(bmaxx), t.rx1);
						int ceY = (bmaxy >= (float)t.ry1) ? t.ry1 : min(lceilF(bmaxy), t.ry1);
						const int csX = (bminx < (float)t.rx0) ? t.rx0 : max(floorI(bminx), t.rx0);
						int csY = (bminy < (float)t.ry0) ? t.ry0 : max(floorI(bminy), t.ry0);
						csY = max(csY, by0); ceY = min(ceY, by1);
						for(int iY = csY; iY < ceY; ++iY)
							for(int iX = csX; iX < ceX; ++iX)
							{
								const int pixLocal = (iY - t.tileY0)*f.tileW + (iX - t.tileX0);
								// hierarchical z, per pixel: every sample of the pixel would fail "Bound.zmin > occlZ"
								if(zminKey > s.pixZ[pixLocal]) continue;
								// GetDofOffsetIndex(cell) = shuffledIndices[cell] of the pixel's shuffle pattern
								const int index = f.shufTab[(size_t)s.shufPat[pixLocal]*n + cell];
								const int idx = sampleIdx(f, s, iX - t.tileX0, iY - t.tileY0, index);
								if(moving)
								{
									const float time = s.time[idx];
									if(time < time0 || time > time1) continue;
								}
								const float x = s.posx[idx], y = s.posy[idx];
								if((x < bminx || x > bmaxx) || (y < bminy || y > bmaxy)) continue;
								if(zminKey > (uint32_t)(s.keys[idx] >> 32)) continue;
								movPush(f, ws, idx);
							}
					}
				}
			}
			else
			{
				const int Wp = eX - sX;
				const int k0 = round << 10, kEnd = min(total, k0 + 1024);
				// (k < 2^16 and the quotients are small: a float quotient is off by at most one, which the remainder corrects --
				// the integer divisions were 6.5 % of the kernel's instructions)
				const float rCnt = 1.0f/(float)cnt, rWp = 1.0f/(float)Wp;
				for(int k = k0 + lane; k < kEnd; k += 32)
				{
					int pi = (int)((float)k*rCnt), j = k - pi*cnt;
					if(j < 0) { --pi; j += cnt; } else if(j >= cnt) { ++pi; j -= cnt; }
					int qy = (int)((float)pi*rWp), qx = pi - qy*Wp;
					if(qx < 0) { --qy; qx += Wp; } else if(qx >= Wp) { ++qy; qx -= Wp; }
					const int iY = sY + qy, iX = sX + qx;
					const int pixLocal = (iY - t.tileY0)*f.tileW + (iX - t.tileX0);
					if(zminKey > s.pixZ[pixLocal]) continue;
					const int index = indexT0 + j;
					float bminx = Bnd.mnx, bmaxx = Bnd.mxx, bminy = Bnd.mny, bmaxy = Bnd.mxy;
					if(byIndex)
					{
						// the lens cell whose sample this is, its shifted bound, and whether the reference's walk over the pixels of
						// that bound reaches this pixel: iX in [floor(bminx), ceil(bmaxx)), i.e. bminx < iX + 1 and bmaxx > iX
						const int cell = f.shufInvTab[(size_t)s.shufPat[pixLocal]*n + index];
						const float4 db = f.dofBounds[cell];
						bminx = Bnd.mnx - db.z*maxCocX; bmaxx = Bnd.mxx - db.x*maxCocX;
						bminy = Bnd.mny - db.w*maxCocY; bmaxy = Bnd.mxy - db.y*maxCocY;
						if(!(bminx < (float)(iX + 1)) || !(bmaxx > (float)iX) || !(bminy < (float)(iY + 1)) || !(bmaxy > (float)iY)) continue;
					}
					const int idx = sampleIdx(f, s, iX - t.tileX0, iY - t.tileY0, index);
					const float time = s.time ? s.time[idx] : f.shutterOpen;
					if(moving && (time < time0 || time > time1)) continue;
					const float x = s.posx[idx], y = s.posy[idx];
					if((x < bminx || x > bmaxx) || (y < bminy || y > bmaxy)) continue;
					if(zminKey > (uint32_t)(s.keys[idx] >> 32)) continue;
					movPush(f, ws, idx);
				}
			}
			++round;
		}
		movDrain<PLAIN>(f, t, s, dc, c, ws, lane, last ? 1u : 32u);
		if(last) break;
	}
	return true;
}

// ---- resolve one sample: colour/opacity of the winning opaque hit, composite of the deep
// list (CqImagePixel::Combine, imagepixel.cpp:144-332; depth filter "min" only).
template<bool MBDOF>
__device__ __forceinline__ void hitUV(const DevFrame& f, const GridRec& g, uint32_t p, float2 pos, float2 dofOff, float time, float2& uv)
{
	if(g.flags & AQH_GRID_POINTS) { uv = make_float2(0.f, 0.f); return; }     // a disc has no parametrisation: constant shading
	if(!MBDOF)
	{
		// no motion and no depth of field anywhere in the frame: the vertices are those of the grid
		const uint32_t cu = g.cu_cv & 0xffffu;
		const float4 P0 = f.P4[p], P1 = f.P4[p+1], P2 = f.P4[p+cu+1], P3 = f.P4[p+cu+2];
		const float qx[4] = {P0.x, P1.x, P2.x, P3.x}, qy[4] = {P0.y, P1.y, P2.y, P3.y}, qz[4] = {P0.z, P1.z, P2.z, P3.z};
		HitCache c;
		cachePointInPolyTest(c, qx, qy, qz, 0xE4 /* any valid code: only the uv part is used */);
		uv = invBilinear(c.Ax, c.Ay, c.Ex, c.Ey, c.Fx, c.Fy, c.Gx, c.Gy, c.linear, pos.x, pos.y);
		return;
	}
	MovingMP m;
	m.p = p; m.g = g; m.cu = g.cu_cv & 0xffffu; m.nverts = g.nverts; m.nkeys = g.nkeys_koff & 0xffu;
	m.times = f.keyTimes + (g.nkeys_koff >> 8);
	float px[4], py[4], pz[4];
	B2 dummy;
	samplePoints<false>(f, m, nullptr, m.nkeys > 1, dummy, make_float2(0.f, 0.f), make_float2(0.f, 0.f), pos, dofOff, time, px, py, pz, false);
	HitCache c;
	cachePointInPolyTest(c, px, py, pz, 0xE4 /* any valid code: only the uv part is used */);
	uv = invBilinear(c.Ax, c.Ay, c.Ex, c.Ey, c.Fx, c.Fy, c.Gx, c.Gy, c.linear, pos.x, pos.y);
}

// Fast path: depth filter "min" (imagepixel.cpp:144-262, 301-318).
template<bool MBDOF>
__device__ __forceinline__ void resolveSampleMin(const DevFrame& f, const TileCtx& t, const HideSmem& s, const DeepCtx& dc, int idx,
                              float out[7], bool& valid, uint32_t& pNear)
{
	const unsigned long long key = s.keys[idx];
	const uint32_t p = (uint32_t)key;
	pNear = p;                       // the nearest entry hands its arbitrary output variables to the sample (imagepixel.cpp:249-251)
	const bool haveOpaque = (p != 0xffffffffu);
	const float2 pos = make_float2(s.posx[idx], s.posy[idx]);
	float col[3] = {0.f, 0.f, 0.f}, opa[3] = {0.f, 0.f, 0.f};
	float depth = 0.f;
	bool opaqueMatte = false;
	if(haveOpaque)
	{
		const float4 a = f.P4[p];
		const GridRec g = f.grids[infoOf(a) & VINFO_GRID_MASK];
		float2 uv;
		const float2 dofOff = s.dof ? s.dof[idx] : make_float2(0.f, 0.f);
		const float time = s.time ? s.time[idx] : 0.f;
		hitUV<MBDOF>(f, g, p, pos, dofOff, time, uv);
		shadeHit(f, g, p, uv, col, opa);
		depth = keyDepth((uint32_t)(key >> 32));
		opaqueMatte = (g.flags & AQH_GRID_MATTE) != 0;
	}
	const uint32_t word = s.head ? s.head[idx] : 0u;
	const uint32_t nDeep = word & DEEP_COUNT_MASK;
	if(nDeep == 0u)
	{
		valid = haveOpaque;
		if(opaqueMatte) { col[0] = col[1] = col[2] = 0.f; opa[0] = opa[1] = opa[2] = 0.f; }  // imagepixel.cpp:308-318
		out[0] = col[0]; out[1] = col[1]; out[2] = col[2]; out[3] = opa[0]; out[4] = opa[1]; out[5] = opa[2]; out[6] = depth;
		return;
	}
	// back-to-front over {opaque hit} ∪ deep list, i.e. descending (depth, order)
	float sc[3] = {0.f, 0.f, 0.f}, so[3] = {0.f, 0.f, 0.f};
	float opaqueDepth0 = haveOpaque ? depth : FLT_MAX;      // opaqueDepths[0] = occlZ
	if(haveOpaque)
	{
		if(opaqueMatte)
		{
#pragma unroll
			for(int k = 0; k < 3; ++k) { sc[k] = (1.f-opa[k])*sc[k] + opa[k]*0.0f; so[k] = (1.f-col[k])*so[k] + col[k]*0.0f; }
		}
		else
		{
#pragma unroll
			for(int k = 0; k < 3; ++k)
			{
				sc[k] = (sc[k] * (1.0f - fminf(fmaxf(opa[k], 0.0f), 1.0f))) + col[k];
				so[k] = ((1.0f - so[k]) * opa[k]) + so[k];
			}
		}
		if(opa[0] >= f.zthr[0] && opa[1] >= f.zthr[1] && opa[2] >= f.zthr[2]) opaqueDepth0 = depth;
	}
	// One walk collects up to DEEP_SORT entries into a register-resident, descending (depth, submission) order (the
	// in-line records of neighbouring samples are neighbours in memory); longer lists fall back to repeated selection
	// of the farthest not yet composited entry.
	constexpr int DEEP_SORT = 8;
	unsigned long long ks[DEEP_SORT];
	uint32_t sl[DEEP_SORT];
#pragma unroll
	for(int j = 0; j < DEEP_SORT; ++j) { ks[j] = 0ull; sl[j] = DEEP_NIL; }
	int nList = 0;
	const bool shortList = nDeep <= (uint32_t)(DEEP_SORT/2);
	// (a list of at most four entries -- the common case -- only ever touches the first four slots)
	auto insert = [&](const uint2 A, uint32_t ce)
	{
		unsigned long long k = ((unsigned long long)depthKey(__uint_as_float(A.x)) << 32) | A.y;
#pragma unroll
		for(int j = 0; j < DEEP_SORT; ++j)
		{
			if(j == DEEP_SORT/2 && shortList) break;
			// strict '>' keeps equal keys (a hit stored twice on a time sub-bound boundary) both in the list
			if(k > ks[j] || (sl[j] == DEEP_NIL && ce != DEEP_NIL))
			{
				const unsigned long long tk = ks[j]; const uint32_t ts = sl[j];
				ks[j] = k; sl[j] = ce; k = tk; ce = ts;
			}
		}
		++nList;
	};
	for(uint32_t e = deepFirst(word); e != DEEP_NIL; e = deepNext(dc, word, e))
		insert(dc.A[deepAt(dc, s.nsP, idx, e)], e);
	unsigned long long prev = ~0ull;
	for(int step = 0; ; ++step)
	{
		uint32_t bestSlot = DEEP_NIL;
		if(nList <= DEEP_SORT)
		{
			if(step >= nList) break;
#pragma unroll
			for(int j = 0; j < DEEP_SORT; ++j) if(j == step) bestSlot = sl[j];
		}
		else
		{
			// farthest not yet composited entry: largest (depthKey, p) strictly below prev
			unsigned long long best = 0;
			for(uint32_t e = deepFirst(word); e != DEEP_NIL; e = deepNext(dc, word, e))
			{
				const uint2 A = dc.A[deepAt(dc, s.nsP, idx, e)];
				unsigned long long k = ((unsigned long long)depthKey(__uint_as_float(A.x)) << 32) | A.y;
				if(k < prev && (bestSlot == DEEP_NIL || k > best)) { best = k; bestSlot = e; }
			}
			if(bestSlot == DEEP_NIL) break;
			prev = best;
		}
		// one entry over the running composite (imagepixel.cpp:191-262)
		const size_t at = deepAt(dc, s.nsP, idx, bestSlot);
		const uint2 A = dc.A[at];
		const uint4 B = dc.B[at];
		float c2[3], o2[3];
		shadeAt(f, B.z, B.w & 0xffffu, (B.w & HITF_SMOOTH) != 0, make_float2(__uint_as_float(B.x), __uint_as_float(B.y)), c2, o2);
		pNear = A.y;                 // composited back to front: the last entry is the nearest
		if(B.w & HITF_MATTE)
		{
#pragma unroll
			for(int k = 0; k < 3; ++k) { sc[k] = (1.f-o2[k])*sc[k] + o2[k]*0.0f; so[k] = (1.f-c2[k])*so[k] + c2[k]*0.0f; }
		}
		else
		{
#pragma unroll
			for(int k = 0; k < 3; ++k)
			{
				sc[k] = (sc[k] * (1.0f - fminf(fmaxf(o2[k], 0.0f), 1.0f))) + c2[k];
				so[k] = ((1.0f - so[k]) * o2[k]) + so[k];
			}
		}
		if(o2[0] >= f.zthr[0] && o2[1] >= f.zthr[1] && o2[2] >= f.zthr[2]) opaqueDepth0 = __uint_as_float(A.x);
	}
	valid = true;
	out[0] = sc[0]; out[1] = sc[1]; out[2] = sc[2]; out[3] = so[0]; out[4] = so[1]; out[5] = so[2];
	out[6] = opaqueDepth0;
}

template<bool MBDOF>
__device__ __forceinline__ void resolveSampleGeneral(const DevFrame& f, const TileCtx& t, const HideSmem& s, const DeepCtx& dc, int idx,
                              float out[7], bool& valid, uint32_t& pNear)
{
	// keys[nsP + idx]: nearest opaque hit; keys[idx]: the occlusion depth occlZ (the same array unless the midpoint depth
	// filter keeps the SECOND nearest opaque depth there, bucketprocessor.cpp:1502-1529)
	const unsigned long long key = s.keys[idx + (f.midpointZ ? s.nsP : 0)];
	const float occlZ = keyDepth((uint32_t)(s.keys[idx] >> 32));
	const uint32_t p = (uint32_t)key;
	pNear = p;
	const bool haveOpaque = (p != 0xffffffffu);
	const float2 pos = make_float2(s.posx[idx], s.posy[idx]);
	float col[3] = {0.f, 0.f, 0.f}, opa[3] = {0.f, 0.f, 0.f};
	float depth = 0.f;
	bool opaqueMatte = false;
	if(haveOpaque)
	{
		const float4 a = f.P4[p];
		const GridRec g = f.grids[infoOf(a) & VINFO_GRID_MASK];
		float2 uv;
		const float2 dofOff = s.dof ? s.dof[idx] : make_float2(0.f, 0.f);
		const float time = s.time ? s.time[idx] : 0.f;
		hitUV<MBDOF>(f, g, p, pos, dofOff, time, uv);
		shadeHit(f, g, p, uv, col, opa);
		depth = keyDepth((uint32_t)(key >> 32));
		opaqueMatte = (g.flags & AQH_GRID_MATTE) != 0;
	}
	const uint32_t word = s.head ? s.head[idx] : 0u;
	const uint32_t head = deepFirst(word);
	if(head == DEEP_NIL)
	{
		// opaque-only sample: imagepixel.cpp:304-327
		valid = haveOpaque;
		if(opaqueMatte) { col[0] = col[1] = col[2] = 0.f; opa[0] = opa[1] = opa[2] = 0.f; }
		if(haveOpaque && f.depthFilter == AQH_DEPTHFILTER_MIDPOINT) depth = 0.5f*(depth + occlZ);
		out[0] = col[0]; out[1] = col[1]; out[2] = col[2]; out[3] = opa[0]; out[4] = opa[1]; out[5] = opa[2]; out[6] = depth;
		return;
	}
	// ---- CqImagePixel::Combine (imagepixel.cpp:144-300): the valid opaque hit joins the list, the list is
	// ordered by depth and composited back to front.  One walk over the sample's list collects up to
	// DEEP_SORT entries into a register-resident, descending (depth, submission) order -- the pointer chase
	// through the pool is paid once; longer lists fall back to repeated selection.
	constexpr int DEEP_SORT = 8;
	constexpr uint32_t NIL = 0xffffffffu, OPQ = 0xfffffffeu;      // OPQ: "the entry is the opaque hit"
	unsigned long long ks[DEEP_SORT];
	uint32_t sl[DEEP_SORT];
#pragma unroll
	for(int j = 0; j < DEEP_SORT; ++j) { ks[j] = 0ull; sl[j] = NIL; }
	int nList = 0;
	{
		uint32_t e = haveOpaque ? OPQ : head;
		while(e != NIL)
		{
			unsigned long long k; uint32_t nextE;
			if(e == OPQ) { k = key; nextE = head; }
			else
			{
				const uint2 A = dc.A[deepAt(dc, s.nsP, idx, e)];
				k = ((unsigned long long)depthKey(__uint_as_float(A.x)) << 32) | A.y;
				nextE = deepNext(dc, word, e);
			}
			uint32_t ce = e;
#pragma unroll
			for(int j = 0; j < DEEP_SORT; ++j)
			{
				// strict '>' keeps equal keys (a hit stored twice on a time sub-bound boundary) both in the list
				if(k > ks[j] || (sl[j] == NIL && ce != NIL))
				{
					const unsigned long long tk = ks[j]; const uint32_t ts = sl[j];
					ks[j] = k; sl[j] = ce; k = tk; ce = ts;
				}
			}
			++nList;
			e = nextE;
		}
	}
	float sc[3] = {0.f, 0.f, 0.f}, so[3] = {0.f, 0.f, 0.f};
	float opaqueDepth0 = occlZ, opaqueDepth1 = FLT_MAX, maxOpaqueDepth = FLT_MAX;
	float totDepth = 0.0f; int totCount = 0;
	const bool average = f.depthFilter == AQH_DEPTHFILTER_AVERAGE;
	// pass 0: back to front (descending) compositing; pass 1 (depth filter "average" only): front to back sum
	// of the depths of the entries that reach the z threshold in ANY channel (imagepixel.cpp:279-296)
	for(int pass = 0; pass < (average ? 2 : 1); ++pass)
	{
		unsigned long long prev = (pass == 0) ? ~0ull : 0ull;
		for(int step = 0; ; ++step)
		{
			uint32_t bestSlot = NIL;
			if(nList <= DEEP_SORT)
			{
				if(step >= nList) break;
				const int want = (pass == 0) ? step : nList - 1 - step;
#pragma unroll
				for(int j = 0; j < DEEP_SORT; ++j) if(j == want) bestSlot = sl[j];
			}
			else
			{
				// next entry in (depthKey, p) order strictly beyond prev
				unsigned long long best = 0;
				uint32_t e = haveOpaque ? OPQ : head;
				while(e != NIL)
				{
					unsigned long long k; uint32_t nextE;
					if(e == OPQ) { k = key; nextE = head; }
					else
					{
						const uint2 A = dc.A[deepAt(dc, s.nsP, idx, e)];
						k = ((unsigned long long)depthKey(__uint_as_float(A.x)) << 32) | A.y;
						nextE = deepNext(dc, word, e);
					}
					const bool beyond = (pass == 0) ? (k < prev) : (k > prev);
					const bool better = (pass == 0) ? (k > best) : (k < best);
					if(beyond && (bestSlot == NIL || better)) { best = k; bestSlot = e; }
					e = nextE;
				}
				if(bestSlot == NIL) break;
				prev = best;
			}
			float c2[3], o2[3], d2;
			bool matte;
			if(bestSlot == OPQ)
			{
#pragma unroll
				for(int k = 0; k < 3; ++k) { c2[k] = col[k]; o2[k] = opa[k]; }
				d2 = depth; matte = opaqueMatte;
				if(pass == 0) pNear = p;
			}
			else
			{
				const size_t at = deepAt(dc, s.nsP, idx, bestSlot);
				const uint2 A = dc.A[at];
				const uint4 B = dc.B[at];
				if(pass == 0) pNear = A.y;
				shadeAt(f, B.z, B.w & 0xffffu, (B.w & HITF_SMOOTH) != 0, make_float2(__uint_as_float(B.x), __uint_as_float(B.y)), c2, o2);
				d2 = __uint_as_float(A.x);
				matte = (B.w & HITF_MATTE) != 0;
			}
			if(pass == 1)
			{
				// the nearest entry shares its data with the composited result in the reference (the hit is
				// an index into the pixel's data pool): its test sees the composited opacity
				const float* od = (step == 0) ? so : o2;
				if(od[0] >= f.zthr[0] || od[1] >= f.zthr[1] || od[2] >= f.zthr[2]) { totDepth += d2; totCount++; }
				continue;
			}
			if(matte)
			{
#pragma unroll
				for(int k = 0; k < 3; ++k) { sc[k] = (1.f-o2[k])*sc[k] + o2[k]*0.0f; so[k] = (1.f-c2[k])*so[k] + c2[k]*0.0f; }
			}
			else
			{
#pragma unroll
				for(int k = 0; k < 3; ++k)
				{
					sc[k] = (sc[k] * (1.0f - fminf(fmaxf(o2[k], 0.0f), 1.0f))) + c2[k];
					so[k] = ((1.0f - so[k]) * o2[k]) + so[k];
				}
			}
			if(o2[0] >= f.zthr[0] && o2[1] >= f.zthr[1] && o2[2] >= f.zthr[2])
			{
				opaqueDepth1 = opaqueDepth0;
				opaqueDepth0 = d2;
				if(!(maxOpaqueDepth < FLT_MAX)) maxOpaqueDepth = d2;
			}
		}
	}
	valid = true;
	out[0] = sc[0]; out[1] = sc[1]; out[2] = sc[2]; out[3] = so[0]; out[4] = so[1]; out[5] = so[2];
	float zout = opaqueDepth0;
	if(f.depthFilter == AQH_DEPTHFILTER_MIDPOINT) zout = (nList > 1) ? ((opaqueDepth0 + opaqueDepth1) * 0.5f) : FLT_MAX;
	else if(f.depthFilter == AQH_DEPTHFILTER_MAX) zout = maxOpaqueDepth;
	else if(average) zout = totDepth / (float)totCount;
	out[6] = zout;
}

// ---- CSG: CqCSGTreeNode::ProcessTree / ProcessSampleList / EvaluateState (csgtree.cpp:144-351) on the sample's depth-sorted
// list.  Samples with a list in a frame that has CSG grids take this path (rare: out of line, arrays in local memory).
#define CSG_MAX_LIST 48
__device__ __forceinline__ bool csgEvaluate(int type, uint32_t state, int nKids)
{
	const uint32_t all = nKids >= 32 ? 0xffffffffu : ((1u << nKids) - 1u);
	if(type == AQH_CSG_UNION) return (state & all) != 0u;
	if(type == AQH_CSG_INTERSECTION) return (state & all) == all;
	if(type == AQH_CSG_DIFFERENCE) return (state & 1u) && !(state & all & ~1u);
	return false;
}
template<bool MBDOF>
__device__ __noinline__ void resolveSampleCSG(const DevFrame& f, const HideSmem s, const DeepCtx dc, int idx,
                                              float out[7], bool& valid, uint32_t& pNear)
{
	constexpr uint32_t NIL = 0xffffffffu, OPQ = 0xfffffffeu;
	const unsigned long long key = s.keys[idx + (f.midpointZ ? s.nsP : 0)];
	const float occlZ = keyDepth((uint32_t)(s.keys[idx] >> 32));
	const uint32_t p = (uint32_t)key;
	const bool haveOpaque = (p != NIL);
	const float2 pos = make_float2(s.posx[idx], s.posy[idx]);
	float col[3] = {0.f, 0.f, 0.f}, opa[3] = {0.f, 0.f, 0.f};
	float depth = 0.f;
	bool opaqueMatte = false;
	if(haveOpaque)
	{
		const float4 a = f.P4[p];
		const GridRec g = f.grids[infoOf(a) & VINFO_GRID_MASK];
		float2 uv;
		const float2 dofOff = s.dof ? s.dof[idx] : make_float2(0.f, 0.f);
		const float time = s.time ? s.time[idx] : 0.f;
		hitUV<MBDOF>(f, g, p, pos, dofOff, time, uv);
		shadeHit(f, g, p, uv, col, opa);
		depth = keyDepth((uint32_t)(key >> 32));
		opaqueMatte = (g.flags & AQH_GRID_MATTE) != 0;
	}
	// the list in ascending (depth, submission) order, the valid opaque hit included (imagepixel.cpp:157-164)
	unsigned long long ks[CSG_MAX_LIST];
	uint32_t sl[CSG_MAX_LIST];
	int nd[CSG_MAX_LIST];
	int cnt = 0;
	{
		const uint32_t word = s.head[idx];
		uint32_t e = haveOpaque ? OPQ : deepFirst(word);
		while(e != NIL)
		{
			unsigned long long k; uint32_t nextE; int node = -1;
			if(e == OPQ) { k = key; nextE = deepFirst(word); }
			else
			{
				const uint2 A = dc.A[deepAt(dc, s.nsP, idx, e)];
				k = ((unsigned long long)depthKey(__uint_as_float(A.x)) << 32) | A.y;
				nextE = deepNext(dc, word, e);
				const uint32_t gi = infoOf(f.P4[A.y]) & VINFO_GRID_MASK;
				if(f.grids[gi].flags & AQH_GRID_USES_CSG) node = f.gridCsg[gi];
			}
			if(cnt >= CSG_MAX_LIST) { atomicOr(f.errorFlags, 8u); break; }
			int j = cnt++;
			while(j > 0 && ks[j-1] > k) { ks[j] = ks[j-1]; sl[j] = sl[j-1]; nd[j] = nd[j-1]; --j; }
			ks[j] = k; sl[j] = e; nd[j] = node;
			e = nextE;
		}
	}
	// as long as any entry belongs to a CSG node, run that node's whole tree over the list (imagepixel.cpp:166-189)
	for(int guard = 0; guard < 64; ++guard)
	{
		int first = -1;
		for(int j = 0; j < cnt && first < 0; ++j) if(nd[j] >= 0) first = j;
		if(first < 0) break;
		int root = nd[first];
		while(f.csgParent[root] >= 0) root = f.csgParent[root];
		if(f.csgType[root] == AQH_CSG_PRIMITIVE)
		{
			// CqCSGNodePrimitive::ProcessSampleList: a lone primitive just releases its samples
			for(int j = 0; j < cnt; ++j) if(nd[j] == root) nd[j] = -1;
			continue;
		}
		for(int oi = 0; oi < f.nCsgOrder; ++oi)
		{
			const int N = f.csgOrder[oi];
			int r = N;
			while(f.csgParent[r] >= 0) r = f.csgParent[r];
			if(r != root) continue;
			// ProcessSampleList of node N (its non-primitive children come earlier in csgOrder)
			const int type = f.csgType[N], nKids = f.csgKids[N];
			const int promoted = (f.csgParent[N] >= 0) ? N : -1;
			uint32_t state = 0u;
			bool current = csgEvaluate(type, state, nKids);
			int w = 0;
			for(int j = 0; j < cnt; ++j)
			{
				const int n = nd[j];
				const int ci = (n >= 0 && f.csgParent[n] == N) ? f.csgSlot[n] : -1;
				bool keep = true;
				int newNode = n;
				if(ci >= 0)
				{
					state ^= 1u << ci;
					const bool next = csgEvaluate(type, state, nKids);
					if(next == current) keep = false;          // the solid's state does not change here: the wall is not visible
					else { current = next; newNode = promoted; }
				}
				if(keep) { ks[w] = ks[j]; sl[w] = sl[j]; nd[w] = newNode; ++w; }
			}
			cnt = w;
		}
	}
	if(cnt == 0) { valid = false; pNear = NIL; return; }     // nothing left and no opaque hit: the sample stays empty
	// ---- composite what is left back to front (imagepixel.cpp:191-300), exactly like resolveSampleGeneral
	float sc[3] = {0.f, 0.f, 0.f}, so[3] = {0.f, 0.f, 0.f};
	float opaqueDepth0 = occlZ, opaqueDepth1 = FLT_MAX, maxOpaqueDepth = FLT_MAX;
	float totDepth = 0.0f; int totCount = 0;
	const bool average = f.depthFilter == AQH_DEPTHFILTER_AVERAGE;
	for(int pass = 0; pass < (average ? 2 : 1); ++pass)
		for(int step = 0; step < cnt; ++step)
		{
			const int j = (pass == 0) ? cnt - 1 - step : step;
			float c2[3], o2[3], d2;
			bool matte;
			if(sl[j] == OPQ)
			{
#pragma unroll
				for(int k = 0; k < 3; ++k) { c2[k] = col[k]; o2[k] = opa[k]; }
				d2 = depth; matte = opaqueMatte;
			}
			else
			{
				const size_t at = deepAt(dc, s.nsP, idx, sl[j]);
				const uint2 A = dc.A[at];
				const uint4 B = dc.B[at];
				shadeAt(f, B.z, B.w & 0xffffu, (B.w & HITF_SMOOTH) != 0, make_float2(__uint_as_float(B.x), __uint_as_float(B.y)), c2, o2);
				d2 = __uint_as_float(A.x);
				matte = (B.w & HITF_MATTE) != 0;
			}
			if(pass == 1)
			{
				const float* od = (step == 0) ? so : o2;      // the nearest entry shares its data with the composited result
				if(od[0] >= f.zthr[0] || od[1] >= f.zthr[1] || od[2] >= f.zthr[2]) { totDepth += d2; totCount++; }
				continue;
			}
			if(matte)
			{
#pragma unroll
				for(int k = 0; k < 3; ++k) { sc[k] = (1.f-o2[k])*sc[k] + o2[k]*0.0f; so[k] = (1.f-c2[k])*so[k] + c2[k]*0.0f; }
			}
			else
			{
#pragma unroll
				for(int k = 0; k < 3; ++k)
				{
					sc[k] = (sc[k] * (1.0f - fminf(fmaxf(o2[k], 0.0f), 1.0f))) + c2[k];
					so[k] = ((1.0f - so[k]) * o2[k]) + so[k];
				}
			}
			if(o2[0] >= f.zthr[0] && o2[1] >= f.zthr[1] && o2[2] >= f.zthr[2])
			{
				opaqueDepth1 = opaqueDepth0;
				opaqueDepth0 = d2;
				if(!(maxOpaqueDepth < FLT_MAX)) maxOpaqueDepth = d2;
			}
		}
	valid = true;
	pNear = (sl[0] == OPQ) ? p : dc.A[deepAt(dc, s.nsP, idx, sl[0])].y;
	out[0] = sc[0]; out[1] = sc[1]; out[2] = sc[2]; out[3] = so[0]; out[4] = so[1]; out[5] = so[2];
	float zout = opaqueDepth0;
	if(f.depthFilter == AQH_DEPTHFILTER_MIDPOINT) zout = (cnt > 1) ? ((opaqueDepth0 + opaqueDepth1) * 0.5f) : FLT_MAX;
	else if(f.depthFilter == AQH_DEPTHFILTER_MAX) zout = maxOpaqueDepth;
	else if(average) zout = totDepth / (float)totCount;
	out[6] = zout;
}

// Depth filter "min" without the midpoint bookkeeping is the common case and keeps its own lean code; every
// other depth filter goes through the general restatement of CqImagePixel::Combine above; samples with a hit list in a
// frame with CSG solids go through the CSG resolve.
template<bool MBDOF, bool DFGEN, bool PL = false>
__device__ __forceinline__ void resolveSample(const DevFrame& f, const TileCtx& t, const HideSmem& s, const DeepCtx& dc, int idx,
                                              float out[7], bool& valid, uint32_t& pNear)
{
	if(!PL && f.anyCSG && s.head && (s.head[idx] & DEEP_COUNT_MASK))
	{
		// results through temporaries: only they (not the caller's registers) have their address taken by the call
		float o[7]; bool v; uint32_t pn;
		resolveSampleCSG<MBDOF>(f, s, dc, idx, o, v, pn);
#pragma unroll
		for(int k = 0; k < 7; ++k) out[k] = o[k];
		valid = v; pNear = pn;
		return;
	}
	if(!DFGEN) resolveSampleMin<MBDOF>(f, t, s, dc, idx, out, valid, pNear);
	else resolveSampleGeneral<MBDOF>(f, t, s, dc, idx, out, valid, pNear);
}

// Per-tap inclusion bits of one sample (bucketprocessor.cpp:609-612), evaluated with the true
// pixel coordinates: the sample of pixel (X,Y) belongs to filter tap (fx,fy) of output pixel
// (X-fx, Y-fy).  Bits 0-14: x taps, 15-29: y taps, 31: the sample holds a valid hit.
__device__ __forceinline__ uint32_t tapMask(const DevFrame& f, float posx, float posy, int X, int Y, bool valid)
{
	uint32_t mask = valid ? 0x80000000u : 0u;
	for(int fx = -f.shiftX; fx <= f.shiftX; ++fx)
	{
		float vx = posx - ((float)(X - fx) + 0.5f);
		if(vx >= -f.xfwo2 && vx <= f.xfwo2) mask |= 1u << (fx + f.shiftX);
	}
	for(int fy = -f.shiftY; fy <= f.shiftY; ++fy)
	{
		float vy = posy - ((float)(Y - fy) + 0.5f);
		if(vy >= -f.yfwo2 && vy <= f.yfwo2) mask |= 1u << (15 + fy + f.shiftY);
	}
	return mask;
}

// The compact per-slot word of the sample planes (hider_device.h): x-tap bits, y-tap bits, valid bit.
__device__ __forceinline__ uint32_t packMask(const DevFrame& f, uint32_t m)
{
	const int nx = 2*f.shiftX + 1, ny = 2*f.shiftY + 1;
	return (m & ((1u << nx) - 1u)) | (((m >> 15) & ((1u << ny) - 1u)) << nx) | ((m >> 31) << (nx + ny));
}
// Stored INVERTED (a set bit = "not in this tap" / "no valid hit"), so that the filter's inclusion test is
// a single AND-and-test-zero; padding slots hold all ones.
__device__ __forceinline__ void storeMask(const DevFrame& f, size_t at, uint32_t compact)
{
	compact = ~compact;
	if(f.maskBytes == 1) f.maskPlane[at] = (unsigned char)compact;
	else if(f.maskBytes == 2) reinterpret_cast<uint16_t*>(f.maskPlane)[at] = (uint16_t)compact;
	else reinterpret_cast<uint32_t*>(f.maskPlane)[at] = compact;
}
__device__ __forceinline__ uint32_t loadMask(const DevFrame& f, size_t at)
{
	if(f.maskBytes == 1) return (~(uint32_t)f.maskPlane[at]) & 0xffu;
	if(f.maskBytes == 2) return (~(uint32_t)reinterpret_cast<const uint16_t*>(f.maskPlane)[at]) & 0xffffu;
	return ~reinterpret_cast<const uint32_t*>(f.maskPlane)[at];
}

// Per-phase timing of k_hide for development builds (python -m aqsis_b200.build --phase-timing, i.e. -DAQH_PHASE_TIMING):
// every warp adds, per phase, the cycles it worked and the cycles it then waited at the phase's barrier to
// f.counters[4 + 2*phase], [5 + 2*phase].  Phases: 0 prepare, 1 opaque pass, 2 deep pass, 3 resolve, 4 tile fetch,
// 5 refresh between the passes, 6 occlusion image.
#ifdef AQH_PHASE_TIMING
#define PHASE_BARRIER(ph) do { const long long a_ = clock64(); __syncthreads(); const long long b_ = clock64(); \
	if(lane == 0) { atomicAdd(&s_ph[2*(ph)], (unsigned long long)(a_ - tPh)); atomicAdd(&s_ph[2*(ph) + 1], (unsigned long long)(b_ - a_)); } \
	tPh = b_; } while(0)
#else
#define PHASE_BARRIER(ph) __syncthreads()
#endif

// ------------------------------------------------------------------------------------
// k_hide: persistent CTAs pull tiles from a counter.  Inside a tile every warp runs its own
// pipeline -- grab RECS_PER_WARP micropolygons of the tile's bin, set them up (one lane each,
// records in the warp's shared-memory slots), then sample them one after the other with all 32
// lanes -- so there is no CTA-wide barrier inside the micropolygon loop.
// DFGEN: a depth filter other than "min" (kept out of the common kernels: its bookkeeping costs registers)
template<bool MBDOF, int THREADS, bool PARTIALS, bool DFGEN, bool PLAIN = false>
__global__ void __launch_bounds__(THREADS, (!MBDOF && THREADS < 512) ? 1024/THREADS : 2) k_hide(const __grid_constant__ DevFrame f, uint32_t slotBeg, uint32_t slotEnd, uint32_t* cursor)
{
	extern __shared__ __align__(16) unsigned char smemRaw[];
	__shared__ uint32_t s_tile;
	__shared__ uint32_t s_deepCount;
	__shared__ uint32_t s_next, s_resNext;
	__shared__ uint32_t s_tileZ, s_dirty, s_lastRef;
	constexpr int NWARPS = THREADS/32;
	HideSmem s = carveSmem(f, smemRaw, NWARPS*RECS_PER_WARP, MBDOF ? NWARPS*sizeof(MovScratch) : 0);
	MovScratch* ws = MBDOF ? reinterpret_cast<MovScratch*>(reinterpret_cast<unsigned char*>(s.mov) + (size_t)(threadIdx.x >> 5)*sizeof(MovScratch)) : nullptr;
	s.tileZ = &s_tileZ; s.dirty = &s_dirty;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int n = f.n, xs = f.xs, ys = f.ys;
	StaticRec* myRecs = s.recs + warp*RECS_PER_WARP;
	DeepCtx dc;
	// this CTA's part of the transparent hit pool: the in-line slots of the tile's samples, then the overflow records
	dc.ovBase = (uint32_t)DEEP_INLINE*(uint32_t)s.nsP;
	dc.A = f.deepA + (size_t)blockIdx.x*((size_t)dc.ovBase + f.deepCapPerCta);
	dc.B = f.deepB + (size_t)blockIdx.x*((size_t)dc.ovBase + f.deepCapPerCta);
	dc.ovNext = f.deepNext + (size_t)blockIdx.x*f.deepCapPerCta;
	dc.ovCap = f.deepCapPerCta;
	dc.ovCount = &s_deepCount;
	for(int i = tid; i < n; i += THREADS) s.subOfs[i] = (uint16_t)((i / xs)*s.stride + i % xs);
#ifdef AQH_PHASE_TIMING
	__shared__ unsigned long long s_ph[14];
	if(tid < 14) s_ph[tid] = 0;
	__syncthreads();
	long long tPh = clock64();
#endif
	for(;;)
	{
		PHASE_BARRIER(3);
		// one band of tile rows per launch: active tiles [slotBeg, slotEnd), handed out through the band's own cursor
		if(tid == 0) { s_tile = slotBeg + atomicAdd(cursor, 1u); s_deepCount = 0; s_next = 0; s_resNext = 0; }
		PHASE_BARRIER(4);
		const uint32_t slot = s_tile;
		if(slot >= slotEnd) break;
		// an incremental flush leaves tiles without new micropolygons as they are (their occlusion image is already there)
		if((!PLAIN && f.zOnly) && (PLAIN ? (unsigned long long*)nullptr : f.zKeys) && f.binOffset[slot+1] == f.binOffset[slot]) continue;
		const uint32_t tile = f.activeTiles[slot];
		TileCtx t;
		t.tileX0 = f.sx0 + (int)(tile % f.ntx)*f.tileW;
		t.tileY0 = f.sy0 + (int)(tile / f.ntx)*f.tileH;
		t.rx0 = t.tileX0; t.ry0 = t.tileY0;
		t.rx1 = min(t.tileX0 + f.tileW, f.sx0 + f.sw);
		t.ry1 = min(t.tileY0 + f.tileH, f.sy0 + f.sh);
		// ---- Prepare_bucket: CqImagePixel::clear + setSamples (imagepixel.cpp:105-122, 334-359)
		// every slot (padding included) starts empty and far outside any bound ...
		for(int idx = tid; idx < s.nsP; idx += THREADS)
		{
			s.keys[idx] = KEY_EMPTY;
			if((!PLAIN && f.midpointZ)) s.keys[s.nsP + idx] = KEY_EMPTY;
			if(s.head) s.head[idx] = 0u;
			s.posx[idx] = -1e30f;
			s.posy[idx] = -1e30f;
		}
		PHASE_BARRIER(0);
		// ... then one warp per pixel of the tile, lanes over its samples (no per-sample index arithmetic)
		for(int pix = warp; pix < f.tileW*f.tileH; pix += NWARPS)
		{
			const int ply = pix / f.tileW, plx = pix - ply*f.tileW;
			const int X = t.tileX0 + plx, Y = t.tileY0 + ply;
			if(X >= t.rx1 || Y >= t.ry1) continue;
			const size_t pp = (size_t)(Y - f.sy0)*f.sw + (X - f.sx0);
			const size_t plane = (size_t)f.sw*f.sh;
			const int patPos = f.patPlanes[plane + pp];
			const int patT = s.time ? f.patPlanes[3*plane + pp] : 0;
			const int patS = s.dof ? f.patPlanes[pp] : 0, patD = s.dof ? f.patPlanes[2*plane + pp] : 0;
			const int base = (ply*ys)*s.stride + plx*xs;
			for(int i = lane; i < n; i += 32)
			{
				const int idx = base + s.subOfs[i];
				const float2 o = f.posTab[(size_t)patPos*n + i];
				s.posx[idx] = (float)X + o.x;
				s.posy[idx] = (float)Y + o.y;
				if((PLAIN ? (unsigned long long*)nullptr : f.zKeys))
				{
					// the occlusion state earlier flushes of this frame left behind
					s.keys[idx] = (PLAIN ? (unsigned long long*)nullptr : f.zKeys)[pp*n + i];
					if((!PLAIN && f.midpointZ)) s.keys[s.nsP + idx] = (PLAIN ? (unsigned long long*)nullptr : f.zKeys2)[pp*n + i];
				}
				if(s.time)
					s.time[idx] = (f.shutterClose - f.shutterOpen) * f.val1d[(size_t)patT*n + i] + f.shutterOpen;
				if(s.dof)
				{
					// samples[shuffled[i]].dofOffset = projectToCircle(-1 + 2*dofOffsets[i])
					const int j = f.shufTab[(size_t)patS*n + i];
					const float2 d = f.posTab[(size_t)patD*n + i];
					const float vx = -1.f + 2.f*d.x, vy = -1.f + 2.f*d.y;
					float m2 = (vy == 0.f) ? vx*vx : ((vx == 0.f) ? vy*vy : vx*vx + vy*vy);
					float r = sqrtf(m2);
					float2 o2 = make_float2(0.f, 0.f);
					if(r != 0.f)
					{
						float adj = maxA(fabsf(vx), fabsf(vy)) / r;
						o2 = make_float2(adj*vx, adj*vy);
					}
					s.dof[base + s.subOfs[j]] = o2;
				}
			}
			if(lane == 0 && s.dof) s.shufPat[pix] = (uint8_t)patS;
		}
		for(int i = tid; i < f.tileW*f.tileH; i += THREADS) s.pixZ[i] = 0xffffffffu;
		if(tid == 0) { s_tileZ = 0xffffffffu; s_dirty = 0; s_lastRef = 0; }
		PHASE_BARRIER(0);
		if((PLAIN ? (unsigned long long*)nullptr : f.zKeys))
		{
			if(warp == 0) refreshPixZ(f, t, s, lane);        // start from the hierarchical z of the stored keys
			PHASE_BARRIER(0);
		}
		const uint32_t binBeg = f.binOffset[slot], binCnt = f.binOffset[slot+1] - binBeg;
		// Static frames: the opaque micropolygons come first in a (single-run) bin and each pass walks its own part
		// (tileFlags = where the transparent part starts).  Motion blur / depth of field frames keep whole bins in both
		// passes (tileFlags bit 0 = the bin has micropolygons for the transparent pass): any extra state in that kernel's
		// main loop costs more in instruction fetch than the skipped entries would save.
		const uint32_t tflags = f.tileFlags[slot];
		const uint32_t split = MBDOF ? 0xffffffffu : (f.anyTransparent ? tflags : binCnt);
		const bool parted = !MBDOF && split != 0xffffffffu;
		const bool hasDeep = MBDOF ? (!(!PLAIN && f.zOnly) && f.anyTransparent && (tflags & 1u))
		                           : (f.anyTransparent && !(!PLAIN && f.zOnly) && (split == 0xffffffffu || split < binCnt));
		// ---- Render_MPGs: opaque pass, then (if the tile saw non-opaque micropolygons) deep pass
		// MBDOF kernel: keep ONE copy of the (large) pass body in the instruction stream; the compiler would
		// otherwise peel the loop.  The static kernel is small enough to profit from the specialised copies.
#pragma unroll (MBDOF ? 1 : 2)
		for(int pass = 0; pass < 2; ++pass)
		{
			if(MBDOF) asm volatile("" : "+r"(pass));
			if(pass == 1)
			{
				if(!hasDeep) break;      // (occlusion only needs the opaque pass)
				PHASE_BARRIER(1);
				if(tid == 0) { s_next = parted ? split : 0u; s_dirty = 0; s_lastRef = parted ? split : 0u; }
				if(warp == 0) refreshPixZ(f, t, s, lane);        // the opaque depths are final now
				PHASE_BARRIER(5);
			}
			const uint32_t passBeg = (pass == 1 && parted) ? split : 0u, passEnd = (pass == 0 && parted) ? split : binCnt;
			// Motion blur / depth of field.  The general kernel is bound by instruction fetch: free-running warps are spread all over
			// its per-micropolygon code (57.7 % of the stall samples were `no_inst`, profiles/r02b_k_hide_config3_scale0.5.txt), so
			// its warps take their micropolygons IN LOCK STEP -- one CTA-wide barrier per micropolygon, -22 % -- and run the set-up
			// and the first enumeration rounds on the same cache lines; barriers further inside (per enumeration round, before the
			// final drain) cost more in waiting than they save in fetch.  The PLAIN kernel is short enough to stay in the
			// instruction cache: there the barrier only adds waiting (38 % of its stall samples) and the warps run free
			// (config 3: general 426 ms free / 357 ms lock step, PLAIN 298 ms lock step / 238 ms free; profiles/README.md).
			if(MBDOF)
			{
				bool done = false;
				for(;;)
				{
					bool have = false;
					uint32_t p = 0;
					if(!done)
					{
						uint32_t base = 0;
						if(lane == 0) base = atomicAdd(&s_next, 1u);
						base = __shfl_sync(0xffffffffu, base, 0);
						if(base >= passEnd) done = true;
						else
						{
							// hierarchical-z refresh, paced tile-wide (see the static loop below)
							const uint32_t REFRESH_EVERY = f.tune[2] ? (uint32_t)f.tune[2] : (PLAIN ? 64u : 32u);
							const uint32_t last = *(volatile uint32_t*)&s_lastRef;
							if(base - passBeg >= (uint32_t)NWARPS && base - last >= REFRESH_EVERY && *(volatile uint32_t*)s.dirty)
							{
								uint32_t mine = 0;
								if(lane == 0) mine = (atomicCAS(&s_lastRef, last, base) == last) ? 1u : 0u;
								mine = __shfl_sync(0xffffffffu, mine, 0);
								if(mine)
								{
									if(lane == 0) *(volatile uint32_t*)s.dirty = 0;
									__syncwarp();
									refreshPixZ(f, t, s, lane);
									__syncwarp();
								}
							}
							const unsigned long long ent = f.binEntries[binBeg + base];
							// sorted by nearest depth: the first micropolygon behind the whole tile ends its sorted run
							if(!(pass == 1 && (!PLAIN && f.anyCSG)) && (uint32_t)(ent >> 32) > *(volatile uint32_t*)s.tileZ)
							{
								const uint32_t runEnd = min(passEnd, (base / (uint32_t)f.sortRun + 1u)*(uint32_t)f.sortRun);
								if(lane == 0) atomicMax(&s_next, runEnd);
							}
							else { have = true; p = (uint32_t)ent; }
						}
					}
					if(PLAIN) { if(done) break; }
					else if(__syncthreads_and(done ? 1 : 0)) break;
					if(have)
					{
						const bool handled = renderMBOrDof<PLAIN>(f, t, s, dc, ws, p, lane, pass == 0);
						if(!handled)
						{
							// static micropolygon in a frame without depth of field
							if(lane == 0) setupStaticRecCall(f, t, s.pixZ, p, pass == 0, &myRecs[0]);
							__syncwarp();
							if(pass == 0) sampleStaticRec<true, false, PLAIN>(f, t, s, dc, myRecs[0], lane);
							else sampleStaticRec<false, false, PLAIN>(f, t, s, dc, myRecs[0], lane);
							__syncwarp();
						}
					}
				}
			}
			else
			for(;;)
			{
				uint32_t base = 0;
				// Small grabs keep few micropolygons in flight, so that the front-to-back order of the
				// bin turns into culling early.  Once the first wave (one grab per warp) is under way the
				// hierarchical z is refreshed whenever new hits have landed, and because the bin is sorted
				// by nearest depth, the first micropolygon found behind the whole tile ends its sorted run.
				// a pass with few entries (the opaque part of a bin under many transparent layers) is spread over all warps
				const uint32_t GRAB = f.tune[pass] ? (uint32_t)f.tune[pass] : min(8u, max(2u, (passEnd - passBeg + NWARPS - 1u)/NWARPS));
				if(lane == 0) base = atomicAdd(&s_next, GRAB);
				base = __shfl_sync(0xffffffffu, base, 0);
				if(base >= passEnd) break;
				const int cnt = min(GRAB, passEnd - base);
				// tile-wide pacing: one refresh per REFRESH_EVERY bin entries handed out (whichever warp crosses the
				// mark takes it) -- 16 warps each refreshing on their own schedule spent 12 % of the kernel here
				const uint32_t REFRESH_EVERY = f.tune[2] ? (uint32_t)f.tune[2] : (MBDOF ? 16u : 96u);
				{
					const uint32_t last = *(volatile uint32_t*)&s_lastRef;
					if(base - passBeg >= GRAB*NWARPS && base - last >= REFRESH_EVERY && *(volatile uint32_t*)s.dirty)
					{
						uint32_t mine = 0;
						if(lane == 0) mine = (atomicCAS(&s_lastRef, last, base) == last) ? 1u : 0u;
						mine = __shfl_sync(0xffffffffu, mine, 0);
						if(mine)
						{
							if(lane == 0) *(volatile uint32_t*)s.dirty = 0;
							__syncwarp();
							refreshPixZ(f, t, s, lane);
							__syncwarp();
						}
					}
				}
				// the grab's entries in one coalesced load (one lane each); the first one also decides the early out
				const unsigned long long ent = (lane < cnt) ? f.binEntries[binBeg + base + lane] : 0ull;
				{
					// (partitioned bins carry the depth key without its last bit: a lower bound of the real one)
					const uint32_t zfirst = MBDOF ? __shfl_sync(0xffffffffu, (uint32_t)(ent >> 32), 0)
					                              : (__shfl_sync(0xffffffffu, (uint32_t)(ent >> 32), 0) << 1);
					// (CSG micropolygons are not cullable: with any in the frame the deep pass visits every entry)
					if(!(pass == 1 && (!PLAIN && f.anyCSG)) && zfirst > *(volatile uint32_t*)s.tileZ)
					{
						const uint32_t runEnd = min(passEnd, (base / (uint32_t)f.sortRun + 1u)*(uint32_t)f.sortRun);
						if(lane == 0) atomicMax(&s_next, runEnd);
						continue;
					}
				}
				{
					if(lane < cnt) setupStaticRec(f, t, s.pixZ, (uint32_t)ent, pass == 0, myRecs[lane]);
					__syncwarp();
					for(int j = 0; j < cnt; ++j)
					{
						if(pass == 0) sampleStaticRec<true, true, PLAIN>(f, t, s, dc, myRecs[j], lane);
						else sampleStaticRec<false, true, PLAIN>(f, t, s, dc, myRecs[j], lane);
					}
					__syncwarp();
				}
			}
		}
		PHASE_BARRIER(hasDeep ? 2 : 1);
		if(s.head && !(!PLAIN && f.zOnly))
		{
			// statistics: transparent hits kept by the tile's samples (padding slots hold 0)
			uint32_t mine = 0;
			for(int idx = tid; idx < s.nsP; idx += THREADS) mine += s.head[idx] & DEEP_COUNT_MASK;
			mine = __reduce_add_sync(0xffffffffu, mine);
			if(lane == 0 && mine) atomicAdd(&f.counters[2], (unsigned long long)mine);
		}
		// ---- occlusion feedback for the front end (CqOcclusionTree, occlusion.cpp:54-225, at pixel granularity)
		if((PLAIN ? (float*)nullptr : f.occlImage))
		{
			if(warp == 0) refreshPixZ(f, t, s, lane);
			PHASE_BARRIER(6);
			const int tw0 = t.rx1 - t.rx0, th0 = t.ry1 - t.ry0;
			for(int pix = tid; pix < tw0*th0; pix += THREADS)
			{
				const int ly = pix / tw0, lx = pix - ly*tw0;
				const int X = t.rx0 + lx, Y = t.ry0 + ly;
				(PLAIN ? (float*)nullptr : f.occlImage)[(size_t)(Y - f.sy0)*f.sw + (X - f.sx0)] = keyDepth(s.pixZ[ly*f.tileW + lx]);
			}
			if((!PLAIN && f.zOnly) && (PLAIN ? (unsigned long long*)nullptr : f.zKeys))
			{
				// keep the per-sample occlusion keys for the next flush / the final frame: one warp per pixel
				for(int pix = warp; pix < tw0*th0; pix += NWARPS)
				{
					const int ly = pix / tw0, lx = pix - ly*tw0;
					const size_t pp = (size_t)(t.ry0 + ly - f.sy0)*f.sw + (size_t)(t.rx0 + lx - f.sx0);
					const int base = (ly*ys)*s.stride + lx*xs;
					for(int i = lane; i < n; i += 32)
					{
						(PLAIN ? (unsigned long long*)nullptr : f.zKeys)[pp*n + i] = s.keys[base + s.subOfs[i]];
						if((!PLAIN && f.midpointZ)) (PLAIN ? (unsigned long long*)nullptr : f.zKeys2)[pp*n + i] = s.keys[s.nsP + base + s.subOfs[i]];
					}
				}
			}
			if((!PLAIN && f.zOnly)) continue;          // the next tile's first barrier orders the reads of pixZ above
		}
		// ---- Combine_samples + hand the resolved samples to the filter stage.
		const int tw = t.rx1 - t.rx0, th = t.ry1 - t.ry0;
		if(!PARTIALS)
		{
			// Planes are [k][y][chunk][x][slot].  The unit of work is 32 consecutive sample slots of one pixel, one per
			// lane; warps take units from a tile-wide counter (samples with long transparent lists cost several times
			// the others: a static pixel-per-warp split left most warps waiting for the slowest one).
			const int SC = f.planeSC, nCh = f.planeChunks, nSlots = SC*nCh;
			const int unitsPerPix = (nSlots + 31) >> 5, nUnits = tw*th*unitsPerPix;
			for(;;)
			{
				int unit = 0;
				if(lane == 0) unit = (int)atomicAdd(&s_resNext, 1u);
				unit = __shfl_sync(0xffffffffu, unit, 0);
				if(unit >= nUnits) break;
				const int pix = unit / unitsPerPix;
				const int ly = pix / tw, lx = pix - ly*tw;
				const int X = t.rx0 + lx, Y = t.ry0 + ly;
				// the planes hold ringRows sample rows at a time (a ring over the bands of the frame)
				const size_t rowAt = (size_t)((Y - f.sy0) % f.ringRows)*nCh*f.planeW + (size_t)(X - f.sx0);
				const int i = ((unit - pix*unitsPerPix) << 5) + lane;
				if(i < nSlots)
				{
					const int c = i / SC, j = i - c*SC;
					const size_t at = (rowAt + (size_t)c*f.planeW)*SC + j;
					if(i >= n) storeMask(f, at, 0u);       // padding slot: never included
					else
					{
					const int idx = sampleIdx(f, s, lx, ly, i);
					float out[7]; bool valid; uint32_t pNear;
					resolveSample<MBDOF, DFGEN, PLAIN>(f, t, s, dc, idx, out, valid, pNear);
					storeMask(f, at, packMask(f, tapMask(f, s.posx[idx], s.posy[idx], X, Y, valid)));
					if(valid)
					{
#pragma unroll
						for(int k = 0; k < 7; ++k) f.planes[(size_t)k*f.planeStride + at] = out[k];
						if((PLAIN ? 0 : f.aovFloats))
						{
							// StoreExtraData (bucketprocessor.cpp:1573-1643): the values at the index of the micropolygon of the
							// NEAREST hit, uninterpolated; planes 7.. are filtered like colour
							const GridRec g = f.grids[infoOf(f.P4[pNear]) & VINFO_GRID_MASK];
							const float* av = f.aov + ((size_t)g.vbase + (pNear - g.pbase))*(PLAIN ? 0 : f.aovFloats);
							for(int k = 0; k < (PLAIN ? 0 : f.aovFloats); ++k) f.planes[(size_t)(7 + k)*f.planeStride + at] = av[k];
						}
					}
					}
				}
			}
		}
		else
		{
			// Tile-partials mode: one warp per pixel.  32 samples at a time are resolved (one per
			// lane) into the warp's scratch; then lane (tap, group) runs over them in sample order and
			// adds weight*value into its three of the nine per-(pixel,tap) sums
			// (0 gTot, 1 hit count | R G B | Or Og Ob Z is split 3/3/3 over the groups).
			float4* scratch = reinterpret_cast<float4*>(myRecs);     // 32 samples x 2 float4
			const size_t planeSz = (size_t)f.sw*f.sh;
			for(int pix = warp; pix < tw*th; pix += NWARPS)
			{
				const int lx = pix % tw, ly = pix / tw;
				const int X = t.rx0 + lx, Y = t.ry0 + ly;
				const size_t at = (size_t)(Y - f.sy0)*f.sw + (size_t)(X - f.sx0);
				for(int c0 = 0; c0 < n; c0 += 32)
				{
					const int i = c0 + lane;
					if(i < n)
					{
						const int idx = sampleIdx(f, s, lx, ly, i);
						float out[7]; bool valid; uint32_t pNear;
						resolveSample<MBDOF, DFGEN, PLAIN>(f, t, s, dc, idx, out, valid, pNear);
						const uint32_t m = tapMask(f, s.posx[idx], s.posy[idx], X, Y, valid);
						if(!valid) { out[0] = out[1] = out[2] = out[3] = out[4] = out[5] = out[6] = 0.f; }
						scratch[2*lane] = make_float4(out[0], out[1], out[2], out[3]);
						scratch[2*lane+1] = make_float4(out[4], out[5], out[6], __uint_as_float(m));
					}
					__syncwarp();
					const int cn = min(32, n - c0);
					for(int t0 = 0; t0 < f.ntaps; t0 += 10)
					{
						const int tap = t0 + lane/3, grp = lane - (lane/3)*3;
						if(lane < 30 && tap < f.ntaps)
						{
							const int fy = tap / (2*f.shiftX+1), fx = tap - fy*(2*f.shiftX+1);
							const uint32_t need = (1u << fx) | (1u << (15 + fy));
							const float* g = f.filterTab + (size_t)tap*n + c0;
							float a0 = 0.f, a1 = 0.f, a2 = 0.f;
							for(int j = 0; j < cn; ++j)
							{
								const float4 hi = scratch[2*j+1];
								const uint32_t m = __float_as_uint(hi.w);
								if((m & need) != need) continue;
								const float w = g[j];
								if(grp == 0)
								{
									a0 += w;
									if(m & 0x80000000u) { a1 += 1.0f; a2 += scratch[2*j].x * w; }
								}
								else if(m & 0x80000000u)
								{
									const float4 lo = scratch[2*j];
									if(grp == 1) { a0 += lo.y * w; a1 += lo.z * w; a2 += lo.w * w; }
									else { a0 += hi.x * w; a1 += hi.y * w; a2 += hi.z * w; }
								}
							}
							float* dst = f.partials + ((size_t)tap*9 + grp*3)*planeSz + at;
							if(c0 == 0) { dst[0] = a0; dst[planeSz] = a1; dst[2*planeSz] = a2; }
							else { dst[0] += a0; dst[planeSz] += a1; dst[2*planeSz] += a2; }
						}
					}
					__syncwarp();
				}
			}
		}
	}
#ifdef AQH_PHASE_TIMING
	__syncthreads();
	if(tid < 14 && s_ph[tid]) atomicAdd(&f.counters[4 + tid], s_ph[tid]);
#endif
}

// Per tile: where the transparent part of its bin starts.  k_bin<fill> puts "goes to the transparent pass" in the top bit of
// every entry, so a sorted bin is [opaque micropolygons, front to back | the others, front to back] and each pass of k_hide
// only walks its own part.  A bin longer than one sorted run is only partitioned within its runs: it gets 0xffffffff
// ("both passes walk everything") unless it is all of one kind.
__global__ void __launch_bounds__(256) k_tile_flags(const __grid_constant__ DevFrame f)
{
	const int slot = blockIdx.x;
	__shared__ uint32_t s_opaque;
	if(threadIdx.x == 0) s_opaque = 0;
	__syncthreads();
	const uint32_t beg = f.binOffset[slot], end = f.binOffset[slot+1];
	if(!f.binPartition)
	{
		// motion blur / depth of field frames: bit 0 = the bin holds micropolygons for the transparent pass
		uint32_t any = 0;
		for(uint32_t e = beg + threadIdx.x; e < end && !any; e += 256)
		{
			const uint32_t p = (uint32_t)f.binEntries[e];
			const uint32_t info = infoOf(f.P4[p]);
			if(!mpOpaqueSlot(f, f.grids[info & VINFO_GRID_MASK], p, info)) any = 1;
		}
		if(any) s_opaque = 1;
		__syncthreads();
		if(threadIdx.x == 0) f.tileFlags[slot] = s_opaque;
		return;
	}
	uint32_t mine = 0;
	for(uint32_t e = beg + threadIdx.x; e < end; e += 256)
		mine += (f.binEntries[e] >> 63) ? 0u : 1u;
	mine = __reduce_add_sync(0xffffffffu, mine);
	if((threadIdx.x & 31) == 0 && mine) atomicAdd(&s_opaque, mine);
	__syncthreads();
	if(threadIdx.x == 0)
	{
		const uint32_t cnt = end - beg, nOpaque = s_opaque;
		f.tileFlags[slot] = (cnt <= (uint32_t)f.sortRun || nOpaque == 0u || nOpaque == cnt) ? nOpaque : 0xffffffffu;
	}
}

// ------------------------------------------------------------------------------------
// k_filter: one thread per output pixel, taps in the reference's fy, fx, sy, sx order so the
// float sums round identically (bucketprocessor.cpp:584-664); then coverage/alpha (:695-707),
// ExposeBucket (:766-806) and FormatBucketForDisplay (ddmanager.cpp:1046-1113).
__device__ __forceinline__ double clampD(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }

// FormatBucketForDisplay (ddmanager.cpp:1046-1113) for one pixel: px = the pixel's floats of the channel buffer.
__device__ __forceinline__ void quantisePixel(const DevFrame& f, const DevDisplays& disp, int x, int y, const float* px)
{
	for(int d = 0; d < disp.n; ++d)
	{
		const DevDisplay& dd = disp.d[d];
		const double s = (double)f.dither[((size_t)d*f.yres + y)*f.xres + x];
		unsigned char* pd = dd.out + ((size_t)y*f.xres + x)*dd.entrySize;
		for(int c = 0; c < dd.nChannels; ++c)
		{
			double value = (double)px[dd.channel[c]];
			if(dd.qOne != 0.f)
			{
				double v = (double)dd.qZero + value * (double)(dd.qOne - dd.qZero) + ((double)dd.qDither * s);
				// lround(x) = lfloor(x - 0.5) + 1, math.h:47-70
				double xm = v - 0.5;
				long long li = (long long)xm;
				li = li - ((xm < 0.0 && xm != (double)li) ? 1 : 0);
				value = (double)(li + 1);
				value = clampD(value, (double)dd.qMin, (double)dd.qMax);
			}
			switch(dd.type)
			{
				case AQH_FLOAT32: { float v = (float)value; memcpy(pd, &v, 4); pd += 4; break; }
				case AQH_UNSIGNED32: { value = clampD(value, 0.0, 4294967295.0); uint32_t v = (uint32_t)value; memcpy(pd, &v, 4); pd += 4; break; }
				case AQH_SIGNED32: { value = clampD(value, -2147483648.0, 2147483647.0); int32_t v = (int32_t)value; memcpy(pd, &v, 4); pd += 4; break; }
				case AQH_UNSIGNED16: { uint16_t v = (uint16_t)(int)value; memcpy(pd, &v, 2); pd += 2; break; }
				case AQH_SIGNED16: { int16_t v = (int16_t)(int)value; memcpy(pd, &v, 2); pd += 2; break; }
				case AQH_UNSIGNED8: { *pd++ = (unsigned char)(int)value; break; }
				case AQH_SIGNED8: { *pd++ = (unsigned char)(signed char)(int)value; break; }
			}
		}
	}
}

// ExposeBucket (bucketprocessor.cpp:766-806) on Ci of one pixel.
__device__ __forceinline__ void exposePixel(const DevFrame& f, float* out)
{
	if(f.expGain == 1.0f && f.expGamma == 1.0f) return;
	const float oneovergamma = 1.0f / f.expGamma;
#pragma unroll
	for(int k = 0; k < 3; ++k)
	{
		if(f.expGain != 1.0f) out[k] *= f.expGain;
		if(f.expGamma != 1.0f) out[k] = (float)pow((double)out[k], (double)oneovergamma);
	}
}

// Everything after the sums: normalise, coverage/alpha (bucketprocessor.cpp:633-707),
// ExposeBucket (:766-806), store the 9 standard floats of the pixel and quantise per display
// (FormatBucketForDisplay, ddmanager.cpp:1046-1113).  With f.deferDisplay the quantisation (and with f.deferExpose the
// exposure) is left to k_finish: the displays may show arbitrary output variables, an imager may still change the pixel.
__device__ __forceinline__ void finishPixel(const DevFrame& f, const DevDisplays& disp, int x, int y,
                                            const float acc[7], float gTot, int SampleCount)
{
	const int n = f.n;
	float out[9];
	float coverage;
	if(SampleCount == 0)
	{
#pragma unroll
		for(int k = 0; k < 9; ++k) out[k] = 0.f;
		out[AQH_CH_Z] = FLT_MAX;
		coverage = 0.f;
	}
	else
	{
		const float oneOverGTot = 1.0f / gTot;
#pragma unroll
		for(int k = 0; k < 6; ++k) out[k] = acc[k] * oneOverGTot;
		out[AQH_CH_Z] = acc[6] * oneOverGTot;
		coverage = (SampleCount >= n) ? 1.0f : (float)SampleCount / (float)n;
	}
	const float a = (out[3] + out[4] + out[5]) / 3.0f;
	out[AQH_CH_ALPHA] = a * coverage;
	out[AQH_CH_COVERAGE] = coverage;
	if(!f.deferExpose) exposePixel(f, out);
	// A NaN (0/0 of the "average" depth filter over no qualifying hit, inf - inf of filtered FLT_MAX depths with
	// negative lobes) is the x86 default NaN 0xFFC00000 in the reference's build; the GPU's canonical NaN is
	// 0x7FFFFFFF.  Same value class, made the same bits.
#pragma unroll
	for(int k = 0; k < 9; ++k) if(out[k] != out[k]) out[k] = __uint_as_float(0xffc00000u);
	float* dst = f.channels + ((size_t)y*f.xres + x)*f.nch;
#pragma unroll
	for(int k = 0; k < 9; ++k) dst[k] = out[k];
	if(!f.deferDisplay) quantisePixel(f, disp, x, y, out);
}

// The filtered floats of an arbitrary-output-variable pass: values kBase-7 .. of the pixel's AOV block
// (bucketprocessor.cpp:633-653: zero when no sample of the footprint holds a hit, else sum / gTot).
__device__ __forceinline__ void finishAovPixel(const DevFrame& f, int x, int y, const float* acc, int nVal, int kBase, float gTot, int SampleCount)
{
	float* dst = f.channels + ((size_t)y*f.xres + x)*f.nch + 9 + (kBase - 7);
	const float oneOverGTot = 1.0f / gTot;
	for(int k = 0; k < nVal; ++k)
	{
		float v = (SampleCount == 0) ? 0.f : acc[k] * oneOverGTot;
		if(v != v) v = __uint_as_float(0xffc00000u);
		dst[k] = v;
	}
}

// The deferred tail of the frame: exposure (when an imager ran in between) and the quantisation of every display, one
// thread per pixel of the crop window.
__global__ void __launch_bounds__(256) k_finish(const __grid_constant__ DevFrame f, const __grid_constant__ DevDisplays disp, int expose)
{
	const int x = f.cropX0 + blockIdx.x*blockDim.x + threadIdx.x;
	const int y = f.cropY0 + blockIdx.y*blockDim.y + threadIdx.y;
	if(x >= f.cropX1 || y >= f.cropY1) return;
	if(f.rowOwned && !f.rowOwned[y]) return;
	float* px = f.channels + ((size_t)y*f.xres + x)*f.nch;
	if(expose)
	{
		float ci[3] = {px[0], px[1], px[2]};
		exposePixel(f, ci);
#pragma unroll
		for(int k = 0; k < 3; ++k) { if(ci[k] != ci[k]) ci[k] = __uint_as_float(0xffc00000u); px[k] = ci[k]; }
	}
	quantisePixel(f, disp, x, y, px);
}

// Reference-order filter: one running sum per output pixel over the per-sample planes.
__global__ void __launch_bounds__(256) k_filter(const __grid_constant__ DevFrame f, const __grid_constant__ DevDisplays disp, int weightsInSmem, int yBeg, int yEnd, int kBase, int nVal)
{
	extern __shared__ float s_filtBuf[];
	const int taps = (2*f.shiftX+1)*(2*f.shiftY+1)*f.n;
	// very wide filters at many samples per pixel (15x15 taps x 256 samples = 230 KB) do not fit: read the table from HBM/L2
	const float* s_filt = weightsInSmem ? s_filtBuf : f.filterTab;
	if(weightsInSmem)
		for(int i = threadIdx.x + threadIdx.y*blockDim.x; i < taps; i += blockDim.x*blockDim.y) s_filtBuf[i] = f.filterTab[i];
	__syncthreads();
	const int x = f.cropX0 + blockIdx.x*blockDim.x + threadIdx.x;
	const int y = yBeg + blockIdx.y*blockDim.y + threadIdx.y;
	if(x >= f.cropX1 || y >= yEnd) return;
	if(f.rowOwned && !f.rowOwned[y]) return;
	const int n = f.n, xmax = f.shiftX, ymax = f.shiftY;
	float acc[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
	float gTot = 0.f;
	int SampleCount = 0;
	for(int fy = -ymax; fy <= ymax; ++fy)
		for(int fx = -xmax; fx <= xmax; ++fx)
		{
			const float* g = s_filt + ((fy + ymax)*(2*xmax+1) + fx + xmax)*n;
			const uint32_t need = (1u << (fx + xmax)) | (1u << (2*xmax + 1 + fy + ymax));
			const uint32_t validBit = 1u << (2*xmax + 2*ymax + 2);
			const int SC = f.planeSC, nCh = f.planeChunks;
			for(int sIdx = 0; sIdx < n; ++sIdx)
			{
				const int c = sIdx / SC;
				const size_t at = (((size_t)(y + fy - f.sy0)*nCh + c)*f.planeW + (size_t)(x + fx - f.sx0))*SC + (sIdx - c*SC);
				const uint32_t m = loadMask(f, at);
				if((m & need) == need)
				{
					const float w = g[sIdx];
					gTot += w;
					if(m & validBit)
					{
#pragma unroll
						for(int k = 0; k < 7; ++k)
							if(k < nVal) acc[k] += f.planes[(size_t)(kBase + k)*f.planeStride + at] * w;
						SampleCount++;
					}
				}
			}
		}
	if(kBase == 0) finishPixel(f, disp, x, y, acc, gTot, SampleCount);
	else finishAovPixel(f, x, y, acc, nVal, kBase, gTot, SampleCount);
}

// Reference-order filter, pixel spans staged by bulk copies and channel split.
// A CTA owns W consecutive output pixels of ONE image row and has 8*W threads: thread (px, ch)
// accumulates channel ch of pixel px -- ch 0..6 = R G B Or Og Ob Z, ch 7 = the weight total -- over the
// taps in exactly the reference's fy, fx, sy, sx order (bucketprocessor.cpp:597-629), so every
// per-channel float sum is the reference's own sequence of roundings.
// The resolved samples lie in HBM as [plane][row][chunk][pixel][slot] (hider_device.h), so the samples of a
// span of pixels of one source row (and one chunk of <= 64 slots) are ONE contiguous piece per plane: a
// stage is eight cp.async.bulk copies (TMA engine, completion on an mbarrier) issued by one thread.
//   n <= 64 : one stage per fy holds the W + 2*xmax pixels all fx taps need (each sample travels
//             L2 -> SM (2*ymax+1)*(W+2*xmax)/W times);
//   n  > 64 : the tile would not fit, so a stage is one (fy, fx, chunk) of exactly W pixels.
// In shared memory the planes are skewed by 16 bytes each, and the eight channel threads of a pixel sit in
// adjacent lanes: a quarter warp's 16-byte loads of four consecutive samples fall into eight different bank
// groups (values) or on one address (mask), i.e. no bank conflicts.  The weight total is "channel 7" with
// a value of 1.0 (1.0f*w == w exactly) so that all eight lanes run the same instructions.
// Weights come from constant memory (uniform index).  Excluded samples are skipped by predication,
// never multiplied by zero.
__constant__ __align__(16) float c_filt[49*256 + 64];     // [tap][chunk][planeSC] (slots past n hold 0)
// per (tap, chunk): which groups of four consecutive sample slots can hold a sample inside the tap's filter support at all
// (launchFilter derives it from the sub-pixel cell every slot's sample lies in); an empty chunk is not even copied
__constant__ uint16_t c_tapGroups[49*4];

__device__ __forceinline__ uint32_t smemAddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbarInit(uint64_t* bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smemAddr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbarExpectTx(uint64_t* bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smemAddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbarWait(uint64_t* bar, uint32_t parity)
{
	asm volatile(
		"{\n"
		".reg .pred p;\n"
		"WAIT_%=:\n"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
		"@p bra DONE_%=;\n"
		"bra WAIT_%=;\n"
		"DONE_%=:\n"
		"}\n" :: "r"(smemAddr(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulkLoad(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
	             :: "r"(smemAddr(dst)), "l"(src), "r"(bytes), "r"(smemAddr(bar)) : "memory");
}

// ld.shared.v4: the compiler splits a float4 load from a runtime-strided shared address into two 8-byte
// loads, which puts two pixels in one wavefront and makes the skewed planes collide.
__device__ __forceinline__ float4 lds128(const float* p)
{
	float4 v;
	asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smemAddr(p)));
	return v;
}
#define FILTER_W 32
#define FILTER_ONES 128       /* floats of 1.0 read by channel 7 ("the weight total is the sum of 1.0*w") */
__host__ __device__ __forceinline__ int filterSpan(const DevFrame& f)      // staged pixels per stage
{
	return (f.planeChunks == 1) ? FILTER_W + ((2*f.shiftX + 3) & ~3) : FILTER_W;
}
__host__ __device__ __forceinline__ int filterPlaneFloats(const DevFrame& f)
{
	return ((filterSpan(f)*f.planeSC + 31) & ~31) + 4;     // + 4: the 16-byte skew between planes
}

// MB = bytes per mask word (1, 2 or 4).  Four consecutive slots are tested per step:
// one 32-bit shared load for byte masks, one 64-bit load for 16-bit masks, one 128-bit load for 32-bit masks.
// CHUNKED = more than 64 samples per pixel: a stage is one (fy, fx, 64-slot chunk) -- the reference sums tap by tap, all samples of
// a pixel inside each tap, so the chunks of one fx cannot be shared with the next (measured: sharing them would take 20 % off the
// filter of config 4, and breaks bit parity); chunks and groups of four slots that cannot lie inside the tap's support are skipped
// (c_tapGroups).  With one chunk per pixel the skip test would cost more than it saves.
template<int MB, bool CHUNKED>
__global__ void __launch_bounds__(8*FILTER_W) k_filter_spans(const __grid_constant__ DevFrame f, const __grid_constant__ DevDisplays disp, int yBeg, int kBase, int nVal)
{
	constexpr int W = FILTER_W;
	extern __shared__ __align__(128) unsigned char fsm[];
	__shared__ __align__(8) uint64_t s_bar;
	const int n = f.n, xmax = f.shiftX, ymax = f.shiftY;
	const int SC = f.planeSC, nCh = f.planeChunks;
	const bool halo = !CHUNKED;
	const int span = filterSpan(f);
	const int planeS = filterPlaneFloats(f);
	float* tile = reinterpret_cast<float*>(fsm);         // [7 value planes][span][SC], planes skewed; ones; mask words
	const int tid = threadIdx.x, ch = tid & 7, px = tid >> 3;
	const int x0 = f.cropX0 + blockIdx.x*W, y = yBeg + blockIdx.y;
	if(f.rowOwned && !f.rowOwned[y]) return;             // uniform over the CTA
	// kBase / nVal: the planes this pass filters -- 0 / 7 for R G B Or Og Ob Z, then groups of up to seven AOV floats;
	// channel threads beyond nVal idle, channel 7 always sums the weights
	const bool pixLive = (x0 + px) < f.cropX1;
	const bool live = pixLive && (ch < nVal || ch == 7);
	// channel 7 reads 1.0f from a region laid out so that its banks continue the skew of the seven planes
	float* ones = tile + (size_t)7*planeS;
	const unsigned char* mbase = reinterpret_cast<const unsigned char*>(ones + FILTER_ONES);
	if(tid == 0)
	{
		mbarInit(&s_bar, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	if(tid < FILTER_ONES) ones[tid] = 1.0f;
	float acc = 0.f;
	int count = 0;
	uint32_t phase = 0;
	const uint32_t vbytes = (uint32_t)(span*SC*4), mbytes = (uint32_t)(span*SC*MB);
	const int nStageFx = halo ? 1 : (2*xmax + 1);
	const uint32_t validBit = (ch < 7) ? (1u << (2*xmax + 2*ymax + 2)) : 0u;
	for(int fy = 0; fy <= 2*ymax; ++fy)
	{
		const size_t srow = (size_t)((y + fy - ymax - f.sy0) % f.ringRows);
		for(int sfx = 0; sfx < nStageFx; ++sfx)
			for(int c = 0; c < nCh; ++c)
			{
				if(CHUNKED && c_tapGroups[(fy*(2*xmax + 1) + sfx)*4 + c] == 0) continue;      // nothing of this chunk can be inside the tap (CTA-uniform)
				__syncthreads();                         // barrier initialised / everyone is done with the previous stage
				if(tid == 0)
				{
					// order the CTA's generic-proxy reads of the tile before the async-proxy overwrite
					asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
					mbarExpectTx(&s_bar, (uint32_t)nVal*vbytes + mbytes);
					// first staged pixel column of the sample region: x0 - xmax - sx0 (+ fx when every fx is staged on its own)
					const size_t col = (size_t)(x0 - f.cropX0) + (size_t)sfx;
					const size_t at = ((srow*nCh + c)*f.planeW + col)*SC;
					for(int k = 0; k < nVal; ++k)
						bulkLoad(tile + (size_t)k*planeS, f.planes + (size_t)(kBase + k)*f.planeStride + at, vbytes, &s_bar);
					bulkLoad(const_cast<unsigned char*>(mbase), f.maskPlane + at*MB, mbytes, &s_bar);
				}
				mbarWait(&s_bar, phase);
				phase ^= 1u;
				if(!pixLive) continue;
				const int fx0 = halo ? 0 : sfx, fx1 = halo ? 2*xmax : sfx;
				for(int fx = fx0; fx <= fx1; ++fx)
				{
					const int tap = fy*(2*xmax + 1) + fx;
					const uint32_t tapBits = (1u << fx) | (1u << (2*xmax + 1 + fy));
					const uint32_t need = tapBits | validBit;
					// four weights per constant-bank load (the table is padded to whole groups of four slots, launchFilter)
					const float4* w4 = reinterpret_cast<const float4*>(c_filt) + ((tap*nCh + c)*SC >> 2);
					const uint32_t groups = c_tapGroups[tap*4 + c];
					const int o = (halo ? px + fx : px)*SC;
					const float* vp = (ch < 7) ? tile + (size_t)ch*planeS + o : ones + (o & 31);
					// SampleCount (the valid samples of the footprint): the eight channel threads of a pixel each count an eighth
					// of the mask words by bit arithmetic, instead of one add per sample in every thread
					const uint32_t needV = tapBits | (1u << (2*xmax + 2*ymax + 2));
					if(MB == 1)
					{
						const uint32_t* mp = reinterpret_cast<const uint32_t*>(mbase + o);
						const uint32_t nv4 = needV*0x01010101u;
						for(int wI = ch; wI < SC/4; wI += 8)
						{
							const uint32_t t = mp[wI] & nv4;
							count += __popc(~(((t & 0x7f7f7f7fu) + 0x7f7f7f7fu) | t | 0x7f7f7f7fu));       // zero bytes
						}
						if(!live) continue;
						const uint32_t n0 = need, n1 = need << 8, n2 = need << 16, n3 = need << 24;
#pragma unroll 4
						for(int s4 = 0; s4 < SC/4; ++s4)
						{
							if(CHUNKED && !((groups >> s4) & 1u)) continue;
							const uint32_t m = mp[s4];
							const float4 v = lds128(vp + 4*s4);
							const float4 w = w4[s4];
							if((m & n0) == 0u) acc += v.x * w.x;
							if((m & n1) == 0u) acc += v.y * w.y;
							if((m & n2) == 0u) acc += v.z * w.z;
							if((m & n3) == 0u) acc += v.w * w.w;
						}
					}
					else if(MB == 2)
					{
						const uint2* mp = reinterpret_cast<const uint2*>(mbase + (size_t)o*2);
						const uint32_t nv2 = needV*0x00010001u;
						for(int wI = ch; wI < SC/4; wI += 8)
						{
							const uint2 m = mp[wI];
							const uint32_t t0 = m.x & nv2, t1 = m.y & nv2;
							count += __popc(~(((t0 & 0x7fff7fffu) + 0x7fff7fffu) | t0 | 0x7fff7fffu))      // zero halves
							       + __popc(~(((t1 & 0x7fff7fffu) + 0x7fff7fffu) | t1 | 0x7fff7fffu));
						}
						if(!live) continue;
						const uint32_t n0 = need, n1 = need << 16;
#pragma unroll 4
						for(int s4 = 0; s4 < SC/4; ++s4)
						{
							if(CHUNKED && !((groups >> s4) & 1u)) continue;
							const uint2 m = mp[s4];
							const float4 v = lds128(vp + 4*s4);
							const float4 w = w4[s4];
							if((m.x & n0) == 0u) acc += v.x * w.x;
							if((m.x & n1) == 0u) acc += v.y * w.y;
							if((m.y & n0) == 0u) acc += v.z * w.z;
							if((m.y & n1) == 0u) acc += v.w * w.w;
						}
					}
					else
					{
						const uint4* mp = reinterpret_cast<const uint4*>(mbase + (size_t)o*4);
						for(int wI = ch; wI < SC/4; wI += 8)
						{
							const uint4 m = mp[wI];
							count += ((m.x & needV) == 0u) + ((m.y & needV) == 0u) + ((m.z & needV) == 0u) + ((m.w & needV) == 0u);
						}
						if(!live) continue;
#pragma unroll 4
						for(int s4 = 0; s4 < SC/4; ++s4)
						{
							if(CHUNKED && !((groups >> s4) & 1u)) continue;
							const uint4 m = mp[s4];
							const float4 v = lds128(vp + 4*s4);
							const float4 w = w4[s4];
							if((m.x & need) == 0u) acc += v.x * w.x;
							if((m.y & need) == 0u) acc += v.y * w.y;
							if((m.z & need) == 0u) acc += v.z * w.z;
							if((m.w & need) == 0u) acc += v.w * w.w;
						}
					}
				}
			}
	}
	// gather the eight sums of a pixel; the first warp finishes the W pixels (one lane each)
	__syncthreads();
	float* sums = reinterpret_cast<float*>(fsm);         // [9][W]
	sums[ch*W + px] = acc;
	// the eight channel threads of a pixel (adjacent lanes) each counted an eighth of the footprint's valid samples
	count += __shfl_xor_sync(0xffffffffu, count, 1);
	count += __shfl_xor_sync(0xffffffffu, count, 2);
	count += __shfl_xor_sync(0xffffffffu, count, 4);
	if(ch == 0) sums[8*W + px] = __int_as_float(count);
	__syncthreads();
	if(tid < W && (x0 + tid) < f.cropX1)
	{
		float a[7];
#pragma unroll
		for(int k = 0; k < 7; ++k) a[k] = sums[k*W + tid];
		if(kBase == 0) finishPixel(f, disp, x0 + tid, y, a, sums[7*W + tid], __float_as_int(sums[8*W + tid]));
		else finishAovPixel(f, x0 + tid, y, a, nVal, kBase, sums[7*W + tid], __float_as_int(sums[8*W + tid]));
	}
}

static size_t filterSpansSmem(const DevFrame& f)
{
	const size_t tile = ((size_t)7*filterPlaneFloats(f) + FILTER_ONES)*4 + (size_t)filterSpan(f)*f.planeSC*f.maskBytes;
	const size_t sums = (size_t)9*FILTER_W*4;
	return ((tile > sums ? tile : sums) + 15) & ~(size_t)15;
}

// Tile-partials filter: sum the nine per-(pixel,tap) partial sums over the taps in fy, fx order.
__global__ void __launch_bounds__(256) k_filter_partials(const __grid_constant__ DevFrame f, const __grid_constant__ DevDisplays disp)
{
	const int x = f.cropX0 + blockIdx.x*blockDim.x + threadIdx.x;
	const int y = f.cropY0 + blockIdx.y*blockDim.y + threadIdx.y;
	if(x >= f.cropX1 || y >= f.cropY1) return;
	if(f.rowOwned && !f.rowOwned[y]) return;
	const int xmax = f.shiftX, ymax = f.shiftY;
	const size_t planeSz = (size_t)f.sw*f.sh;
	float acc[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
	float gTot = 0.f, cnt = 0.f;
	int tap = 0;
	for(int fy = -ymax; fy <= ymax; ++fy)
		for(int fx = -xmax; fx <= xmax; ++fx, ++tap)
		{
			const float* src = f.partials + (size_t)tap*9*planeSz + (size_t)(y + fy - f.sy0)*f.sw + (size_t)(x + fx - f.sx0);
			gTot += src[0];
			cnt += src[planeSz];
#pragma unroll
			for(int k = 0; k < 7; ++k) acc[k] += src[(size_t)(2 + k)*planeSz];
		}
	finishPixel(f, disp, x, y, acc, gTot, (int)cnt);
}

// ------------------------------------------------------------------------------------
// launchers
// Threads per CTA of the static kernel: 512 (two CTAs per SM) for tiles of 4096 samples, 256 (four CTAs per SM) for
// tiles of at most 2048 samples (chooseTile, hider_api.cpp).
#ifndef AQH_MBP_THREADS
#define AQH_MBP_THREADS 256      /* threads per CTA of the short motion kernel */
#endif
static inline bool smallStaticTile(const DevFrame& f) { return f.tileW*f.tileH*f.n <= 2048; }
// Project + count the bin entries of the positions [pA, pB) (whole grids): the two steps that only need
// the grids uploaded so far, so that they can run while the next chunk of the frame is still on the bus.
cudaError_t launchProjectCount(const DevFrame& f, int64_t pA, int64_t pB, cudaStream_t st)
{
	if(pB <= pA) return cudaSuccess;
	const unsigned blocks = (unsigned)((pB - pA + 255) / 256);
	k_project<<<blocks, 256, 0, st>>>(f, pA, pB);
	k_bin<false><<<blocks, 256, 0, st>>>(f, pA, pB);
	return cudaGetLastError();
}
cudaError_t launchSplitLines(const DevFrame& f, cudaStream_t st)
{
	if(f.nGrids == 0) return cudaSuccess;
	k_splitlines<<<(f.nGrids + 127)/128, 128, 0, st>>>(f);
	return cudaGetLastError();
}
cudaError_t launchBinScan(const DevFrame& f, cudaStream_t st)
{
	k_bin_scan<<<1, 1024, 0, st>>>(f);
	return cudaGetLastError();
}
__global__ void __launch_bounds__(256) k_fill_keys(unsigned long long* keys, size_t n)
{
	const size_t i = (size_t)blockIdx.x*256 + threadIdx.x;
	if(i < n) keys[i] = KEY_EMPTY;
}
cudaError_t launchFillKeys(unsigned long long* keys, size_t n, cudaStream_t st)
{
	if(n) k_fill_keys<<<(unsigned)((n + 255)/256), 256, 0, st>>>(keys, n);
	return cudaGetLastError();
}
cudaError_t launchProject(const DevFrame& f, int64_t pA, int64_t pB, cudaStream_t st)
{
	if(pB <= pA) return cudaSuccess;
	k_project<<<(unsigned)((pB - pA + 255)/256), 256, 0, st>>>(f, pA, pB);
	return cudaGetLastError();
}
cudaError_t launchBinCount(const DevFrame& f, int64_t pA, int64_t pB, cudaStream_t st)
{
	if(pB <= pA) return cudaSuccess;
	k_bin<false><<<(unsigned)((pB - pA + 255)/256), 256, 0, st>>>(f, pA, pB);
	return cudaGetLastError();
}
cudaError_t launchBinFill(const DevFrame& f, int64_t pA, int64_t pB, cudaStream_t st)
{
	if(pB > pA)
		k_bin<true><<<(unsigned)((pB - pA + 255)/256), 256, 0, st>>>(f, pA, pB);
	if(f.nActiveTiles)
	{
		// function attributes are per device: set on every launch (a process may drive hiders on several GPUs)
		cudaError_t e = cudaFuncSetAttribute(k_bin_sort, cudaFuncAttributeMaxDynamicSharedMemorySize, SORT_MAX*8);
		if(e != cudaSuccess) return e;
		if(f.sortRun > SORT_MAX) return cudaErrorInvalidValue;
		if(pB > pA) k_bin_sort<<<f.nActiveTiles, 256, (size_t)f.sortRun*8, st>>>(f, f.sortRun);
		// only the deep pass reads the flags: frames without a non-opaque vertex (and with cullable hits) skip the scan
		if(f.anyTransparent) k_tile_flags<<<f.nActiveTiles, 256, 0, st>>>(f);
	}
	return cudaGetLastError();
}

template<bool MBDOF, int THREADS, bool PARTIALS, bool DFGEN, bool PLAIN = false>
static cudaError_t configHide(const DevFrame& f, int smCount, LaunchCfg& cfg)
{
	cfg.hideThreads = THREADS;
	cfg.batchMPs = (THREADS/32)*RECS_PER_WARP;
	cfg.hideSmemBytes = hideSmemBytes(f, cfg.batchMPs, MBDOF ? (THREADS/32)*sizeof(MovScratch) : 0);
	if(cfg.hideSmemBytes > 227*1024) return cudaErrorInvalidValue;
	cudaError_t e = cudaFuncSetAttribute(k_hide<MBDOF, THREADS, PARTIALS, DFGEN, PLAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.hideSmemBytes);
	if(e != cudaSuccess) return e;
	e = cudaFuncSetAttribute(k_hide<MBDOF, THREADS, PARTIALS, DFGEN, PLAIN>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
	if(e != cudaSuccess) return e;
	int perSm = 0;
	e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k_hide<MBDOF, THREADS, PARTIALS, DFGEN, PLAIN>, THREADS, cfg.hideSmemBytes);
	if(e != cudaSuccess) return e;
	if(perSm < 1) perSm = 1;
	cfg.hideCtas = smCount * perSm;
	return cudaSuccess;
}

cudaError_t hideKernelConfig(const DevFrame& f, int smCount, LaunchCfg& cfg)
{
	cfg.smCount = smCount;
	const bool mbdof = f.useDof || f.anyMotion;
	const bool partials = f.filterMode != AQH_FILTER_REFERENCE_ORDER;
	const bool dfgen = f.depthFilter != AQH_DEPTHFILTER_MIN;
#define AQH_CFG(MB, TH) (partials ? (dfgen ? configHide<MB, TH, true, true>(f, smCount, cfg) : configHide<MB, TH, true, false>(f, smCount, cfg)) \
                                  : (dfgen ? configHide<MB, TH, false, true>(f, smCount, cfg) : configHide<MB, TH, false, false>(f, smCount, cfg)))
	if(f.plain && !f.tune[4] && !partials && !dfgen)
	{
		if(mbdof) return configHide<true, AQH_MBP_THREADS, false, false, true>(f, smCount, cfg);
		return smallStaticTile(f) ? configHide<false, 256, false, false, true>(f, smCount, cfg) : configHide<false, 512, false, false, true>(f, smCount, cfg);
	}
	return mbdof ? AQH_CFG(true, 256) : (smallStaticTile(f) ? AQH_CFG(false, 256) : AQH_CFG(false, 512));
#undef AQH_CFG
}

cudaError_t launchHide(const DevFrame& f, const LaunchCfg& cfg, uint32_t slotBeg, uint32_t slotEnd, uint32_t* cursor, cudaStream_t st)
{
	if(slotEnd <= slotBeg) return cudaSuccess;
	const bool mbdof = f.useDof || f.anyMotion;
	const bool partials = f.filterMode != AQH_FILTER_REFERENCE_ORDER;
	const bool dfgen = f.depthFilter != AQH_DEPTHFILTER_MIN;
	const int ctas = (int)std::min<uint32_t>((uint32_t)cfg.hideCtas, slotEnd - slotBeg);
#define AQH_LAUNCH(MB, TH, PA, DF) k_hide<MB, TH, PA, DF><<<ctas, TH, cfg.hideSmemBytes, st>>>(f, slotBeg, slotEnd, cursor)
#define AQH_LAUNCH2(MB, TH) do { if(partials) { if(dfgen) AQH_LAUNCH(MB, TH, true, true); else AQH_LAUNCH(MB, TH, true, false); } \
                                 else { if(dfgen) AQH_LAUNCH(MB, TH, false, true); else AQH_LAUNCH(MB, TH, false, false); } } while(0)
	if(f.plain && !f.tune[4] && !partials && !dfgen)
	{
		if(mbdof) k_hide<true, AQH_MBP_THREADS, false, false, true><<<ctas, AQH_MBP_THREADS, cfg.hideSmemBytes, st>>>(f, slotBeg, slotEnd, cursor);
		else if(smallStaticTile(f)) k_hide<false, 256, false, false, true><<<ctas, 256, cfg.hideSmemBytes, st>>>(f, slotBeg, slotEnd, cursor);
		else k_hide<false, 512, false, false, true><<<ctas, 512, cfg.hideSmemBytes, st>>>(f, slotBeg, slotEnd, cursor);
	}
	else if(mbdof) AQH_LAUNCH2(true, 256); else if(smallStaticTile(f)) AQH_LAUNCH2(false, 256); else AQH_LAUNCH2(false, 512);
#undef AQH_LAUNCH2
#undef AQH_LAUNCH
	return cudaGetLastError();
}

// Filter + expose + quantise the output rows [yBeg, yEnd) (all of their footprint rows are in the sample planes).
// One pass for the seven standard planes, then one per group of up to seven AOV floats.
// uploadTable: the first launch of a frame loads the weight table into constant memory.
int filterLaunchCount(const DevFrame& f)
{
	return (f.filterMode != AQH_FILTER_REFERENCE_ORDER) ? 1 : 1 + (f.aovFloats + 6)/7;
}
cudaError_t launchFilter(const DevFrame& f, const DevDisplays& disp, const float* hostFilterTab, int yBeg, int yEnd, bool uploadTable, cudaStream_t st)
{
	const int w = f.cropX1 - f.cropX0, h = yEnd - yBeg;
	if(w <= 0 || h <= 0) return cudaSuccess;
	if(f.filterMode != AQH_FILTER_REFERENCE_ORDER)
	{
		dim3 block(32, 8), grid((w + 31)/32, (h + 7)/8);
		k_filter_partials<<<grid, block, 0, st>>>(f, disp);
		return cudaGetLastError();
	}
	const int ntaps = (2*f.shiftX+1)*(2*f.shiftY+1);
	const int ntapw = ntaps*f.n;
	const int nPad = f.planeChunks*f.planeSC;                       // slots per pixel of the sample planes (a multiple of four, >= n)
	const size_t spanSmem = filterSpansSmem(f);
	const bool spans = ntaps*nPad <= 49*256 && spanSmem <= 113*1024;     // tap bits fit the mask word (shift <= 7 always), weights in 64 KB of constant memory
	cudaError_t e = cudaSuccess;
	if(spans && uploadTable)
	{
		// the span filter reads its weights four at a time: [tap][slot] padded to the planes' slots per pixel
		// (pageable source: the copy has left the buffer when the call returns)
		std::vector<float> padded((size_t)ntaps*nPad, 0.f);
		for(int t = 0; t < ntaps; ++t) std::memcpy(&padded[(size_t)t*nPad], hostFilterTab + (size_t)t*f.n, (size_t)f.n*4);
		e = cudaMemcpyToSymbolAsync(c_filt, padded.data(), padded.size()*4, 0, cudaMemcpyHostToDevice, st);
	}
	if(e != cudaSuccess) return e;
	if(spans && uploadTable)
	{
		// A sample of sub-pixel cell (sx, sy) lies at an offset in [sx/xs, (sx+1)/xs) x [sy/ys, (sy+1)/ys) of its pixel (jittered
		// or not); tap (fx, fy) includes it when offset + fx - 0.5 lies within +-xfwo2 (tapMask): cells that cannot satisfy
		// that -- with one cell of margin -- need not be looked at.  Only a skip list: the per-sample test stays.
		uint16_t groupsHost[49*4] = {};
		const int nx = 2*f.shiftX + 1, ny = 2*f.shiftY + 1;
		for(int tap = 0; tap < nx*ny && tap < 49; ++tap)
		{
			const int fy = tap / nx - f.shiftY, fx = tap % nx - f.shiftX;
			const double lox = 0.5 - fx - f.xfwo2, hix = 0.5 - fx + f.xfwo2, loy = 0.5 - fy - f.yfwo2, hiy = 0.5 - fy + f.yfwo2;
			const int sxA = (int)std::floor(lox*f.xs) - 1, sxB = (int)std::ceil(hix*f.xs) + 1;
			const int syA = (int)std::floor(loy*f.ys) - 1, syB = (int)std::ceil(hiy*f.ys) + 1;
			for(int c = 0; c < 4; ++c)
			{
				uint16_t g = 0;
				for(int s4 = 0; s4 < f.planeSC/4 && c < f.planeChunks; ++s4)
					for(int k = 0; k < 4; ++k)
					{
						const int i = c*f.planeSC + 4*s4 + k;
						if(i >= f.n) continue;
						const int sy = i / f.xs, sx = i % f.xs;
						if(sx >= sxA && sx <= sxB && sy >= syA && sy <= syB) g |= (uint16_t)(1u << s4);
					}
				groupsHost[tap*4 + c] = g;
			}
		}
		e = cudaMemcpyToSymbolAsync(c_tapGroups, groupsHost, sizeof groupsHost, 0, cudaMemcpyHostToDevice, st);
		if(e != cudaSuccess) return e;
	}
	for(int kBase = 0; kBase < 7 + f.aovFloats; kBase += 7)
	{
		const int nVal = kBase == 0 ? 7 : std::min(7, 7 + f.aovFloats - kBase);
		if(spans)
		{
			dim3 grid((w + FILTER_W - 1)/FILTER_W, h);
#define AQH_SPANS(MB, CH) do { e = cudaFuncSetAttribute(k_filter_spans<MB, CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)spanSmem); \
			if(e != cudaSuccess) return e; \
			k_filter_spans<MB, CH><<<grid, 8*FILTER_W, spanSmem, st>>>(f, disp, yBeg, kBase, nVal); } while(0)
			const bool chunked = f.planeChunks > 1;
			if(f.maskBytes == 1) { if(chunked) AQH_SPANS(1, true); else AQH_SPANS(1, false); }
			else if(f.maskBytes == 2) { if(chunked) AQH_SPANS(2, true); else AQH_SPANS(2, false); }
			else { if(chunked) AQH_SPANS(4, true); else AQH_SPANS(4, false); }
#undef AQH_SPANS
		}
		else
		{
			dim3 block(32, 8), grid((w + 31)/32, (h + 7)/8);
			size_t smem = (size_t)ntapw*sizeof(float);
			const int inSmem = smem <= 200*1024 ? 1 : 0;
			if(!inSmem) smem = 0;
			if(smem > 48*1024)
			{
				e = cudaFuncSetAttribute(k_filter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
				if(e != cudaSuccess) return e;
			}
			k_filter<<<grid, block, smem, st>>>(f, disp, inSmem, yBeg, yEnd, kBase, nVal);
		}
		e = cudaGetLastError();
		if(e != cudaSuccess) return e;
	}
	return cudaSuccess;
}

cudaError_t launchFinish(const DevFrame& f, const DevDisplays& disp, int expose, cudaStream_t st)
{
	const int w = f.cropX1 - f.cropX0, h = f.cropY1 - f.cropY0;
	if(w <= 0 || h <= 0) return cudaSuccess;
	dim3 block(32, 8), grid((w + 31)/32, (h + 7)/8);
	k_finish<<<grid, block, 0, st>>>(f, disp, expose);
	return cudaGetLastError();
}

int kernelsArchOk()
{
	cudaFuncAttributes a;
	return cudaFuncGetAttributes(&a, k_project) == cudaSuccess ? 1 : 0;
}

} // namespace aqh
