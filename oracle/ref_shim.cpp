// ref_shim.cpp -- TEST INFRASTRUCTURE ONLY.
//
// A thin extern "C" window onto the reference's own leaf sources, which the
// Makefile in this directory compiles IN PLACE from /root/reference (nothing is
// copied into this repository): libs/math/random.cpp, libs/core/multijitter.cpp,
// libs/core/grid.cpp, libs/core/filters.cpp, libs/core/bound.cpp and the
// header-only libs/core/bilinear.h.  The result, oracle/_ref/libaqsis_refleaf.so,
// is used by tests/ to pin (i) the oracle restatement and (ii) the product's
// host-side RNG / sampler / filter code bit-for-bit, and by
// tests/golden/make_golden.py to generate the committed golden vectors.
// Nothing in the product may load it.
#include <aqsis/aqsis.h>
#include <aqsis/math/random.h>
#include <aqsis/math/math.h>
#include <aqsis/math/vector2d.h>
#include <aqsis/math/vector3d.h>
#include <aqsis/ri/ri.h>
#include "multijitter.h"
#include "grid.h"
#include "bilinear.h"
#include "bound.h"

#include <cstring>

using namespace Aqsis;

extern "C" {

// ---- CqRandom (libs/math/random.cpp) ----
void ref_random_reseed(unsigned seed) { CqRandom r; r.Reseed(seed); }
unsigned ref_random_uint() { CqRandom r; return r.RandomInt(); }
float ref_random_float() { CqRandom r; return r.RandomFloat(); }
unsigned ref_random_int(unsigned range) { CqRandom r; return r.RandomInt(range); }

// ---- samplers (libs/core/multijitter.cpp, grid.cpp) ----
// Builds the sampler from the CURRENT global RNG state (the caller reseeds first, as
// RiWorldBegin does with 545), then performs `ndraws` rounds of the per-pixel draw
// sequence of CqImagePixel::setSamples (imagepixel.cpp:338-347) and returns, per round,
// the five table offsets (in units of patterns) that were handed out.
// It also dumps the complete tables by probing: we recover each table through the
// pointers returned by the sampler, relative to the base found by exhaustive draws.
struct RefSampler
{
	IqSampler* s;
	int n;
	bool jitter;
};

void* ref_sampler_create(int xs, int ys, int jitter)
{
	RefSampler* r = new RefSampler;
	r->n = xs*ys;
	r->jitter = jitter != 0;
	if(jitter)
		r->s = new CqMultiJitteredSampler(xs, ys);
	else
		r->s = new CqGridSampler(xs, ys);
	return r;
}
void ref_sampler_destroy(void* p)
{
	RefSampler* r = static_cast<RefSampler*>(p);
	delete r->s;
	delete r;
}
// One setSamples()-style round: copies the n entries each getter returns.
void ref_sampler_draw(void* p, int* shuffled, float* pos_xy, float* dof_xy, float* times, float* lods)
{
	RefSampler* r = static_cast<RefSampler*>(p);
	const TqInt* sh = r->s->getShuffledIndices();
	const CqVector2D* pos = r->s->get2DSamples();
	const CqVector2D* dof = r->s->get2DSamples();
	const TqFloat* t = r->s->get1DSamples();
	const TqFloat* l = r->s->get1DSamples();
	for(int i = 0; i < r->n; ++i)
	{
		shuffled[i] = sh[i];
		pos_xy[2*i] = pos[i].x(); pos_xy[2*i+1] = pos[i].y();
		dof_xy[2*i] = dof[i].x(); dof_xy[2*i+1] = dof[i].y();
		times[i] = t[i];
		lods[i] = l[i];
	}
}

// ---- pixel filters (libs/core/filters.cpp) ----
float ref_filter(int which, float x, float y, float xw, float yw)
{
	switch(which)
	{
		case 0: return RiBoxFilter(x, y, xw, yw);
		case 1: return RiTriangleFilter(x, y, xw, yw);
		case 2: return RiGaussianFilter(x, y, xw, yw);
		case 3: return RiCatmullRomFilter(x, y, xw, yw);
		case 4: return RiSincFilter(x, y, xw, yw);
		case 5: return RiMitchellFilter(x, y, xw, yw);
		case 6: return RiDiskFilter(x, y, xw, yw);
		case 7: return RiBesselFilter(x, y, xw, yw);
	}
	return 0;
}

// ---- inverse bilinear + bilerp (libs/core/bilinear.h) ----
// verts: A,B,C,D as x,y pairs; returns uv for P and whether forward bilerp of z works.
void ref_invbilinear(const float* verts, float px, float py, float* uv)
{
	CqInvBilinear inv(CqVector2D(verts[0], verts[1]), CqVector2D(verts[2], verts[3]),
	                  CqVector2D(verts[4], verts[5]), CqVector2D(verts[6], verts[7]));
	CqVector2D r = inv(CqVector2D(px, py));
	uv[0] = r.x(); uv[1] = r.y();
}
float ref_bilerp(float a, float b, float c, float d, float u, float v)
{
	return bilerp(a, b, c, d, CqVector2D(u, v));
}
void ref_bilerp2(const float* verts, float u, float v, float* out)
{
	CqVector2D r = bilerp(CqVector2D(verts[0], verts[1]), CqVector2D(verts[2], verts[3]),
	                      CqVector2D(verts[4], verts[5]), CqVector2D(verts[6], verts[7]), CqVector2D(u, v));
	out[0] = r.x(); out[1] = r.y();
}

// ---- rounding helpers (include/aqsis/math/math.h:47-70) ----
long ref_lfloor(double x) { return lfloor(x); }
long ref_lceil(double x) { return lceil(x); }
long ref_lround(double x) { return Aqsis::lround(x); }
long ref_lfloorf(float x) { return lfloor(x); }
long ref_lceilf(float x) { return lceil(x); }

// ---- CqBound (libs/core/bound.h/.cpp) ----
int ref_bound_contains2d(const float* b6, float x, float y)
{
	CqBound b(b6[0], b6[1], b6[2], b6[3], b6[4], b6[5]);
	return b.Contains2D(CqVector2D(x, y)) ? 1 : 0;
}
int ref_bound_intersects(const float* b6, float minx, float miny, float maxx, float maxy)
{
	CqBound b(b6[0], b6[1], b6[2], b6[3], b6[4], b6[5]);
	return b.Intersects(CqVector2D(minx, miny), CqVector2D(maxx, maxy)) ? 1 : 0;
}

} // extern "C"
