// ref_hider.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Drives the REFERENCE's OWN hider, compiled in place from /root/reference by oracle/Makefile
// (target `refhider`; nothing is copied into this repository, the output is
// oracle/_ref/libaqsis_refhider.so).  These reference translation units run unmodified:
//
//   libs/core/imagebuffer.cpp      CqImageBuffer::SetImage / AddMPG / RenderImage (the bucket loop)
//   libs/core/bucketprocessor.cpp  preProcess, RenderMicroPoly, RenderMPG_Static, RenderMPG_MBOrDof,
//                                  StoreSample, CombineElements, FilterBucket, ExposeBucket, overlap caches
//   libs/core/imagepixel.cpp       CqImagePixel::setSamples / Combine
//   libs/core/micropolygon.cpp     CqMicroPolygon(+Motion): Initialise, ComputeVertexOrder, bounds, BuildBoundList,
//                                  CacheHitTestValues, fContains, Sample, CacheOutputInterpCoeffs
//   libs/core/occlusion.cpp bucket.cpp bound.cpp optioncache.cpp options.cpp parameters.cpp stats.cpp
//   libs/core/multijitter.cpp grid.cpp filters.cpp csgtree.cpp threadscheduler.cpp
//   libs/math/random.cpp matrix.cpp color.cpp   libs/util/sstring.cpp logging.cpp
//
// What this file supplies around them (and therefore what is NOT the reference's code):
//   * the render context: a CqRenderer whose constructor and the few members the hider calls are
//     defined here (renderer.cpp would drag in the RI front end, shader VM and texture system);
//     every other virtual is a generated abort stub (oracle/gen_ref_stubs.py);
//     CqRenderer::MinCoCForBound restates libs/core/renderer.cpp:1602-1617;
//   * a grid class over the caller's P/Ci/Oi arrays (CqMicroPolyGridBase subclass) whose busting
//     loop restates CqMicroPolyGrid::Split (micropolygon.cpp:770-856) -- the real Split needs
//     surfaces, attributes and transforms from the geometry front end;
//   * a display manager that captures each bucket's float channel buffer and quantises it like
//     CqDisplayRequest::FormatBucketForDisplay (ddmanager.cpp:1022-1118), drawing the dither value
//     from the reference's global CqRandom in the reference's order.
//   * the pixel filter handed to CqOptions::SetfuncFilter is the REFERENCE'S OWN Ri*Filter (libs/core/filters.cpp is part
//     of this library): chosen by name (ref_set_filter) or by recognising which standard filter the caller's function
//     pointer computes; only a genuinely user-defined filter is called through the pointer;
//   * ref_can_cull: the answers of the reference's CqOcclusionTree::canCull (occlusion.cpp:161-225) for a list of
//     bounds, asked of every bucket's tree once the bucket's micropolygons are rendered.
// The product never links or loads this library; tests/ and bench.py's CPU legs do.
#include <vector>
#include <string>
#include <aqsis/aqsis.h>
#include <aqsis/math/random.h>
#include <aqsis/math/math.h>
#include <aqsis/math/region.h>
#include <aqsis/ri/ri.h>
#include <aqsis/util/file.h>
#include <aqsis/util/logging.h>
#include <aqsis/util/logging_streambufs.h>
#include <aqsis/shadervm/ishaderdata.h>
#include <aqsis/shadervm/ishaderexecenv.h>

// The trim loops of a surface are private members filled by CqTrimLoop::Prepare from NURBS curves (the front end's
// work); ref_set_trim_loops below fills the tessellated points in directly.
#define private public
#include "trimcurve.h"
#undef private
#include "renderer.h"
#include "imagebuffer.h"
// The occlusion-feedback checker (ref_can_cull below) asks the bucket processor's own CqOcclusionTree; the tree is a
// private member and the processors are locals of CqImageBuffer::RenderImage, so this translation unit -- and only
// this one -- reads the class with its access specifiers lifted (the object layout is the same).
#define private public
#include "bucketprocessor.h"
#undef private
#include "micropolygon.h"
#include "imagers.h"
#include "options.h"
#include "attributes.h"
#include "stats.h"
#include "iddmanager.h"
#include "csgtree.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <list>
#include <memory>
#include <vector>

#include "oracle_hider.h"

[[noreturn]] static void refStubAbort(const char* what)
{
	std::fprintf(stderr, "ref_hider: %s is outside the hide+filter path and was stubbed out\n", what);
	std::abort();
}

namespace Aqsis {

// ---------------------------------------------------------------------------------------
// Symbols the hider sources reference but never reach on this path.
CqRenderer* pCurrRenderer = 0;
std::list<CqAttributes*> Attribute_stack;
IqRenderer* QGetRenderContextI() { return pCurrRenderer; }
boost::shared_ptr<IqShaderExecEnv> IqShaderExecEnv::create(IqRenderer*) { refStubAbort("IqShaderExecEnv::create"); }
boostfs::path findFileNothrow(const std::string&, const std::string&) { return boostfs::path(); }
CqImagersource::CqImagersource(const boost::shared_ptr<IqShader>&, bool) { refStubAbort("CqImagersource"); }
CqImagersource::~CqImagersource() {}
void CqImagersource::Initialise(const CqRegion&, IqChannelBuffer*) { refStubAbort("CqImagersource::Initialise"); }
CqColor CqImagersource::Color(TqFloat, TqFloat) { refStubAbort("CqImagersource::Color"); }
CqColor CqImagersource::Opacity(TqFloat, TqFloat) { refStubAbort("CqImagersource::Opacity"); }
TqFloat CqImagersource::Alpha(TqFloat, TqFloat) { refStubAbort("CqImagersource::Alpha"); }

// ---------------------------------------------------------------------------------------
// The render context.
CqRenderer::CqRenderer()
	: m_pImageBuffer(0), m_pDDManager(0), m_Mode(RenderMode_Image), m_fSaveGPrims(false), m_fWorldBegin(false),
	  m_DofMultiplier(0), m_OneOverFocalDistance(FLT_MAX), m_UsingDepthOfField(false), m_DepthOfFieldScale(1, 1),
	  m_OutputDataOffset(9), m_OutputDataTotalSize(9), m_FrameNo(0), m_pErrorHandler(0), m_abortRender(false),
	  m_pProgressHandler(0), m_pRaytracer(0),
	  m_cropWindowXMin(0), m_cropWindowXMax(0), m_cropWindowYMin(0), m_cropWindowYMax(0)
{
	m_poptDefault = CqOptionsPtr(new CqOptions);
}
CqRenderer::~CqRenderer() {}
// Private state is reachable only from members, so the frame set-up rides on Initialise().
static struct RefSetup { const AqhFrameParams* p; CqImageBuffer* image; IqDDManager* dd; } g_refSetup;
void CqRenderer::Initialise()
{
	const AqhFrameParams& p = *g_refSetup.p;
	m_pImageBuffer = g_refSetup.image;
	m_pDDManager = g_refSetup.dd;
	// integer crop window as CqRenderer::initialiseCropWindow leaves it (renderer.cpp:1619-1627)
	m_cropWindowXMin = p.crop_xmin; m_cropWindowXMax = p.crop_xmax;
	m_cropWindowYMin = p.crop_ymin; m_cropWindowYMax = p.crop_ymax;
	// lens constants as SetDepthOfFieldData derives them (renderer.h:368-377); the C ABI carries them derived
	m_UsingDepthOfField = p.use_dof != 0;
	m_DofMultiplier = p.dof_multiplier;
	m_OneOverFocalDistance = p.dof_one_over_focal_distance;
	m_DepthOfFieldScale = CqVector2D(p.dof_scale_x, p.dof_scale_y);
	// what CqRenderer::RegisterOutputData leaves behind for every AOV of a RiDisplay (renderer.cpp:1520-1546; the token
	// dictionary that derives the float count from the declaration is the front end's: the ABI carries the count)
	for(int a = 0; a < p.n_aovs; ++a)
	{
		SqOutputDataEntry e;
		e.m_Offset = m_OutputDataOffset;
		e.m_NumSamples = p.aov[a].n_floats;
		m_OutputDataOffset += e.m_NumSamples;
		m_OutputDataTotalSize += e.m_NumSamples;
		m_OutputDataEntries[std::string(p.aov[a].name)] = e;
	}
}
const IqOptionsPtr CqRenderer::poptCurrent() const { return m_poptDefault; }
TqFloat CqRenderer::Time() const { return 0; }
const TqInt* CqRenderer::GetIntegerOption(const char* n, const char* p) const { return m_poptDefault->GetIntegerOption(n, p); }
const TqFloat* CqRenderer::GetFloatOption(const char* n, const char* p) const { return m_poptDefault->GetFloatOption(n, p); }
// libs/core/renderer.cpp:1602-1617 (restated: renderer.cpp is not part of this build)
const TqFloat CqRenderer::MinCoCForBound(const CqBound& bound) const
{
	TqFloat z1 = bound.vecMin().z();
	TqFloat z2 = bound.vecMax().z();
	TqFloat focalDist = 1/m_OneOverFocalDistance;
	if((z1 - focalDist)*(z2 - focalDist) < 0)
		return 0;
	TqFloat minBlur = min(std::fabs(1/z1 - m_OneOverFocalDistance), std::fabs(1/z2 - m_OneOverFocalDistance));
	return m_DofMultiplier * min(m_DepthOfFieldScale.x(), m_DepthOfFieldScale.y()) * minBlur;
}
// renderer.cpp:1549-1567 (restated: renderer.cpp is not part of this build); registration happens in Initialise() above
TqInt CqRenderer::RegisterOutputData(const char* name) { return OutputDataIndex(name); }
TqInt CqRenderer::OutputDataIndex(const char* name)
{
	std::map<std::string, SqOutputDataEntry>::const_iterator i = m_OutputDataEntries.find(name);
	return i == m_OutputDataEntries.end() ? -1 : i->second.m_Offset;
}
TqInt CqRenderer::OutputDataSamples(const char* name)
{
	std::map<std::string, SqOutputDataEntry>::const_iterator i = m_OutputDataEntries.find(name);
	return i == m_OutputDataEntries.end() ? 0 : i->second.m_NumSamples;
}
#include "renderer_stubs.inc"

} // namespace Aqsis

using namespace Aqsis;

namespace {

// ---------------------------------------------------------------------------------------
// The reference's own pixel filters (libs/core/filters.cpp), by name or by recognising the function behind a pointer.
struct RefFilter { const char* name; RtFilterFunc fn; };
const RefFilter kRefFilters[] = {
	{"box", RiBoxFilter}, {"triangle", RiTriangleFilter}, {"gaussian", RiGaussianFilter}, {"catmull-rom", RiCatmullRomFilter},
	{"sinc", RiSincFilter}, {"mitchell", RiMitchellFilter}, {"disk", RiDiskFilter}, {"bessel", RiBesselFilter},
};
RtFilterFunc g_forcedFilter = 0;

RtFilterFunc referenceFilterFor(AqhFilterFunc f)
{
	if(g_forcedFilter) return g_forcedFilter;
	if(!f) return RiGaussianFilter;
	static const float probes[5][2] = {{0.3f, 0.2f}, {-0.9f, 0.6f}, {1.3f, -1.2f}, {0.05f, -0.45f}, {-1.7f, 0.1f}};
	for(const RefFilter& rf : kRefFilters)
	{
		bool same = true;
		for(int w = 2; w <= 4 && same; ++w)
			for(int i = 0; i < 5 && same; ++i)
				same = f(probes[i][0], probes[i][1], float(w), float(w)) == rf.fn(probes[i][0], probes[i][1], float(w), float(w));
		if(same) return rf.fn;
	}
	return reinterpret_cast<RtFilterFunc>(f);       // a user filter: tabulated through the pointer, as aqsis would
}

// Occlusion queries of the current ref_can_cull call (empty during ref_render).
struct CullQueries { int n; const float* bounds; unsigned char* culledEverywhere; } g_cull = {0, 0, 0};

// ---------------------------------------------------------------------------------------
// A shader variable that is just a window onto the caller's AoS float array
// (CqVector3D and CqColor are three packed floats).
class RefShaderData : public IqShaderData
{
public:
	explicit RefShaderData(float* data) : m_data(data) {}
	virtual void GetPointPtr(const CqVector3D*& res) const { res = reinterpret_cast<const CqVector3D*>(m_data); }
	virtual void GetPointPtr(CqVector3D*& res) { res = reinterpret_cast<CqVector3D*>(m_data); }
	virtual void GetColorPtr(const CqColor*& res) const { res = reinterpret_cast<const CqColor*>(m_data); }
	virtual void GetColorPtr(CqColor*& res) { res = reinterpret_cast<CqColor*>(m_data); }
#include "shaderdata_stubs.inc"
private:
	float* m_data;
};

// One arbitrary output variable of a grid: what FindStandardVar(name) hands StoreExtraData (bucketprocessor.cpp:1573-1643).
// The caller's array is vertex-major with all AOV floats of a vertex side by side.
class RefAovData : public RefShaderData
{
public:
	RefAovData(const float* data, int stride, int offset, int nFloats)
		: RefShaderData(0), m_aov(data), m_stride(stride), m_offset(offset), m_n(nFloats) {}
	virtual EqVariableType Type() const { return m_n == 1 ? type_float : (m_n == 3 ? type_color : type_matrix); }
	virtual void GetFloat(TqFloat& res, TqInt index = 0) const { res = at(index)[0]; }
	virtual void GetColor(CqColor& res, TqInt index = 0) const { const float* v = at(index); res = CqColor(v[0], v[1], v[2]); }
	virtual void GetPoint(CqVector3D& res, TqInt index = 0) const { const float* v = at(index); res = CqVector3D(v[0], v[1], v[2]); }
	virtual void GetMatrix(CqMatrix& res, TqInt index = 0) const
	{
		TqFloat m[16];
		for(int i = 0; i < 16; ++i) m[i] = at(index)[i];
		res = CqMatrix(m);
		res.SetfIdentity(false);
	}
private:
	const float* at(TqInt index) const { return m_aov + size_t(index)*m_stride + m_offset; }
	const float* m_aov;
	int m_stride, m_offset, m_n;
};

// The surface parameters u, v of a trimmed grid (pVar(EnvVars_u) ->GetFloat(u, index), micropolygon.cpp:793-800, 1605-1618).
class RefFloatData : public RefShaderData
{
public:
	explicit RefFloatData(const float* data) : RefShaderData(0), m_f(data) {}
	virtual EqVariableType Type() const { return type_float; }
	virtual void GetFloat(TqFloat& res, TqInt index = 0) const { res = m_f[index]; }
private:
	const float* m_f;
};

// ---------------------------------------------------------------------------------------
// Trim curves.  CqMicroPolygon::Sample asks the grid's SURFACE (bCanBeTrimmed, bIsPointTrimmed) and its ATTRIBUTES
// ("trimcurve" "sense").  A real CqSurfaceNURBS needs the whole geometry front end; the three trim virtuals it
// implements are one-liners over its CqTrimLoopArray (geometry/nurbs.h:346-357).  The stand-in below is an object
// with nothing but a vtable pointer whose table carries those three entries at the slots the compiler assigned to
// CqSurface's virtuals (read off the pointers to the member functions: Itanium C++ ABI, 1 + byte offset in the
// vtable), bound to the REFERENCE'S OWN CqTrimLoopArray::TrimPoint / LineIntersects (geometry/trimcurve.cpp, compiled
// in place).
struct RefTrimSurface
{
	void** vptr;
	const CqTrimLoopArray* loops;
};
static const bool refSurfCanBeTrimmed(const RefTrimSurface*) { return true; }                      // nurbs.h:346-349
static const bool refSurfIsPointTrimmed(const RefTrimSurface* s, const CqVector2D& p) { return s->loops->TrimPoint(p); }
static const bool refSurfIsLineIntersecting(const RefTrimSurface* s, const CqVector2D& a, const CqVector2D& b) { return s->loops->LineIntersects(a, b); }
static void refSurfOther() { refStubAbort("a CqSurface virtual other than the trim queries"); }
template<class PMF> static size_t vtableSlot(PMF pmf)
{
	struct Raw { std::ptrdiff_t ptr, adj; } raw;
	static_assert(sizeof(PMF) == sizeof(Raw), "pointer to member function layout");
	std::memcpy(&raw, &pmf, sizeof raw);
	return size_t(raw.ptr - 1)/sizeof(void*);
}
static void** refSurfaceVtable()
{
	static void* table[512];
	static bool built = false;
	if(!built)
	{
		for(void*& e : table) e = reinterpret_cast<void*>(&refSurfOther);
		typedef const bool (CqSurface::*Q0)() const;
		typedef const bool (CqSurface::*Q1)(const CqVector2D&) const;
		typedef const bool (CqSurface::*Q2)(const CqVector2D&, const CqVector2D&) const;
		table[vtableSlot<Q0>(&CqSurface::bCanBeTrimmed)] = reinterpret_cast<void*>(&refSurfCanBeTrimmed);
		table[vtableSlot<Q1>(&CqSurface::bIsPointTrimmed)] = reinterpret_cast<void*>(&refSurfIsPointTrimmed);
		table[vtableSlot<Q2>(&CqSurface::bIsLineIntersecting)] = reinterpret_cast<void*>(&refSurfIsLineIntersecting);
		built = true;
	}
	return table;
}
// Attribute "trimcurve" "sense": the only attribute the hit test reads (micropolygon.cpp:1597-1601).
class RefAttributes : public IqAttributes
{
public:
	explicit RefAttributes(bool outside) : m_sense(outside ? "outside" : "inside") {}
	virtual const CqString* GetStringAttribute(const char* strName, const char* strParam) const
	{
		return (std::strcmp(strName, "trimcurve") == 0 && std::strcmp(strParam, "sense") == 0) ? &m_sense : 0;
	}
#include "attributes_stubs.inc"
private:
	CqString m_sense;
};
std::vector<CqTrimLoopArray> g_trimSets;        // ref_set_trim_loops
std::vector<RefTrimSurface> g_trimSurfaces;

// ---------------------------------------------------------------------------------------
// One shaded grid as the hider sees it.
class RefGrid : public CqMicroPolyGridBase
{
public:
	RefGrid(int cu, int cv, float* P, float* Ci, float* Oi, uint32_t flags, const float* lod)
		: m_cu(cu), m_cv(cv), m_P(P), m_Ci(Ci), m_Oi(Oi)
	{
		m_lod[0] = lod ? lod[0] : -1.0f; m_lod[1] = lod ? lod[1] : -1.0f;
		m_fTriangular = (flags & AQH_GRID_TRIANGULAR) != 0;
		// CacheGridInfo, micropolygon.cpp:46-66
		m_CurrentGridInfo.lodBounds = m_lod;
		m_CurrentGridInfo.matteFlag = (flags & AQH_GRID_MATTE_ALPHA) ? SqImageSample::Flag_MatteAlpha
		                              : ((flags & AQH_GRID_MATTE) ? SqImageSample::Flag_Matte : 0);
		m_CurrentGridInfo.usesDataMap = !(QGetRenderContext()->GetMapOfOutputDataEntries().empty());   // micropolygon.cpp:61-62
		m_CurrentGridInfo.useSmoothShading = (flags & AQH_GRID_SMOOTH) != 0;
	}
	virtual void Split(long, long, long, long) {}
	virtual void Shade(bool) {}
	virtual void TransferOutputVariables() {}
	virtual void DeleteVariables(bool) {}
	virtual CqSurface* pSurface() const { return m_surface; }
	virtual const IqConstAttributesPtr pAttributes() const { return m_attributes; }
	void setTrim(CqSurface* surface, const IqConstAttributesPtr& attributes, const float* uv, int nverts)
	{
		m_surface = surface; m_attributes = attributes;
		m_u.assign(nverts, 0.f); m_v.assign(nverts, 0.f);
		for(int i = 0; i < nverts; ++i) { m_u[i] = uv[2*i]; m_v[i] = uv[2*i+1]; }
		m_uVar.reset(new RefFloatData(m_u.data())); m_vVar.reset(new RefFloatData(m_v.data()));
	}
	virtual bool usesCSG() const { return m_csg.get() != 0; }
	virtual boost::shared_ptr<CqCSGTreeNode> pCSGNode() const { return m_csg; }
	void setCSGNode(const boost::shared_ptr<CqCSGTreeNode>& n) { m_csg = n; }
	void addAov(const std::string& name, const float* data, int stride, int offset, int nFloats)
	{
		m_aovs.push_back(std::make_pair(name, new RefAovData(data, stride, offset, nFloats)));
	}
	virtual ~RefGrid() { for(size_t i = 0; i < m_aovs.size(); ++i) delete m_aovs[i].second; }
	virtual TqInt uGridRes() const { return m_cu; }
	virtual TqInt vGridRes() const { return m_cv; }
	virtual TqUint numMicroPolygons(TqInt cu, TqInt cv) const { return cu*cv; }
	virtual TqUint numShadingPoints(TqInt cu, TqInt cv) const { return (cu+1)*(cv+1); }
	virtual bool hasValidDerivatives() const { return true; }
	virtual IqShaderData* pVar(TqInt index)
	{
		switch(index)
		{
			case EnvVars_P: return &m_P;
			case EnvVars_Ci: return m_Ci.valid() ? &m_Ci : 0;
			case EnvVars_Oi: return m_Oi.valid() ? &m_Oi : 0;
			case EnvVars_u: return m_uVar.get();
			case EnvVars_v: return m_vVar.get();
			default: return 0;
		}
	}
	virtual IqShaderData* FindStandardVar(const char* name)
	{
		for(size_t i = 0; i < m_aovs.size(); ++i) if(m_aovs[i].first == name) return m_aovs[i].second;
		return 0;
	}
	virtual boost::shared_ptr<IqShaderExecEnv> pShaderExecEnv() { return boost::shared_ptr<IqShaderExecEnv>(); }
	void addSplitLine(TqFloat time, const CqVector3D& a, const CqVector3D& b)
	{
		SqTriangleSplitLine sl;
		sl.m_TriangleSplitPoint1 = a; sl.m_TriangleSplitPoint2 = b;
		m_TriangleSplitLine.AddTimeSlot(time, sl);
	}
private:
	struct Var : RefShaderData
	{
		explicit Var(float* d) : RefShaderData(d), m_ok(d != 0) {}
		bool valid() const { return m_ok; }
		bool m_ok;
	};
	int m_cu, m_cv;
	Var m_P, m_Ci, m_Oi;
	float m_lod[2];
	boost::shared_ptr<CqCSGTreeNode> m_csg;
	std::vector<std::pair<std::string, RefAovData*> > m_aovs;
	CqSurface* m_surface = 0;
	IqConstAttributesPtr m_attributes;
	std::vector<float> m_u, m_v;
	std::unique_ptr<RefFloatData> m_uVar, m_vVar;
};

// The frame's CSG tree (ref_set_csg_tree): real CqCSGTreeNode objects, children attached in node-index order.
std::vector<boost::shared_ptr<CqCSGTreeNode> > g_csgNodes;

// ---------------------------------------------------------------------------------------
// Display manager: captures buckets.  Quantisation restates FormatBucketForDisplay
// (ddmanager.cpp:1022-1118) and draws its dither from the reference's global CqRandom.
int typeSize(int type)
{
	switch(type)
	{
		case AQH_FLOAT32: case AQH_UNSIGNED32: case AQH_SIGNED32: return 4;
		case AQH_UNSIGNED16: case AQH_SIGNED16: return 2;
		default: return 1;
	}
}

class RefDDManager : public IqDDManager
{
public:
	RefDDManager(const AqhFrameParams& p, float* channels, unsigned char* const* displays)
		: m_p(p), m_channels(channels), m_displays(displays)
	{
		for(int d = 0; d < p.n_displays; ++d)
		{
			// selectDataFormat, ddmanager.cpp:249-283
			const AqhDisplayDesc& dd = p.display[d];
			int type = dd.type;
			if(type == 0)
			{
				if(dd.quantize_one == 0) type = AQH_FLOAT32;
				else if(dd.quantize_min >= 0)
					type = dd.quantize_max <= 255.0f ? AQH_UNSIGNED8 : (dd.quantize_max <= 65535.0f ? AQH_UNSIGNED16 : AQH_UNSIGNED32);
				else if(dd.quantize_min >= -128.0f && dd.quantize_max <= 127.0f) type = AQH_SIGNED8;
				else if(dd.quantize_min >= -32768.0f && dd.quantize_max <= 32767.0f) type = AQH_SIGNED16;
				else type = AQH_SIGNED32;
			}
			m_type[d] = type;
			m_entry[d] = typeSize(type)*dd.n_channels;
		}
	}
	virtual TqInt Initialise() { return 0; }
	virtual TqInt Shutdown() { return 0; }
	virtual TqInt AddDisplay(const TqChar*, const TqChar*, const TqChar*, TqInt, TqInt, TqInt, std::map<std::string, void*>) { return 0; }
	virtual TqInt ClearDisplays() { return 0; }
	virtual TqInt OpenDisplays(TqInt, TqInt) { return 0; }
	virtual TqInt CloseDisplays() { return 0; }
	virtual bool fDisplayNeeds(const TqChar*) { return false; }
	virtual TqInt Uses() { return 0; }
	virtual TqInt numDisplayRequests() { return 0; }
	virtual boost::shared_ptr<IqDisplayRequest> displayRequest(TqInt) { return boost::shared_ptr<IqDisplayRequest>(); }
	virtual TqInt DisplayBucket(const CqRegion& DRegion, const IqChannelBuffer* pBuffer)
	{
		// CqDDManager::DisplayBucket, ddmanager.cpp:146-171: buckets outside the crop window are skipped
		CqRenderer* rc = QGetRenderContext();
		if(g_cull.n)
		{
			// the bucket processor this channel buffer is a member of (RenderImage passes &processor->getChannelBuffer())
			const CqBucketProcessor* bp = reinterpret_cast<const CqBucketProcessor*>(
				reinterpret_cast<const char*>(pBuffer) - offsetof(CqBucketProcessor, m_channelBuffer));
			// a bound that misses the tree's area is "culled" by canCull itself (the cropped bound is empty), so every
			// bucket can be asked: the surface survives if ANY bucket it could be re-posted to keeps it
			for(int i = 0; i < g_cull.n; ++i)
			{
				const float* b = g_cull.bounds + 6*i;
				CqBound bound(b[0], b[1], b[2], b[3], b[4], b[5]);
				if(!bp->m_OcclusionTree.canCull(bound)) g_cull.culledEverywhere[i] = 0;
			}
		}
		if(pBuffer->width() == 0 || pBuffer->height() == 0) return 0;
		if(DRegion.xMax() <= rc->cropWindowXMin() || DRegion.yMax() <= rc->cropWindowYMin() ||
		   DRegion.xMin() > rc->cropWindowXMax() || DRegion.yMin() > rc->cropWindowYMax())
			return 0;
		static const char* names[5] = {"Ci", "Oi", "a", "z", "coverage"};
		int idx[5];
		for(int i = 0; i < 5; ++i) idx[i] = pBuffer->getChannelIndex(names[i]);
		const int w = pBuffer->width(), h = pBuffer->height();
		int nch = 9, aovIdx[AQH_MAX_AOVS];
		for(int a = 0; a < m_p.n_aovs; ++a) { aovIdx[a] = pBuffer->getChannelIndex(m_p.aov[a].name); nch += m_p.aov[a].n_floats; }
		std::vector<float> px(size_t(w)*h*nch);
		for(int y = 0; y < h; ++y)
			for(int x = 0; x < w; ++x)
			{
				float* o = &px[(size_t(y)*w + x)*nch];
				const float* ci = (*pBuffer)(x, y, idx[0]);
				const float* oi = (*pBuffer)(x, y, idx[1]);
				o[0] = ci[0]; o[1] = ci[1]; o[2] = ci[2]; o[3] = oi[0]; o[4] = oi[1]; o[5] = oi[2];
				o[6] = (*pBuffer)(x, y, idx[2])[0];
				o[7] = (*pBuffer)(x, y, idx[3])[0];
				o[8] = (*pBuffer)(x, y, idx[4])[0];
				for(int a = 0, at = 9; a < m_p.n_aovs; at += m_p.aov[a].n_floats, ++a)
					for(int k = 0; k < m_p.aov[a].n_floats; ++k) o[at + k] = (*pBuffer)(x, y, aovIdx[a])[k];
				const int X = DRegion.xMin() + x, Y = DRegion.yMin() + y;
				if(m_channels && X < m_p.xres && Y < m_p.yres)
					std::memcpy(m_channels + (size_t(Y)*m_p.xres + X)*nch, o, size_t(nch)*4);
			}
		CqRandom random;
		for(int d = 0; d < m_p.n_displays; ++d)
		{
			const AqhDisplayDesc& dd = m_p.display[d];
			for(int y = 0; y < h; ++y)
				for(int x = 0; x < w; ++x)
				{
					double s = random.RandomFloat();
					const int X = DRegion.xMin() + x, Y = DRegion.yMin() + y;
					unsigned char* pdata = (m_displays && m_displays[d]) ? m_displays[d] + (size_t(Y)*m_p.xres + X)*m_entry[d] : 0;
					if(!pdata) continue;
					for(int c = 0; c < dd.n_channels; ++c)
					{
						double value = px[(size_t(y)*w + x)*nch + dd.channel[c]];
						if(dd.quantize_one != 0)
						{
							value = Aqsis::lround(dd.quantize_zero + value * (dd.quantize_one - dd.quantize_zero) + (dd.quantize_dither * s));
							value = clamp<double>(value, dd.quantize_min, dd.quantize_max);
						}
						switch(m_type[d])
						{
							case AQH_FLOAT32: { float v = value; std::memcpy(pdata, &v, 4); pdata += 4; break; }
							case AQH_UNSIGNED32: { value = clamp<double>(value, 0, std::numeric_limits<uint32_t>::max()); uint32_t v = static_cast<uint32_t>(value); std::memcpy(pdata, &v, 4); pdata += 4; break; }
							case AQH_SIGNED32: { value = clamp<double>(value, std::numeric_limits<int32_t>::min(), std::numeric_limits<int32_t>::max()); int32_t v = static_cast<int32_t>(value); std::memcpy(pdata, &v, 4); pdata += 4; break; }
							case AQH_UNSIGNED16: { uint16_t v = static_cast<uint16_t>(value); std::memcpy(pdata, &v, 2); pdata += 2; break; }
							case AQH_SIGNED16: { int16_t v = static_cast<int16_t>(value); std::memcpy(pdata, &v, 2); pdata += 2; break; }
							case AQH_UNSIGNED8: { *pdata++ = static_cast<unsigned char>(value); break; }
							case AQH_SIGNED8: { *pdata++ = static_cast<unsigned char>(static_cast<signed char>(value)); break; }
						}
					}
				}
		}
		return 0;
	}
private:
	const AqhFrameParams& m_p;
	float* m_channels;
	unsigned char* const* m_displays;
	int m_entry[AQH_MAX_DISPLAYS], m_type[AQH_MAX_DISPLAYS];
};

double nowS() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

template<class T> void setOpt(T* dst, std::initializer_list<T> v) { int i = 0; for(T x : v) dst[i++] = x; }

} // namespace

extern "C" {

// Render one frame with the reference's own hider.  Same contract as orc_render (oracle_hider.h);
// single-threaded like the reference.  Returns 0, or AQH_ERR_UNSUPPORTED for options the driver
// above does not map.
int ref_render(const AqhFrameParams* pp, const AqhGridBlock* grids, float* channels,
               unsigned char* const* display_out, OrcStats* stats)
{
	if(!pp || !grids || grids->memory_space != 0) return AQH_ERR_BAD_PARAMS;
	const AqhFrameParams& p = *pp;
	// like aqsis' own main(): only warnings and worse reach std::cerr (tools/aqsis/aqsis.cpp installs the same filter)
	static Aqsis::filter_by_level_buf* quiet = new Aqsis::filter_by_level_buf(Aqsis::WARNING, std::cerr);
	(void)quiet;
	const double t0 = nowS();

	// ---- render context and options (what RiCxxCore::WorldBegin leaves behind, ri.cpp:577-666)
	CqRenderer* rc = new CqRenderer;
	pCurrRenderer = rc;
	CqOptions& opt = *boost::static_pointer_cast<CqOptions>(boost::const_pointer_cast<IqOptions>(rc->poptCurrent()));
	setOpt<TqInt>(opt.GetIntegerOptionWrite("System", "Resolution", 2), {p.xres, p.yres});
	setOpt<TqInt>(opt.GetIntegerOptionWrite("System", "PixelSamples", 2), {p.xsamples, p.ysamples});
	setOpt<TqFloat>(opt.GetFloatOptionWrite("System", "FilterWidth", 2), {p.filter_xwidth, p.filter_ywidth});
	setOpt<TqFloat>(opt.GetFloatOptionWrite("System", "Clipping", 2), {p.clip_near, p.clip_far});
	setOpt<TqFloat>(opt.GetFloatOptionWrite("System", "Shutter", 2), {p.shutter_open, p.shutter_close});
	setOpt<TqFloat>(opt.GetFloatOptionWrite("System", "Exposure", 2), {p.exposure_gain, p.exposure_gamma});
	setOpt<TqInt>(opt.GetIntegerOptionWrite("System", "DisplayMode", 1), {p.display_mode});
	setOpt<TqInt>(opt.GetIntegerOptionWrite("limits", "bucketsize", 2), {p.bucket_xsize, p.bucket_ysize});
	setOpt<TqInt>(opt.GetIntegerOptionWrite("Hider", "jitter", 1), {p.jitter});
	{
		// Hider "depthfilter" and limits "zthreshold" (optioncache.cpp:95-116)
		static const char* const names[4] = {"min", "midpoint", "max", "average"};
		if(p.depth_filter < 0 || p.depth_filter > 3) return AQH_ERR_BAD_PARAMS;
		opt.GetStringOptionWrite("Hider", "depthfilter", 1)[0] = names[p.depth_filter];
		opt.GetColorOptionWrite("limits", "zthreshold", 1)[0] = CqColor(p.zthreshold[0], p.zthreshold[1], p.zthreshold[2]);
	}
	opt.SetfuncFilter(referenceFilterFor(p.filter_func));
	CqImageBuffer* image = new CqImageBuffer;
	RefDDManager dd(p, channels, display_out);
	g_refSetup.p = &p; g_refSetup.image = image; g_refSetup.dd = &dd;
	rc->Initialise();

	// ---- the global random stream: Reseed at WorldBegin (ri.cpp:660), then any front-end draws
	CqRandom rng;
	rng.Reseed(p.rng_seed);
	for(uint32_t i = 0; i < p.rng_predraws; ++i) rng.RandomInt();

	image->SetImage();

	// ---- bust every grid and hand the micropolygons to CqImageBuffer::AddMPG
	const double t1 = nowS();
	CqMatrix camToRaster;
	for(int i = 0; i < 4; ++i) for(int j = 0; j < 4; ++j) camToRaster.SetElement(i, j, p.cam_to_raster[i*4+j]);
	camToRaster.SetfIdentity(false);
	std::vector<RefGrid*> keep;
	std::vector<std::vector<float> > owned;       // projected copies of P (the caller's arrays are const)
	size_t po = 0, vo = 0, ko = 0;
	int64_t nmp = 0;
	int aovFloats = 0;
	for(int a = 0; a < p.n_aovs; ++a) aovFloats += p.aov[a].n_floats;
	for(int64_t g = 0; g < grids->n_grids; ++g)
	{
		const int cu = grids->cu[g], cv = grids->cv[g];
		const int nk = grids->nkeys ? grids->nkeys[g] : 1;
		const uint32_t flags = grids->flags[g];
		if(flags & AQH_GRID_POINTS) return AQH_ERR_UNSUPPORTED;      // geometry/points.cpp needs the whole CqSurface family: restated in the oracle only
		const size_t nv = size_t(cu+1)*(cv+1);
		owned.push_back(std::vector<float>(grids->P + po*3, grids->P + (po + nv*nk)*3));
		float* P = owned.back().data();
		if(flags & AQH_GRID_CAMERA_SPACE)
		{
			// micropolygon.cpp:723-731: raster x,y from the camera-to-raster matrix, camera z kept
			for(size_t i = 0; i < nv*nk; ++i)
			{
				CqVector3D pt(P[3*i], P[3*i+1], P[3*i+2]);
				TqFloat zdepth = pt.z();
				CqVector3D r = camToRaster * pt;
				P[3*i] = r.x(); P[3*i+1] = r.y(); P[3*i+2] = zdepth;
			}
		}
		owned.push_back(std::vector<float>());
		float* Ci = 0; float* Oi = 0;
		if(grids->Ci) { owned.push_back(std::vector<float>(grids->Ci + vo*3, grids->Ci + (vo+nv)*3)); Ci = owned.back().data(); }
		if(grids->Oi) { owned.push_back(std::vector<float>(grids->Oi + vo*3, grids->Oi + (vo+nv)*3)); Oi = owned.back().data(); }
		RefGrid* grid = new RefGrid(cu, cv, P, Ci, Oi, flags, grids->lod_bounds ? grids->lod_bounds + 2*g : 0);
		ADDREF(grid);
		keep.push_back(grid);
		if(flags & AQH_GRID_USES_CSG)
		{
			const int node = grids->csg_node ? grids->csg_node[g] : -1;
			if(node < 0 || node >= (int)g_csgNodes.size()) return AQH_ERR_BAD_PARAMS;
			grid->setCSGNode(g_csgNodes[node]);
		}
		const int trimSet = (grids->trim_set && grids->trim_uv) ? grids->trim_set[g] : 0;
		const bool bOutside = (flags & AQH_GRID_TRIM_OUTSIDE) != 0;
		if(trimSet != 0)
		{
			if(trimSet < 0 || trimSet > (int)g_trimSurfaces.size()) return AQH_ERR_BAD_PARAMS;
			grid->setTrim(reinterpret_cast<CqSurface*>(&g_trimSurfaces[trimSet - 1]), IqConstAttributesPtr(new RefAttributes(bOutside)),
			              grids->trim_uv + vo*2, (int)nv);
		}
		if(aovFloats && grids->aov)
			for(int a = 0, at = 0; a < p.n_aovs; at += p.aov[a].n_floats, ++a)
				grid->addAov(p.aov[a].name, grids->aov + vo*aovFloats, aovFloats, at, p.aov[a].n_floats);
		// the culls CqMicroPolyGrid::Shade applies before the grid reaches Split (micropolygon.cpp:431-474, 493-522),
		// restated: Shade() itself needs the shader VM
		std::vector<unsigned char> culledAll;
		if(flags & (AQH_GRID_CULL_BACKFACING | AQH_GRID_CULL_TRANSPARENT))
		{
			culledAll.assign(nv, 0);
			if(grids->culled) for(size_t i = 0; i < nv; ++i) culledAll[i] = grids->culled[vo + i];
			if((flags & AQH_GRID_CULL_BACKFACING) && grids->Ng && !(flags & AQH_GRID_USES_CSG))
			{
				const CqVector3D* pP = reinterpret_cast<const CqVector3D*>(grids->P + po*3);     // camera space, the shaded key
				const CqVector3D* pNg = reinterpret_cast<const CqVector3D*>(grids->Ng + vo*3);
				const CqVector3D* pN = grids->N ? reinterpret_cast<const CqVector3D*>(grids->N + vo*3) : 0;
				for(TqInt i = TqInt(nv) - 1; i >= 0; i--)
				{
					TqFloat s = 1.0f;
					if(NULL != pN)
						s = ((pN[i] * pNg[i]) < 0.0f) ? -1.0f : 1.0f;
					if(((s * pNg[i]) * pP[i]) >= 0)
						culledAll[i] = 1;
				}
			}
			const CqColor zThr(p.zthreshold[0], p.zthreshold[1], p.zthreshold[2]);
			if((flags & AQH_GRID_CULL_TRANSPARENT) && grids->Oi && !(zThr == gColBlack))
			{
				const CqColor* pOi = reinterpret_cast<const CqColor*>(grids->Oi + vo*3);
				for(TqInt i = TqInt(nv) - 1; i >= 0; i--)
				{
					if(pOi[i] == gColBlack)
						culledAll[i] = 1;
					else
						break;
				}
			}
		}
		// triangle split line per key, micropolygon.cpp:733-749
		for(int k = 0; k < nk; ++k)
		{
			const float* Pk = P + size_t(k)*nv*3;
			CqVector3D v0(Pk[0], Pk[1], Pk[2]), v1(Pk[3*cu], Pk[3*cu+1], Pk[3*cu+2]);
			const size_t c = size_t(cv)*(cu+1);
			CqVector3D v2(Pk[3*c], Pk[3*c+1], Pk[3*c+2]);
			const TqFloat time = nk > 1 ? grids->key_times[ko + k] : 0.0f;
			if(((v1.x() - v0.x())*(v2.y() - v0.y()) - (v1.y() - v0.y())*(v2.x() - v0.x())) >= 0)
				grid->addSplitLine(time, v1, v2);
			else
				grid->addSplitLine(time, v2, v1);
		}
		// micropolygon.cpp:770-856
		for(int iv = 0; iv < cv; ++iv)
			for(int iu = 0; iu < cu; ++iu)
			{
				const int iIndex = iv*(cu+1) + iu;
				if(culledAll.empty() ? (grids->culled && grids->culled[vo + iIndex]) : culledAll[iIndex] != 0) continue;
				// micropolygon.cpp:784-835 (restated like the rest of the loop; the queries are the surface's own)
				bool fTrimmed = false;
				if(trimSet != 0 && grid->pSurface()->bCanBeTrimmed())
				{
					const float* t = grids->trim_uv + vo*2;
					CqVector2D vecA(t[2*iIndex], t[2*iIndex+1]), vecB(t[2*(iIndex+1)], t[2*(iIndex+1)+1]);
					CqVector2D vecC(t[2*(iIndex+cu+2)], t[2*(iIndex+cu+2)+1]), vecD(t[2*(iIndex+cu+1)], t[2*(iIndex+cu+1)+1]);
					bool fTrimA = grid->pSurface()->bIsPointTrimmed(vecA), fTrimB = grid->pSurface()->bIsPointTrimmed(vecB);
					bool fTrimC = grid->pSurface()->bIsPointTrimmed(vecC), fTrimD = grid->pSurface()->bIsPointTrimmed(vecD);
					if(bOutside) { fTrimA = !fTrimA; fTrimB = !fTrimB; fTrimC = !fTrimC; fTrimD = !fTrimD; }
					if(fTrimA && fTrimB && fTrimC && fTrimD)
						if(!grid->pSurface()->bIsLineIntersecting(vecA, vecB) && !grid->pSurface()->bIsLineIntersecting(vecB, vecC) &&
						   !grid->pSurface()->bIsLineIntersecting(vecC, vecD) && !grid->pSurface()->bIsLineIntersecting(vecD, vecA))
							continue;
					if(fTrimA || fTrimB || fTrimC || fTrimD) fTrimmed = true;
				}
				++nmp;
				if(nk > 1)
				{
					boost::shared_ptr<CqMicroPolygonMotion> pNew(new CqMicroPolygonMotion(grid, iIndex));
					if(fTrimmed) pNew->MarkTrimmed();
					for(int k = 0; k < nk; ++k)
					{
						const CqVector3D* Pk = reinterpret_cast<const CqVector3D*>(P + size_t(k)*nv*3);
						pNew->AppendKey(Pk[iIndex], Pk[iIndex+1], Pk[iIndex+cu+1], Pk[iIndex+cu+2], grids->key_times[ko + k]);
					}
					pNew->Initialise();
					boost::shared_ptr<CqMicroPolygon> pTemp(pNew);
					image->AddMPG(pTemp);
				}
				else
				{
					boost::shared_ptr<CqMicroPolygon> pNew(new CqMicroPolygon(grid, iIndex));
					if(fTrimmed) pNew->MarkTrimmed();
					pNew->Initialise();
					image->AddMPG(pNew);
				}
			}
		po += nv*nk; vo += nv; ko += nk;
	}
	const double t2 = nowS();

	// ---- the reference's bucket loop
	image->RenderImage();
	const double t3 = nowS();

	if(stats)
	{
		std::memset(stats, 0, sizeof *stats);
		stats->prepare_s = t1 - t0; stats->bust_s = t2 - t1; stats->render_s = t3 - t2; stats->total_s = t3 - t0;
		stats->n_micropolygons = nmp;
		stats->n_samples = int64_t(p.crop_xmax - p.crop_xmin)*(p.crop_ymax - p.crop_ymin)*p.xsamples*p.ysamples;
		stats->threads = 1;
	}
	delete image;
	for(RefGrid* g : keep) RELEASEREF(g);
	pCurrRenderer = 0;
	delete rc;
	return AQH_OK;
}

// Choose the pixel filter of the following ref_render calls by its RenderMan name ("box", "triangle", "gaussian",
// "catmull-rom", "sinc", "mitchell", "disk", "bessel"): the reference's own Ri*Filter.  NULL or "" returns to
// recognising AqhFrameParams::filter_func.  Returns 0, or AQH_ERR_BAD_PARAMS for an unknown name.
int ref_set_filter(const char* name)
{
	g_forcedFilter = 0;
	if(!name || !*name) return AQH_OK;
	for(const RefFilter& rf : kRefFilters)
		if(std::strcmp(rf.name, name) == 0) { g_forcedFilter = rf.fn; return AQH_OK; }
	return AQH_ERR_BAD_PARAMS;
}

// The CSG tree of the following ref_render calls, built from real CqCSGTreeNode objects (csgtree.cpp): type[i] is an
// AQH_CSG_* value, parent[i] the parent node or -1; children are attached in node-index order.  0 nodes = none.
int ref_set_csg_tree(int n_nodes, const int32_t* type, const int32_t* parent)
{
	g_csgNodes.clear();
	CqCSGTreeNode::SetRequired(false);
	if(n_nodes <= 0) return AQH_OK;
	static const char* const names[4] = {"primitive", "union", "intersection", "difference"};
	for(int i = 0; i < n_nodes; ++i)
	{
		if(type[i] < 0 || type[i] > 3 || parent[i] >= n_nodes || parent[i] == i) { g_csgNodes.clear(); return AQH_ERR_BAD_PARAMS; }
		CqString t(names[type[i]]);
		g_csgNodes.push_back(CqCSGTreeNode::CreateNode(t));
	}
	for(int i = 0; i < n_nodes; ++i)
		if(parent[i] >= 0) g_csgNodes[parent[i]]->AddChild(g_csgNodes[i]);
	return AQH_OK;
}

// The trim loops of the following ref_render calls: one CqTrimLoopArray per set, the tessellated points put straight into
// CqTrimLoop::m_aCurvePoints (what CqTrimLoop::Prepare leaves there).  0 sets = none.
int ref_set_trim_loops(int n_sets, const int32_t* set_first_loop, const int32_t* loop_first_point, const float* points)
{
	g_trimSets.clear(); g_trimSurfaces.clear();
	if(n_sets <= 0) return AQH_OK;
	g_trimSets.resize(n_sets);
	for(int s = 0; s < n_sets; ++s)
		for(int l = set_first_loop[s]; l < set_first_loop[s+1]; ++l)
		{
			CqTrimLoop loop;
			for(int i = loop_first_point[l]; i < loop_first_point[l+1]; ++i)
				loop.m_aCurvePoints.push_back(CqVector2D(points[2*i], points[2*i+1]));
			g_trimSets[s].m_aLoops.push_back(loop);
		}
	g_trimSurfaces.resize(n_sets);
	for(int s = 0; s < n_sets; ++s) { g_trimSurfaces[s].vptr = refSurfaceVtable(); g_trimSurfaces[s].loops = &g_trimSets[s]; }
	return AQH_OK;
}
// CqTrimLoopArray::TrimPoint / LineIntersects of set (1-based) -- the leaves the oracle's restatement is pinned against
int ref_trim_point(int set, float x, float y) { return g_trimSets[set - 1].TrimPoint(CqVector2D(x, y)) ? 1 : 0; }
int ref_trim_line(int set, float x1, float y1, float x2, float y2) { return g_trimSets[set - 1].LineIntersects(CqVector2D(x1, y1), CqVector2D(x2, y2)) ? 1 : 0; }

// What aqsis' occlusion culling would do with surfaces of the given raster bounds (xmin, ymin, zmin, xmax, ymax, zmax)
// arriving after all of `grids` has been rendered: culled[i] = 1 when CqOcclusionTree::canCull(bound) holds in EVERY
// bucket whose sample region the bound touches (RenderSurface then re-posts the surface from bucket to bucket until it
// falls off the image, bucketprocessor.cpp:945-958), or when it touches none.
int ref_can_cull(const AqhFrameParams* pp, const AqhGridBlock* grids, int n_bounds, const float* bounds, unsigned char* culled)
{
	if(n_bounds < 0 || (n_bounds && (!bounds || !culled))) return AQH_ERR_BAD_PARAMS;
	for(int i = 0; i < n_bounds; ++i) culled[i] = 1;
	g_cull.n = n_bounds; g_cull.bounds = bounds; g_cull.culledEverywhere = culled;
	const int rc = ref_render(pp, grids, 0, 0, 0);
	g_cull.n = 0;
	return rc;
}

} // extern "C"
