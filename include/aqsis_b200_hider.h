/* aqsis_b200_hider.h -- C ABI of the B200-native REYES hider + pixel filter.
 *
 * This is the drop-in boundary for the post-shading hot path of aqsis
 * (SURVEY.md section 8b).  The reference has no plugin seam around its hider:
 * it is reached by direct C++ calls.  Every entry point below therefore names
 * the reference interface it stands in for (paths relative to the aqsis tree):
 *
 *   aqh_begin_frame   <- CqImageBuffer::SetImage()            libs/core/imagebuffer.cpp:174-228
 *                        SqOptionCache::cacheOptions()        libs/core/optioncache.cpp:50-117
 *                        CqRenderer::SetDepthOfFieldData/...  libs/core/renderer.h:368-406
 *                        CqMultiJitteredSampler ctor          libs/core/multijitter.h:90-101
 *                        CqBucketProcessor::InitialiseFilterValues  bucketprocessor.cpp:811-856
 *   aqh_add_grid      <- CqMicroPolyGridBase::Split(xmin,xmax,ymin,ymax)  libs/core/micropolygon.h:115
 *                        (call site bucketprocessor.cpp:1036) which busts the grid and feeds
 *                        CqImageBuffer::AddMPG(shared_ptr<CqMicroPolygon>&)  imagebuffer.h:91
 *   aqh_add_grid_block<- the same, for many grids at once (host or device resident)
 *   aqh_end_frame     <- CqImageBuffer::RenderImage()         libs/core/imagebuffer.cpp:605-779
 *                        and, per bucket, IqDDManager::DisplayBucket(CqRegion, IqChannelBuffer*)
 *                        libs/core/ddmanager/iddmanager.h:99 -> FormatBucketForDisplay
 *                        ddmanager.cpp:1022-1118 -> DspyImageData  include/aqsis/ri/ndspy.h:186-192
 *   AqhBucketFunc     <- IqDDManager::DisplayBucket            (float channel buffer per bucket)
 *   AqhDataFunc       <- DspyImageDataMethod                   include/aqsis/ri/ndspy.h:157
 *   AqhProgressFunc   <- RtProgressFunc                        include/aqsis/ri/ritypes.h:73
 *   AqhFilterFunc     <- RtFilterFunc                          include/aqsis/ri/ritypes.h:56
 *   aqh_*_filter      <- RiGaussianFilter & co                 libs/core/filters.cpp:71-348
 *
 * Conventions: plain C, plain pointers and sizes, no exceptions cross the ABI,
 * every function returns an AqhStatus (0 = ok).  One calling thread per handle.
 * All arithmetic on the path is IEEE binary32 evaluated without fused
 * multiply-add, exactly as the reference's x86-64 build does.
 */
#ifndef AQSIS_B200_HIDER_H_INCLUDED
#define AQSIS_B200_HIDER_H_INCLUDED

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#  define AQH_EXPORT __attribute__((visibility("default")))
#else
#  define AQH_EXPORT
#endif

#define AQH_ABI_VERSION 3

typedef enum AqhStatus
{
	AQH_OK = 0,
	AQH_ERR_BAD_PARAMS = 1,     /* like PkDspyErrorBadParams */
	AQH_ERR_NO_MEMORY = 2,      /* host or device allocation failed */
	AQH_ERR_UNSUPPORTED = 3,    /* feature outside the implemented path (CSG, trim curves, ...) */
	AQH_ERR_NO_DEVICE = 4,      /* no CUDA device / kernel image: there is NO cpu fallback */
	AQH_ERR_CUDA = 5,           /* a CUDA call failed; see aqh_last_error */
	AQH_ERR_STATE = 6,          /* call out of order (add_grid outside begin/end, ...) */
	AQH_ERR_DEEP_OVERFLOW = 7,  /* transparent hit pool exhausted; raise deep_hits_per_sample */
	AQH_ERR_CALLBACK = 8        /* a display callback returned non-zero */
} AqhStatus;

/* --- enums mirrored from the reference ---------------------------------- */

/* EqDepthFilter, libs/core/optioncache.h:34-40 */
enum { AQH_DEPTHFILTER_MIN = 0, AQH_DEPTHFILTER_MIDPOINT = 1, AQH_DEPTHFILTER_MAX = 2, AQH_DEPTHFILTER_AVERAGE = 3 };
/* EqDisplayMode, include/aqsis/core/ioptions.h:50-56 */
enum { AQH_DMODE_RGB = 1, AQH_DMODE_A = 2, AQH_DMODE_Z = 4 };
/* PkDspy* data types, include/aqsis/ri/ndspy.h:70-76 */
enum { AQH_FLOAT32 = 1, AQH_UNSIGNED32 = 2, AQH_SIGNED32 = 3, AQH_UNSIGNED16 = 4, AQH_SIGNED16 = 5,
       AQH_UNSIGNED8 = 6, AQH_SIGNED8 = 7 };
/* Slots of the per-pixel float channel buffer handed to DisplayBucket
 * (addChannel order in CqBucketProcessor::FilterBucket, bucketprocessor.cpp:389-393). */
enum { AQH_CH_CI_R = 0, AQH_CH_CI_G = 1, AQH_CH_CI_B = 2, AQH_CH_OI_R = 3, AQH_CH_OI_G = 4, AQH_CH_OI_B = 5,
       AQH_CH_ALPHA = 6, AQH_CH_Z = 7, AQH_CH_COVERAGE = 8, AQH_NUM_CHANNELS = 9 };

/* Grid flags (SqGridInfo, libs/core/micropolygon.h:57-63; fTriangular, usesCSG). */
enum {
	AQH_GRID_SMOOTH       = 1u << 0,  /* useSmoothShading: bilinear Ci/Oi, else corner 0 */
	AQH_GRID_MATTE        = 1u << 1,  /* SqImageSample::Flag_Matte */
	AQH_GRID_MATTE_ALPHA  = 1u << 2,  /* SqImageSample::Flag_MatteAlpha */
	AQH_GRID_TRIANGULAR   = 1u << 3,  /* fTriangular(): reject hits beyond the split line */
	AQH_GRID_CAMERA_SPACE = 1u << 4,  /* P is camera space: project with cam_to_raster keeping camera z
	                                     (micropolygon.cpp:723-731); else P is already raster x,y + camera z */
	AQH_GRID_USES_CSG     = 1u << 5,  /* pCSGNode() != 0: hits are never culled, kept in the sample's list and resolved
	                                     through the CSG tree (imagepixel.cpp:166-189, csgtree.cpp:144-351); csg_node
	                                     names the primitive node in the tree given to aqh_set_csg_tree */
	AQH_GRID_POINTS       = 1u << 6,  /* CqMicroPolyGridPoints (geometry/points.cpp:653-700): every vertex is a disc of
	                                     raster radius `radius`; cv must be 0, the grid has cu+1 points; constant shading */
	AQH_GRID_CULL_BACKFACING  = 1u << 7, /* Sides 1: drop micropolygons whose corner normal faces away, (Ng . P) >= 0 in
	                                     camera space (micropolygon.cpp:431-474); needs Ng and AQH_GRID_CAMERA_SPACE */
	AQH_GRID_CULL_TRANSPARENT = 1u << 8, /* drop the trailing run of micropolygons with Oi == 0 (micropolygon.cpp:493-522,
	                                     including the reference's early break at the first non-black vertex) */
	AQH_GRID_TRIM_OUTSIDE     = 1u << 9  /* Attribute "trimcurve" "sense" "outside" of a trimmed surface (micropolygon.cpp:667-671, 1597-1601) */
};

enum { AQH_MAX_DISPLAYS = 8, AQH_MAX_DISPLAY_CHANNELS = 16, AQH_MAX_RANKS = 64, AQH_MAX_AOVS = 8, AQH_MAX_AOV_FLOATS = 21 };

/* One arbitrary output variable (CqRenderer::RegisterOutputData, libs/core/renderer.cpp:1520-1546): `n_floats`
 * consecutive floats of every hit (StoreExtraData, bucketprocessor.cpp:1573-1643: float 1, point/normal/vector/color 3,
 * matrix 16), filtered like colour (bucketprocessor.cpp:620-653) and appended to the 9 standard floats of a pixel of
 * the channel buffer in registration order: the first AOV starts at slot AQH_NUM_CHANNELS. */
typedef struct AqhAovDesc
{
	char name[32];                  /* the shader output variable, e.g. "N" or "_albedo" (informational on this side) */
	int32_t n_floats;               /* 1, 3 or 16 */
	int32_t reserved;
} AqhAovDesc;

/* AqhDisplayDesc::flags */
enum { AQH_DISPLAY_SCANLINE_ORDER = 1 };  /* PkDspyFlagsWantsScanLineOrder (ndspy.h:132): rows are delivered one at a time once
                                             a row of buckets is complete (CollapseBucketsToScanlines / SendToDisplay,
                                             ddmanager.cpp:1129-1175) */

/* How the pixel filter sums are associated (the SET of samples and the weights are always the
 * reference's: inclusion by jittered position, weight by sub-pixel cell centre).
 *   REFERENCE_ORDER one running sum per output pixel in the reference's fy, fx, sy, sx order
 *                   (bucketprocessor.cpp:597-629): bit-identical float results, at the price of
 *                   writing every resolved sample to HBM once.  The default.
 *   TILE_PARTIALS   per (source pixel, filter tap) partial sums over the pixel's samples, formed
 *                   on chip by the hide kernel, then summed over the taps: deterministic, but the
 *                   different association shows as rounding noise (~1e-6 relative for positive
 *                   filters, up to ~1e-4 on dark pixels under negative-lobed filters). */
enum { AQH_FILTER_REFERENCE_ORDER = 0, AQH_FILTER_TILE_PARTIALS = 1 };

/* RtFilterFunc (include/aqsis/ri/ritypes.h:56): only ever tabulated on the host
 * (bucketprocessor.cpp:850), so a user filter works unchanged. */
typedef float (*AqhFilterFunc)(float x, float y, float xwidth, float ywidth);

/* One display request: channel selection + quantisation
 * (CqDisplayRequest m_formats/m_bufferMap + m_Quantize*, ddmanager.cpp:455-480,1046-1113). */
typedef struct AqhDisplayDesc
{
	int32_t n_channels;                          /* entries of channel[] used */
	int32_t channel[AQH_MAX_DISPLAY_CHANNELS];   /* AQH_CH_* slot per output element, in driver order; slots >= AQH_NUM_CHANNELS
	                                                are floats of the arbitrary output variables */
	int32_t type;                                /* AQH_FLOAT32...; 0 = select from (one,min,max) like selectDataFormat */
	float quantize_zero, quantize_one, quantize_min, quantize_max, quantize_dither;
	int32_t flags;                               /* AQH_DISPLAY_* */
} AqhDisplayDesc;

typedef struct AqhFrameParams
{
	int32_t abi_version;            /* AQH_ABI_VERSION */
	int32_t xres, yres;             /* System:Resolution */
	int32_t crop_xmin, crop_xmax, crop_ymin, crop_ymax; /* integer crop window, renderer.cpp:1619-1627 */
	int32_t xsamples, ysamples;     /* PixelSamples */
	float   filter_xwidth, filter_ywidth;
	AqhFilterFunc filter_func;      /* NULL = gaussian */
	int32_t bucket_xsize, bucket_ysize; /* limits:bucketsize, default 16 16; fixes RNG replay + callback order */
	float   clip_near, clip_far;
	float   shutter_open, shutter_close;
	int32_t use_dof;                /* CqRenderer::UsingDepthOfField() */
	float   dof_multiplier;         /* m_DofMultiplier         renderer.h:368-377 */
	float   dof_one_over_focal_distance;
	float   dof_scale_x, dof_scale_y; /* m_DepthOfFieldScale   options.cpp:162-171 */
	int32_t depth_filter;           /* AQH_DEPTHFILTER_*: min, midpoint, max, average (imagepixel.cpp:264-327) */
	float   zthreshold[3];          /* limits:zthreshold, default 1 1 1 */
	int32_t display_mode;           /* AQH_DMODE_* union over displays */
	float   exposure_gain, exposure_gamma;
	int32_t jitter;                 /* Hider "jitter": 1 = CqMultiJitteredSampler, 0 = CqGridSampler */
	float   cam_to_raster[16];      /* row-major m[i][j] of CqMatrix, used with AQH_GRID_CAMERA_SPACE */
	uint32_t rng_seed;              /* seed of the global CqRandom at WorldBegin: 545 (ri.cpp:660) */
	uint32_t rng_predraws;          /* front-end draws made between the reseed and RenderImage() */
	int32_t n_displays;
	AqhDisplayDesc display[AQH_MAX_DISPLAYS];
	/* arbitrary output variables, in registration order (RiDisplay -> RegisterOutputData) */
	int32_t n_aovs;
	AqhAovDesc aov[AQH_MAX_AOVS];
	/* --- device-side knobs (no reference analogue) --- */
	int32_t rank, world_size;       /* pixel-row strips of the image are dealt to ranks; 0,1 = whole image */
	int32_t strip_rows;             /* > 0: strips of this many rows (rounded down to a multiple of 16) dealt round-robin;
	                                   0: world*k near-equal strips dealt round-robin;
	                                   -1: ONE contiguous strip per rank, near-equal heights (least replication of straddlers);
	                                   -2: one contiguous strip per rank, rank r owns rows [strip_bounds[r], strip_bounds[r+1])
	                                       (see aqh_balance_strips) */
	int32_t strip_bounds[AQH_MAX_RANKS + 1];
	int32_t deep_hits_per_sample;   /* capacity for transparent hits: every sample has 8 in-line slots, the pool behind them holds
	                                   (deep_hits_per_sample - 8) hits per sample averaged over a tile; 0 = default (16) */
	int32_t filter_mode;            /* AQH_FILTER_*; 0 = AQH_FILTER_REFERENCE_ORDER (bit-exact sums) */
	int32_t plane_budget_mb;        /* HBM the resolved samples of the reference-order filter may occupy at a time; the frame
	                                   is hidden and filtered in bands of tile rows that fit (0 = 4608 MB) */
	int32_t reserved[6];
} AqhFrameParams;

/* One shaded grid as CqMicroPolyGrid::Split sees it (micropolygon.cpp:641-892, motion :946-1156).
 * Arrays are AoS xyz / rgb floats exactly like IqShaderData::GetPointPtr / GetColorPtr
 * (include/aqsis/shadervm/ishaderdata.h:72-95); they are copied, the caller keeps ownership. */
typedef struct AqhGridDesc
{
	int32_t cu, cv;                 /* uGridRes(), vGridRes(): (cu+1)*(cv+1) vertices, cu*cv micropolygons */
	int32_t nkeys;                  /* 1 = static; >1 = CqMotionMicroPolyGrid key grids */
	const float* key_times;         /* nkeys shutter times (ignored when nkeys == 1) */
	const float* const* P;          /* nkeys pointers to (cu+1)*(cv+1)*3 floats */
	const float* Ci;                /* (cu+1)*(cv+1)*3 floats, NULL = white (micropolygon.cpp:1472-1475) */
	const float* Oi;                /* idem, NULL = opaque */
	const uint8_t* culled;          /* (cu+1)*(cv+1) bytes, non-zero = m_CulledPolys.Value(iIndex); NULL = none */
	uint32_t flags;                 /* AQH_GRID_* */
	float lod_bounds[2];            /* SqGridInfo::lodBounds; lod_bounds[0] < 0 = no level of detail */
	const float* aov;               /* nverts * (sum of the frame's AqhAovDesc::n_floats) floats, vertex-major: what
	                                   FindStandardVar(name)->Get*(value, index) returns; NULL = zeros (the reference would
	                                   leave whatever the pixel pool held) */
	const float* Ng;                /* nverts*3 camera-space geometric normals (AQH_GRID_CULL_BACKFACING), else NULL */
	const float* N;                 /* nverts*3 user normals deciding the facing of Ng (micropolygon.cpp:452-458), may be NULL */
	const float* radius;            /* AQH_GRID_POINTS: nkeys*nverts raster radii, key-major like P */
	int32_t csg_node;               /* AQH_GRID_USES_CSG: index of the grid's primitive node in the CSG tree */
	int32_t trim_set;               /* index of the surface's trim loops in the table given to aqh_set_trim_loops + 1; 0 = the
	                                   surface cannot be trimmed (CqSurface::bCanBeTrimmed) */
	const float* trim_uv;           /* trimmed surfaces: nverts*2 floats, the surface parameters (u, v) of every vertex
	                                   (pVar(EnvVars_u), pVar(EnvVars_v)) */
} AqhGridDesc;

/* Many grids, concatenated.  memory_space 0 = host pointers, 1 = device pointers
 * (grids already resident in HBM: nothing is staged through the host). */
typedef struct AqhGridBlock
{
	int64_t n_grids;
	const int32_t* cu;              /* n_grids */
	const int32_t* cv;              /* n_grids */
	const int32_t* nkeys;           /* n_grids, NULL = all 1 */
	const uint32_t* flags;          /* n_grids */
	const float* lod_bounds;        /* 2*n_grids, NULL = none */
	const float* key_times;         /* sum(nkeys) floats, grid-major; NULL when all static */
	const float* P;                 /* sum(nkeys*nverts)*3 floats: per grid, key-major, AoS xyz */
	const float* Ci;                /* sum(nverts)*3 floats, NULL = white */
	const float* Oi;                /* sum(nverts)*3 floats, NULL = opaque */
	const uint8_t* culled;          /* sum(nverts) bytes, NULL = none */
	int32_t memory_space;           /* applies to P, Ci, Oi, culled, aov, Ng, N, radius; the per-grid tables are always host */
	int32_t reserved[3];
	const float* aov;               /* sum(nverts) * (sum of AqhAovDesc::n_floats) floats, vertex-major; NULL = zeros */
	const float* Ng;                /* sum(nverts)*3, NULL unless some grid has AQH_GRID_CULL_BACKFACING */
	const float* N;                 /* sum(nverts)*3 or NULL */
	const float* radius;            /* sum(nkeys*nverts) floats (only read for AQH_GRID_POINTS grids) or NULL */
	const int32_t* csg_node;        /* n_grids, NULL = none */
	const int32_t* trim_set;        /* n_grids (see AqhGridDesc::trim_set), NULL = no trimmed surfaces */
	const float* trim_uv;           /* sum(nverts)*2 floats (u, v), read for the grids with a trim set; follows memory_space */
} AqhGridBlock;

/* IqDDManager::DisplayBucket stand-in: region [xmin,xmax1) x [ymin,ymax1) of the bucket and its float channel
 * buffer: pixel_stride_floats (= AQH_NUM_CHANNELS + the frame's AOV floats) interleaved floats per pixel, row stride
 * in floats. */
typedef int (*AqhBucketFunc)(void* user, int xmin, int xmax1, int ymin, int ymax1,
                             const float* channels, int row_stride_floats, int pixel_stride_floats);
/* Imager shader stand-in (CqBucketProcessor::FilterBucket, bucketprocessor.cpp:712-743): called once per bucket in
 * reference bucket order after filtering and BEFORE exposure and quantisation; may overwrite Ci (slots 0-2), Oi (3-5)
 * and alpha (6) of the bucket's pixels in place. */
typedef int (*AqhImagerFunc)(void* user, int xmin, int xmax1, int ymin, int ymax1,
                             float* channels, int row_stride_floats, int pixel_stride_floats);
/* DspyImageDataMethod stand-in (ndspy.h:157), one call per bucket per display, reference bucket order. */
typedef int (*AqhDataFunc)(void* user, int display, int xmin, int xmax1, int ymin, int ymax1,
                           int entrysize, const unsigned char* data);
typedef void (*AqhProgressFunc)(void* user, float percent_complete);

typedef struct AqhCallbacks
{
	void* user;
	AqhBucketFunc on_bucket;        /* may be NULL */
	AqhDataFunc on_data;            /* may be NULL */
	AqhProgressFunc on_progress;    /* may be NULL */
	AqhImagerFunc on_imager;        /* may be NULL */
} AqhCallbacks;

/* Stage timings of the last frame in milliseconds, named after the reference's
 * timers (libs/core/stats.h:70-110). Device stages are CUDA-event times. */
typedef struct AqhFrameStats
{
	double prepare_ms;        /* host: RNG replay + tables (Prepare_bucket) */
	double upload_ms;         /* H2D of grids + tables */
	double project_bust_ms;   /* Project_points + Bust_grids (+ binning) */
	double render_mpgs_ms;    /* Render_MPGs + Combine_samples */
	double filter_ms;         /* Filter_samples (+ expose) */
	double display_ms;        /* Display_bucket: quantise */
	double download_ms;       /* D2H of the image(s) */
	double device_total_ms;   /* first kernel to last kernel */
	int64_t n_grids, n_vertices, n_micropolygons, n_bin_entries, n_samples, n_deep_hits;
	int64_t gpu_launches;     /* kernels launched for the frame */
	int64_t h2d_bytes, d2h_bytes;
	int64_t device_bytes;     /* HBM held by the hider's own buffers after the frame (caller-owned device grids excluded) */
	int64_t n_bands;          /* bands of tile rows the frame was hidden and filtered in (AqhFrameParams::plane_budget_mb) */
	double gather_ms;         /* NCCL gather of the finished strips (aqh_gather), device time */
} AqhFrameStats;

typedef struct AqhHider AqhHider;

/* Create / destroy a hider bound to one CUDA device.  Fails with AQH_ERR_NO_DEVICE when
 * there is no usable sm_100 device: the product has no CPU path. */
AQH_EXPORT int aqh_create(AqhHider** out, int device);
AQH_EXPORT int aqh_destroy(AqhHider* h);
AQH_EXPORT int aqh_abi_version(void);
AQH_EXPORT const char* aqh_last_error(const AqhHider* h);
/* Use a caller-owned cudaStream_t (e.g. torch's current stream) for all device work. */
AQH_EXPORT int aqh_set_stream(AqhHider* h, void* cuda_stream);

AQH_EXPORT int aqh_frame_params_default(AqhFrameParams* p);   /* options.cpp:273-305 defaults */
/* Fill the DoF members from RiDepthOfField values the way SetDepthOfFieldData does. */
AQH_EXPORT int aqh_frame_params_set_dof(AqhFrameParams* p, float fstop, float focallength, float focaldistance,
                                        float scale_x, float scale_y);
/* "rgba"/"rgb"/"a"/"z"/"rgbaz" -> channel list in the core's a,r,g,b,z request order (ddmanager.cpp:455-480)
 * or, with driver_order != 0, the r,g,b,a,z order the file/tiff driver negotiates (display.cpp:454-490). */
AQH_EXPORT int aqh_display_from_mode(AqhDisplayDesc* d, const char* mode, int driver_order,
                                     float one, float min, float max, float dither);

AQH_EXPORT int aqh_begin_frame(AqhHider* h, const AqhFrameParams* p);
/* The frame's CSG tree (RiSolidBegin/End -> CqCSGTreeNode, libs/core/csgtree.h:69-215), between aqh_begin_frame and the
 * first CSG grid: node i has type[i] (AQH_CSG_*) and parent[i] (-1 for a root; several trees may coexist).  The children
 * of a node are ordered by node index, which is the order RiSolidBegin created them in.  Grids name their primitive
 * node in csg_node.  Resolved per sample exactly like CqCSGTreeNode::ProcessTree (csgtree.cpp:144-351). */
enum { AQH_CSG_PRIMITIVE = 0, AQH_CSG_UNION = 1, AQH_CSG_INTERSECTION = 2, AQH_CSG_DIFFERENCE = 3 };
AQH_EXPORT int aqh_set_csg_tree(AqhHider* h, int n_nodes, const int32_t* type, const int32_t* parent);

/* Trim curves (RiTrimCurve) as the hider sees them: per trimmed surface a set of closed loops, every loop the polyline of
 * (u, v) points CqTrimLoop::Prepare leaves in m_aCurvePoints (geometry/trimcurve.cpp:109-135; evaluating the NURBS
 * curves is the front end's work).  set s owns the loops [set_first_loop[s], set_first_loop[s+1]), loop l the points
 * [loop_first_point[l], loop_first_point[l+1]) of `points` (2 floats each).  Call inside a frame before the first grid
 * with a trim set.  Replaces CqSurface::bIsPointTrimmed / bIsLineIntersecting (geometry/nurbs.h:350-357,
 * CqTrimLoopArray::TrimPoint / LineIntersects, geometry/trimcurve.cpp:145-242): the device drops the micropolygons that
 * are trimmed away entirely while busting (micropolygon.cpp:784-835) and tests the hits of the ones the curves cross
 * (micropolygon.cpp:1594-1628). */
AQH_EXPORT int aqh_set_trim_loops(AqhHider* h, int n_sets, const int32_t* set_first_loop, const int32_t* loop_first_point,
                                  const float* points);
/* Floats per pixel of the channel buffer of the current frame: AQH_NUM_CHANNELS + the AOV floats. */
AQH_EXPORT int aqh_channel_count(const AqhHider* h, int* n);
AQH_EXPORT int aqh_add_grid(AqhHider* h, const AqhGridDesc* g);
AQH_EXPORT int aqh_add_grid_block(AqhHider* h, const AqhGridBlock* b);
/* Hide + filter + expose + quantise everything added since begin_frame, download, fire callbacks
 * in reference bucket order. */
AQH_EXPORT int aqh_end_frame(AqhHider* h, const AqhCallbacks* cb);
/* The same device work without download or callbacks: results stay in HBM (see aqh_device_*). */
AQH_EXPORT int aqh_render_device(AqhHider* h);

/* Occlusion feedback to the front end -- what CqOcclusionTree::canCull(const CqBound&) gives aqsis
 * (libs/core/occlusion.h:128, occlusion.cpp:161-225; call site CqBucketProcessor::RenderSurface,
 * bucketprocessor.cpp:945-958: a surface whose raster bound lies behind every sample it could touch is
 * re-posted / dropped before it is diced and shaded).
 * aqh_flush hides the grids submitted so far (the frame stays open: more grids may follow, aqh_end_frame
 * renders all of them) and keeps, per image pixel, the farthest occlusion depth over the pixel's samples --
 * FLT_MAX while any sample of the pixel is uncovered.  aqh_can_cull(bound = xmin, ymin, zmin, xmax, ymax, zmax in
 * raster space, the bound aqsis passes) sets *culled = 1 when every pixel the bound touches (inside the crop
 * window) is occluded nearer than zmin: pixel granularity instead of the reference's per-sample tree, i.e. never
 * culls more than the reference would.  Before the first flush of a frame nothing is culled.  As in the reference
 * nothing is culled when the display mode has z and the depth filter is max or average. */
AQH_EXPORT int aqh_flush(AqhHider* h);
AQH_EXPORT int aqh_can_cull(const AqhHider* h, const float bound[6], int* culled);
AQH_EXPORT int aqh_frame_stats(const AqhHider* h, AqhFrameStats* out);

/* Forget everything cached across frames (the replayed random stream and the tables built from the frame options):
 * the next aqh_begin_frame pays the full Prepare_bucket cost again, like the first frame of a process. */
AQH_EXPORT int aqh_clear_caches(AqhHider* h);

/* A minimal in-process capture display (what aqsis' debugdd / a framebuffer driver does with DspyImageData): ready-made
 * AqhCallbacks targets that copy every bucket into caller-owned full-frame images.  user = AqhCapture*. */
typedef struct AqhCapture
{
	int32_t xres, yres, n_channels;  /* n_channels: floats per pixel of `channels` (aqh_channel_count) */
	float* channels;                 /* xres*yres*n_channels floats or NULL */
	unsigned char* display[AQH_MAX_DISPLAYS];   /* xres*yres*entrysize bytes each or NULL */
	int64_t buckets, bytes;          /* counted by the callbacks */
} AqhCapture;
AQH_EXPORT int aqh_capture_on_bucket(void* user, int xmin, int xmax1, int ymin, int ymax1,
                                     const float* channels, int row_stride_floats, int pixel_stride_floats);
AQH_EXPORT int aqh_capture_on_data(void* user, int display, int xmin, int xmax1, int ymin, int ymax1,
                                   int entrysize, const unsigned char* data);

/* Results of the last frame.  Host copies (valid after aqh_end_frame): full-resolution images,
 * rows owned by other ranks are zero. */
AQH_EXPORT int aqh_image_channels(const AqhHider* h, const float** data, int* width, int* height);
AQH_EXPORT int aqh_image_display(const AqhHider* h, int display, const unsigned char** data,
                                 int* entrysize, int* type);
/* Device copies (valid after aqh_render_device / aqh_end_frame) for NCCL gathers by the caller. */
AQH_EXPORT int aqh_device_channels(const AqhHider* h, void** dev_ptr, size_t* bytes);
AQH_EXPORT int aqh_device_display(const AqhHider* h, int display, void** dev_ptr, size_t* bytes);
/* Pixel rows [y0,y1) of strip i owned by this rank. */
/* Device-less: the strips (pixel-row ranges) `rank` of p->world_size owns, round-robin in strips of
 * p->strip_rows rows rounded down to a multiple of 16, or (strip_rows <= 0, the default) world*k near-equal
 * strips with k = max(1, rows/(64*world)) so every rank owns k of them.  Writes min(*n_strips, capacity) entries. */
AQH_EXPORT int aqh_strip_layout(const AqhFrameParams* p, int rank, int* n_strips, int* y0, int* y1, int capacity);
AQH_EXPORT int aqh_num_strips(const AqhHider* h, int* n);
AQH_EXPORT int aqh_strip(const AqhHider* h, int i, int* y0, int* y1);

/* --- sharding one frame over the GPUs of a box, behind the C ABI so that a C++ host needs nothing else.
 * Device-less.  aqh_grid_rank_masks: bit r of rank_mask[g] is set when rank r must receive grid g of a HOST-memory
 * block, i.e. when the grid's raster row range -- union of its motion keys, grown by the largest circle of confusion
 * and the filter half-width like CqImageBuffer::AddMPG does per micropolygon (imagebuffer.cpp:519-554); camera-space
 * grids are projected with p->cam_to_raster first -- touches a strip of rank r.  Straddlers get several bits: they are
 * replicated.  p->world_size <= AQH_MAX_RANKS. */
AQH_EXPORT int aqh_grid_rank_masks(const AqhFrameParams* p, const AqhGridBlock* b, uint64_t* rank_mask /* n_grids */);
/* Work estimate per pixel row: adds, for every grid of the block, its micropolygon count spread over the rows it
 * covers to row_cost[p->yres] (call once per block; the caller zeroes the array first). */
AQH_EXPORT int aqh_grid_row_cost(const AqhFrameParams* p, const AqhGridBlock* b, double* row_cost /* yres */);
/* Contiguous strips of equal estimated work: fills p->strip_bounds[0..world_size] (multiples of 16 rows for tall strips, of 4
 * rows otherwise, except at the ends of the crop window) and sets p->strip_rows = -2. */
AQH_EXPORT int aqh_balance_strips(AqhFrameParams* p, const double* row_cost /* yres */);

/* The one collective of the path: the finished strips of every rank travel to `root` over NVLink (grouped
 * ncclSend/ncclRecv straight from / into the device images, no staging, no packing: a strip is a contiguous range of
 * rows).  NCCL is loaded at run time (libnccl.so.2); without it these return AQH_ERR_UNSUPPORTED.
 *   aqh_comm_unique_id   rank 0 makes the 128-byte ncclUniqueId; the host carries it to the other ranks by its own means
 *   aqh_comm_init        every rank joins (collective call)
 *   aqh_gather           after aqh_render_device: stream-ordered on the hider's stream; on return from the following
 *                        synchronisation root's device images hold the whole frame.
 * With a communicator set, aqh_end_frame gathers before it downloads: root then owns the complete host images and
 * fires the callbacks; the other ranks download nothing. */
AQH_EXPORT int aqh_comm_unique_id(void* id128);
AQH_EXPORT int aqh_comm_init(AqhHider* h, const void* id128, int rank, int world_size);
AQH_EXPORT int aqh_comm_destroy(AqhHider* h);
AQH_EXPORT int aqh_gather(AqhHider* h, int root);

/* Host-side pieces of the path, exported so the reference-side tests can pin them. */
AQH_EXPORT float aqh_box_filter(float x, float y, float xw, float yw);
AQH_EXPORT float aqh_triangle_filter(float x, float y, float xw, float yw);
AQH_EXPORT float aqh_gaussian_filter(float x, float y, float xw, float yw);
AQH_EXPORT float aqh_catmullrom_filter(float x, float y, float xw, float yw);
AQH_EXPORT float aqh_sinc_filter(float x, float y, float xw, float yw);
AQH_EXPORT float aqh_mitchell_filter(float x, float y, float xw, float yw);
AQH_EXPORT float aqh_disk_filter(float x, float y, float xw, float yw);
AQH_EXPORT float aqh_bessel_filter(float x, float y, float xw, float yw);
AQH_EXPORT AqhFilterFunc aqh_filter_by_name(const char* name);

/* CqRandom (libs/math/random.cpp:97-224): one MT19937 stream per handle-less state. */
typedef struct AqhRandom AqhRandom;
AQH_EXPORT AqhRandom* aqh_random_create(uint32_t seed);
AQH_EXPORT void aqh_random_destroy(AqhRandom* r);
AQH_EXPORT void aqh_random_reseed(AqhRandom* r, uint32_t seed);
AQH_EXPORT uint32_t aqh_random_uint(AqhRandom* r);
AQH_EXPORT float aqh_random_float(AqhRandom* r);
AQH_EXPORT uint32_t aqh_random_int(AqhRandom* r, uint32_t range);

/* Sampler tables (CqMultiJitteredSampler / CqGridSampler) built from a given RNG state:
 * out arrays hold ncache*n entries; positions are x,y pairs.  ncache is 250 (jitter) or 1 (grid). */
AQH_EXPORT int aqh_sampler_tables(AqhRandom* r, int xsamples, int ysamples, int jitter,
                                  float* positions_xy, float* values_1d, int32_t* shuffled, int* ncache);
/* The per-pixel RNG replay of a whole frame: for every pixel of the sample
 * region [crop-shift, crop+shift) five pattern indices (shuffle, position, dof, time, lod) as
 * uint8 planes of size sw*sh, plus one dither float per display per image pixel. */
AQH_EXPORT int aqh_replay_frame_rng(const AqhFrameParams* p, uint8_t* pattern_planes /*5*sw*sh*/,
                                    float* dither /*n_displays*xres*yres, may be NULL*/,
                                    int* sx0, int* sy0, int* sw, int* sh);
/* Filter weight table exactly as InitialiseFilterValues lays it out. */
AQH_EXPORT int aqh_filter_table(const AqhFrameParams* p, float* table, int* n_entries);

#ifdef __cplusplus
}
#endif
#endif /* AQSIS_B200_HIDER_H_INCLUDED */
