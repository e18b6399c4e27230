"""ctypes mirror of include/aqsis_b200_hider.h (the C ABI of the hider).

Keep this file in lock-step with the header: tests/test_abi.py checks struct sizes and
that the shared library exports every symbol the header declares.
"""
import ctypes as C

AQH_ABI_VERSION = 3

# AqhStatus
AQH_OK = 0
AQH_ERR_BAD_PARAMS = 1
AQH_ERR_NO_MEMORY = 2
AQH_ERR_UNSUPPORTED = 3
AQH_ERR_NO_DEVICE = 4
AQH_ERR_CUDA = 5
AQH_ERR_STATE = 6
AQH_ERR_DEEP_OVERFLOW = 7
AQH_ERR_CALLBACK = 8
STATUS_NAMES = {0: "AQH_OK", 1: "AQH_ERR_BAD_PARAMS", 2: "AQH_ERR_NO_MEMORY", 3: "AQH_ERR_UNSUPPORTED",
                4: "AQH_ERR_NO_DEVICE", 5: "AQH_ERR_CUDA", 6: "AQH_ERR_STATE", 7: "AQH_ERR_DEEP_OVERFLOW",
                8: "AQH_ERR_CALLBACK"}

DEPTHFILTER_MIN, DEPTHFILTER_MIDPOINT, DEPTHFILTER_MAX, DEPTHFILTER_AVERAGE = 0, 1, 2, 3
DMODE_RGB, DMODE_A, DMODE_Z = 1, 2, 4
FLOAT32, UNSIGNED32, SIGNED32, UNSIGNED16, SIGNED16, UNSIGNED8, SIGNED8 = 1, 2, 3, 4, 5, 6, 7
CH_CI_R, CH_CI_G, CH_CI_B, CH_OI_R, CH_OI_G, CH_OI_B, CH_ALPHA, CH_Z, CH_COVERAGE = range(9)
NUM_CHANNELS = 9

GRID_SMOOTH = 1 << 0
GRID_MATTE = 1 << 1
GRID_MATTE_ALPHA = 1 << 2
GRID_TRIANGULAR = 1 << 3
GRID_CAMERA_SPACE = 1 << 4
GRID_USES_CSG = 1 << 5
GRID_POINTS = 1 << 6
GRID_CULL_BACKFACING = 1 << 7
GRID_CULL_TRANSPARENT = 1 << 8
GRID_TRIM_OUTSIDE = 1 << 9
CSG_PRIMITIVE, CSG_UNION, CSG_INTERSECTION, CSG_DIFFERENCE = 0, 1, 2, 3
DISPLAY_SCANLINE_ORDER = 1
MAX_RANKS, MAX_AOVS, MAX_AOV_FLOATS = 64, 8, 21

MAX_DISPLAYS = 8
FILTER_REFERENCE_ORDER, FILTER_TILE_PARTIALS = 0, 1
MAX_DISPLAY_CHANNELS = 16

FilterFunc = C.CFUNCTYPE(C.c_float, C.c_float, C.c_float, C.c_float, C.c_float)
BucketFunc = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.c_int, C.c_int)
ImagerFunc = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.c_int, C.c_int)
DataFunc = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                       C.POINTER(C.c_ubyte))
ProgressFunc = C.CFUNCTYPE(None, C.c_void_p, C.c_float)


class DisplayDesc(C.Structure):
    _fields_ = [
        ("n_channels", C.c_int32),
        ("channel", C.c_int32 * MAX_DISPLAY_CHANNELS),
        ("type", C.c_int32),
        ("quantize_zero", C.c_float),
        ("quantize_one", C.c_float),
        ("quantize_min", C.c_float),
        ("quantize_max", C.c_float),
        ("quantize_dither", C.c_float),
        ("flags", C.c_int32),
    ]


class AovDesc(C.Structure):
    _fields_ = [("name", C.c_char * 32), ("n_floats", C.c_int32), ("reserved", C.c_int32)]


class FrameParams(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32),
        ("xres", C.c_int32), ("yres", C.c_int32),
        ("crop_xmin", C.c_int32), ("crop_xmax", C.c_int32), ("crop_ymin", C.c_int32), ("crop_ymax", C.c_int32),
        ("xsamples", C.c_int32), ("ysamples", C.c_int32),
        ("filter_xwidth", C.c_float), ("filter_ywidth", C.c_float),
        ("filter_func", C.c_void_p),
        ("bucket_xsize", C.c_int32), ("bucket_ysize", C.c_int32),
        ("clip_near", C.c_float), ("clip_far", C.c_float),
        ("shutter_open", C.c_float), ("shutter_close", C.c_float),
        ("use_dof", C.c_int32),
        ("dof_multiplier", C.c_float), ("dof_one_over_focal_distance", C.c_float),
        ("dof_scale_x", C.c_float), ("dof_scale_y", C.c_float),
        ("depth_filter", C.c_int32),
        ("zthreshold", C.c_float * 3),
        ("display_mode", C.c_int32),
        ("exposure_gain", C.c_float), ("exposure_gamma", C.c_float),
        ("jitter", C.c_int32),
        ("cam_to_raster", C.c_float * 16),
        ("rng_seed", C.c_uint32), ("rng_predraws", C.c_uint32),
        ("n_displays", C.c_int32),
        ("display", DisplayDesc * MAX_DISPLAYS),
        ("n_aovs", C.c_int32),
        ("aov", AovDesc * MAX_AOVS),
        ("rank", C.c_int32), ("world_size", C.c_int32),
        ("strip_rows", C.c_int32),
        ("strip_bounds", C.c_int32 * (MAX_RANKS + 1)),
        ("deep_hits_per_sample", C.c_int32),
        ("filter_mode", C.c_int32),
        ("plane_budget_mb", C.c_int32),
        ("reserved", C.c_int32 * 6),
    ]

    @property
    def aov_floats(self):
        return sum(self.aov[a].n_floats for a in range(self.n_aovs))


class GridDesc(C.Structure):
    _fields_ = [
        ("cu", C.c_int32), ("cv", C.c_int32),
        ("nkeys", C.c_int32),
        ("key_times", C.POINTER(C.c_float)),
        ("P", C.POINTER(C.POINTER(C.c_float))),
        ("Ci", C.POINTER(C.c_float)),
        ("Oi", C.POINTER(C.c_float)),
        ("culled", C.POINTER(C.c_uint8)),
        ("flags", C.c_uint32),
        ("lod_bounds", C.c_float * 2),
        ("aov", C.POINTER(C.c_float)),
        ("Ng", C.POINTER(C.c_float)),
        ("N", C.POINTER(C.c_float)),
        ("radius", C.POINTER(C.c_float)),
        ("csg_node", C.c_int32),
        ("trim_set", C.c_int32),
        ("trim_uv", C.POINTER(C.c_float)),
    ]


class GridBlock(C.Structure):
    _fields_ = [
        ("n_grids", C.c_int64),
        ("cu", C.c_void_p), ("cv", C.c_void_p), ("nkeys", C.c_void_p), ("flags", C.c_void_p),
        ("lod_bounds", C.c_void_p), ("key_times", C.c_void_p),
        ("P", C.c_void_p), ("Ci", C.c_void_p), ("Oi", C.c_void_p), ("culled", C.c_void_p),
        ("memory_space", C.c_int32),
        ("reserved", C.c_int32 * 3),
        ("aov", C.c_void_p), ("Ng", C.c_void_p), ("N", C.c_void_p), ("radius", C.c_void_p), ("csg_node", C.c_void_p),
        ("trim_set", C.c_void_p), ("trim_uv", C.c_void_p),
    ]


class Callbacks(C.Structure):
    _fields_ = [
        ("user", C.c_void_p),
        ("on_bucket", BucketFunc),
        ("on_data", DataFunc),
        ("on_progress", ProgressFunc),
        ("on_imager", ImagerFunc),
    ]


class Capture(C.Structure):
    _fields_ = [("xres", C.c_int32), ("yres", C.c_int32), ("n_channels", C.c_int32), ("channels", C.c_void_p),
                ("display", C.c_void_p * MAX_DISPLAYS), ("buckets", C.c_int64), ("bytes", C.c_int64)]


class FrameStats(C.Structure):
    _fields_ = [
        ("prepare_ms", C.c_double), ("upload_ms", C.c_double), ("project_bust_ms", C.c_double),
        ("render_mpgs_ms", C.c_double), ("filter_ms", C.c_double), ("display_ms", C.c_double),
        ("download_ms", C.c_double), ("device_total_ms", C.c_double),
        ("n_grids", C.c_int64), ("n_vertices", C.c_int64), ("n_micropolygons", C.c_int64),
        ("n_bin_entries", C.c_int64), ("n_samples", C.c_int64), ("n_deep_hits", C.c_int64),
        ("gpu_launches", C.c_int64), ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
        ("device_bytes", C.c_int64), ("n_bands", C.c_int64), ("gather_ms", C.c_double),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


TYPE_SIZES = {FLOAT32: 4, UNSIGNED32: 4, SIGNED32: 4, UNSIGNED16: 2, SIGNED16: 2, UNSIGNED8: 1, SIGNED8: 1}
TYPE_NUMPY = {FLOAT32: "float32", UNSIGNED32: "uint32", SIGNED32: "int32", UNSIGNED16: "uint16",
              SIGNED16: "int16", UNSIGNED8: "uint8", SIGNED8: "int8"}
