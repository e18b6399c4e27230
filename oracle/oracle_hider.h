/* oracle_hider.h -- C interface of the CPU oracle.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load liboracle_hider.so; the product never does.
 * The oracle consumes the SAME parameter and grid structs as the product boundary
 * (include/aqsis_b200_hider.h) so that parity tests feed both sides identical inputs.
 */
#ifndef ORACLE_HIDER_H_INCLUDED
#define ORACLE_HIDER_H_INCLUDED

#include "aqsis_b200_hider.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct OrcStats
{
	double prepare_s, bust_s, render_s, combine_s, filter_s, display_s, total_s;
	int64_t n_micropolygons;   /* MPs that survived AddMPG's crop reject */
	int64_t n_bucket_entries;  /* MP references over all bucket lists */
	int64_t n_samples;         /* sample points set up */
	int64_t spl_count, spl_bound_hits, spl_hits;  /* CqStats SPL_* counters */
	int64_t n_deep_hits;
	int32_t threads;
} OrcStats;

/* Render one frame.  grids->memory_space must be 0.  channels: xres*yres*(9 + AOV floats) floats (may be NULL);
 * display_out[d]: xres*yres*entrysize(d) bytes (entries may be NULL).  nthreads <= 1 follows the
 * reference's single-threaded bucket loop literally; > 1 distributes buckets over threads
 * (identical results: pixels are owned by exactly one bucket). */
int orc_render(const AqhFrameParams* p, const AqhGridBlock* grids, float* channels,
               unsigned char* const* display_out, int nthreads, OrcStats* stats);

int orc_display_entrysize(const AqhDisplayDesc* d, int* type_out);

/* leaves, for pinning against oracle/_ref */
void orc_random_reseed(uint32_t seed);
uint32_t orc_random_uint(void);
float orc_random_float(void);
uint32_t orc_random_int(uint32_t range);
/* builds tables from the current oracle RNG state (consumes the stream, reseeds 19 when jitter) */
int orc_sampler_tables(int xs, int ys, int jitter, float* pos_xy, float* val1d, int32_t* shuffled);
float orc_filter(int which, float x, float y, float xw, float yw);
int orc_set_csg_tree(int n_nodes, const int32_t* type, const int32_t* parent);   /* CSG tree of the following orc_render calls; 0 nodes = none */
int orc_set_trim_loops(int n_sets, const int32_t* set_first_loop, const int32_t* loop_first_point, const float* points);   /* trim loops of the following orc_render calls; 0 sets = none */
int orc_trim_point(int set, float x, float y);                       /* CqTrimLoopArray::TrimPoint of set (1-based) */
int orc_trim_line(int set, float x1, float y1, float x2, float y2);  /* CqTrimLoopArray::LineIntersects */
void orc_set_filter(int which);   /* pixel filter of the following orc_render calls by index (see oracle_hider.cpp); < 0 = by filter_func */
void orc_invbilinear(const float* verts8, float px, float py, float* uv);
float orc_bilerp(float a, float b, float c, float d, float u, float v);
int orc_filter_table(const AqhFrameParams* p, float* table);
int orc_replay(const AqhFrameParams* p, uint8_t* planes, float* dither, int* sx0, int* sy0, int* sw, int* sh);
void orc_dof_bounds(int xs, int ys, float* bounds4);

#ifdef __cplusplus
}
#endif
#endif
