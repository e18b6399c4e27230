// host_sampling.h -- host half of the hider boundary: the renderer-global random stream,
// the 250 cached jitter patterns and the per-pixel replay of that stream in bucket order.
//
// The reference draws everything from ONE process-global MT19937
// (libs/math/random.cpp:97-224).  The device kernels are bucket-order independent, so
// the only way to reproduce the reference's sample pattern is to replay its draw order
// here and hand the kernels per-pixel pattern indices.
#ifndef AQSIS_B200_HOST_SAMPLING_H
#define AQSIS_B200_HOST_SAMPLING_H

#include <cstdint>
#include <vector>
#include "../../include/aqsis_b200_hider.h"

namespace aqh {

/// MT19937 with the reference's float conversion (random.cpp:194-214).
class Random
{
public:
	explicit Random(uint32_t seed = 5489u) { reseed(seed); }
	void reseed(uint32_t seed);
	uint32_t nextUint()
	{
		if(m_idx >= N)
			refill();
		uint32_t y = m_state[m_idx++];
		y ^= (y >> 11);
		y ^= (y << 7) & 0x9d2c5680u;
		y ^= (y << 15) & 0xefc60000u;
		y ^= (y >> 18);
		return y;
	}
	/// uint32 * (1/(2^32+128)) in double, rounded to float: always < 1.
	float nextFloat() { return static_cast<float>(nextUint() * (1.0 / 4294967424.0)); }
	/// RandomInt(range) = lfloor(double(float(range) * RandomFloat()))
	uint32_t nextInt(uint32_t range)
	{
		float f = static_cast<float>(range) * nextFloat();
		return static_cast<uint32_t>(static_cast<long>(f)); // f >= 0, so lfloor == truncation
	}
	void discard(uint64_t n) { while(n--) nextUint(); }
private:
	enum { N = 624, M = 397 };
	void refill();
	uint32_t m_state[N];
	int m_idx;
};

/// The sampler tables of IqSampler (isampler.h:42-78): `ncache` patterns of n samples.
struct SamplerTables
{
	int xs = 0, ys = 0, n = 0, ncache = 0;
	std::vector<float> pos;       // ncache*n*2 (x,y)
	std::vector<float> val1d;     // ncache*n
	std::vector<int32_t> shuffled; // ncache*n
};

/// CqMultiJitteredSampler ctor (multijitter.h:90-101) incl. the trailing Reseed(19),
/// or CqGridSampler (grid.cpp:37-63).  Note the jittered sampler is ALWAYS constructed
/// by RenderImage (imagebuffer.cpp:694) and so always consumes the stream.
void buildJitterTables(Random& rng, int xs, int ys, SamplerTables& out);
void buildGridTables(int xs, int ys, SamplerTables& out);

/// Geometry of the replay: the global sample region and bucket walk.
struct ReplayLayout
{
	int shiftX, shiftY;           // m_DiscreteShiftX/Y = lfloor(filterwidth/2)
	int sx0, sy0, sw, sh;         // sample region [crop-shift, crop+shift)
	int bx0, by0, bx1, by1;       // m_bucketRegion (bucket index ranges, max exclusive)
};
ReplayLayout replayLayout(const AqhFrameParams& p);

/// Per-pixel pattern indices in reference draw order.  planes: 5 planes (shuffle, position,
/// dof, time, lod) of sw*sh bytes; dither: n_displays planes of xres*yres floats (may be null).
/// `rng` must be in the state the reference has when RenderImage() starts its bucket loop.
void replayFrame(const AqhFrameParams& p, const ReplayLayout& L, Random& rng, bool jitter,
                 uint8_t* planes, float* dither);

/// CqBucketProcessor::InitialiseFilterValues (bucketprocessor.cpp:811-856).
void buildFilterTable(const AqhFrameParams& p, std::vector<float>& table);

/// CqBucketProcessor::CalculateDofBounds (bucketprocessor.cpp:858-913): n entries of
/// (minx, miny, maxx, maxy) in lens space.
void buildDofBounds(int xs, int ys, std::vector<float>& bounds);

/// CqImagePixel::projectToCircle (imagepixel.h:429-436).
void projectToCircle(float x, float y, float& ox, float& oy);

} // namespace aqh
#endif
