// host_sampling.cpp -- see host_sampling.h.
#include "host_sampling.h"

#include <algorithm>
#include <cmath>

namespace aqh {

// ---------------------------------------------------------------------------
// MT19937 -- init_genrand / genrand_int32, libs/math/random.cpp:102-160.
void Random::reseed(uint32_t seed)
{
	m_state[0] = seed;
	for(int i = 1; i < N; ++i)
		m_state[i] = 1812433253u * (m_state[i-1] ^ (m_state[i-1] >> 30)) + static_cast<uint32_t>(i);
	m_idx = N;
}

void Random::refill()
{
	const uint32_t upper = 0x80000000u, lower = 0x7fffffffu, matrixA = 0x9908b0dfu;
	int k = 0;
	for(; k < N - M; ++k)
	{
		uint32_t y = (m_state[k] & upper) | (m_state[k+1] & lower);
		m_state[k] = m_state[k+M] ^ (y >> 1) ^ ((y & 1u) ? matrixA : 0u);
	}
	for(; k < N - 1; ++k)
	{
		uint32_t y = (m_state[k] & upper) | (m_state[k+1] & lower);
		m_state[k] = m_state[k+(M-N)] ^ (y >> 1) ^ ((y & 1u) ? matrixA : 0u);
	}
	uint32_t y = (m_state[N-1] & upper) | (m_state[0] & lower);
	m_state[N-1] = m_state[M-1] ^ (y >> 1) ^ ((y & 1u) ? matrixA : 0u);
	m_idx = 0;
}

// ---------------------------------------------------------------------------
// One cached pattern of n = xs*ys samples: multi-jittered 2-D positions, a stratified 1-D value per sample and a
// permutation of the sample indices.  The ARITHMETIC and the order of the random draws are fixed by the reference
// (CqMultiJitteredSampler, libs/core/multijitter.cpp:83-202) -- the stream is shared with everything else the renderer
// draws, so one draw out of place changes every later pattern; the known-answer tests pin both.

// Descending Fisher-Yates over `count` elements `stride` apart: for k = count .. 2 draw j in [0, k) and exchange
// elements k-1 and j.  This is the one shuffle the reference uses everywhere (:104-127, :195-201).
template<class T> static void shuffleDown(Random& rng, T* first, int count, int stride)
{
	for(int k = count; k > 1; --k)
	{
		const int j = static_cast<int>(rng.nextInt(static_cast<uint32_t>(k)));
		std::swap(first[size_t(k - 1)*stride], first[size_t(j)*stride]);
	}
}

static void jitterPattern(Random& rng, int xs, int ys, float* pos, float* val1d, int32_t* shuffled)
{
	const int n = xs*ys;
	if(n == 1)
	{
		// a 2-D point takes two draws, y first (the reference builds CqVector2D(RandomFloat(), RandomFloat()) and g++
		// evaluates arguments right to left), then one draw for the 1-D value that the code below overwrites
		const float y = rng.nextFloat(), x = rng.nextFloat();
		pos[0] = x; pos[1] = y;
		(void)rng.nextFloat();
	}
	else
	{
		// Sub-pixel (ix, iy) starts in sub-cell (iy, ix) of its own sub-pixel -- the canonical multi-jitter layout where
		// the n x n fine grid has exactly one sample per fine row and per fine column.  Shuffling the fine-x cells among
		// the sub-pixels of a row, then the fine-y cells among those of a column, keeps that property.
		std::vector<int> fineX(n), fineY(n);
		for(int iy = 0; iy < ys; ++iy)
			for(int ix = 0; ix < xs; ++ix) { fineX[iy*xs + ix] = iy; fineY[iy*xs + ix] = ix; }
		for(int iy = 0; iy < ys; ++iy) shuffleDown(rng, &fineY[iy*xs], xs, 1);      // along each row of sub-pixels
		for(int ix = 0; ix < xs; ++ix) shuffleDown(rng, &fineX[ix], ys, xs);         // along each column
		const float cellH = 1.0f / ys, cellW = 1.0f / xs, fine = 1.0f / n;
		for(int i = 0; i < n; ++i)
		{
			const float jy = rng.nextFloat(), jx = rng.nextFloat();                  // y before x, as above
			pos[2*i]     = (fineX[i] + jx)*fine + (i % xs)*cellW;
			pos[2*i + 1] = (fineY[i] + jy)*fine + (i / xs)*cellH;
		}
	}
	// stratified 1-D values (times, levels of detail): i/n plus ONE shared offset in [0, 1/n); the running sum is the
	// reference's (:174-190), not i*step
	const float step = 1.0f / n;
	const float offset = step * rng.nextFloat();
	float base = 0;
	for(int i = 0; i < n; ++i) { val1d[i] = base + offset; base += step; }
	// permutation of the sample indices (depth-of-field offsets are handed out through it)
	for(int i = 0; i < n; ++i) shuffled[i] = i;
	shuffleDown(rng, shuffled, n, 1);
}

void buildJitterTables(Random& rng, int xs, int ys, SamplerTables& out)
{
	out.xs = xs; out.ys = ys; out.n = xs*ys; out.ncache = 250;   // m_cacheSize, multijitter.h:60
	const int n = out.n;
	out.pos.assign(size_t(250)*n*2, 0.f);
	out.val1d.assign(size_t(250)*n, 0.f);
	out.shuffled.assign(size_t(250)*n, 0);
	for(int i = 0; i < 250; ++i)
		jitterPattern(rng, xs, ys, &out.pos[size_t(i)*n*2], &out.val1d[size_t(i)*n], &out.shuffled[size_t(i)*n]);
	rng.reseed(19);                                              // multijitter.h:100
}

// CqGridSampler::setupGridPattern, libs/core/grid.cpp:37-63.
void buildGridTables(int xs, int ys, SamplerTables& out)
{
	out.xs = xs; out.ys = ys; out.n = xs*ys; out.ncache = 1;
	const int n = out.n;
	out.pos.assign(size_t(n)*2, 0.f);
	out.val1d.assign(n, 0.f);
	out.shuffled.assign(n, 0);
	const float xScale = static_cast<float>(1.0/xs);
	const float yScale = static_cast<float>(1.0/ys);
	for(int j = 0; j < ys; ++j)
		for(int i = 0; i < xs; ++i)
		{
			out.pos[2*(j*xs+i)]   = static_cast<float>(xScale*(i+0.5));
			out.pos[2*(j*xs+i)+1] = static_cast<float>(yScale*(j+0.5));
		}
	// "TqFloat dt = 1/nSamples" is an integer division in the reference (grid.cpp:52):
	// every time/lod sample is 0 unless n == 1.
	const float dt = static_cast<float>(1/n);
	float sample = static_cast<float>(dt*0.5);
	for(int i = 0; i < n; ++i)
	{
		out.val1d[i] = sample;
		sample += dt;
	}
	for(int i = 0; i < n; ++i)
		out.shuffled[i] = i;
}

// ---------------------------------------------------------------------------
static inline long lfloorf_(float x) { long i = static_cast<long>(x); return i - (x < 0 && x != static_cast<float>(i)); }
static inline long lceilf_(float x) { long i = static_cast<long>(x); return i + (x > 0 && x != static_cast<float>(i)); }

ReplayLayout replayLayout(const AqhFrameParams& p)
{
	ReplayLayout L;
	L.shiftX = static_cast<int>(lfloorf_(p.filter_xwidth/2.0f));   // bucketprocessor.cpp:36-37
	L.shiftY = static_cast<int>(lfloorf_(p.filter_ywidth/2.0f));
	L.sx0 = p.crop_xmin - L.shiftX;
	L.sy0 = p.crop_ymin - L.shiftY;
	L.sw = p.crop_xmax + L.shiftX - L.sx0;
	L.sh = p.crop_ymax + L.shiftY - L.sy0;
	// m_bucketRegion, imagebuffer.cpp:191-195
	L.bx0 = p.crop_xmin / p.bucket_xsize;
	L.by0 = p.crop_ymin / p.bucket_ysize;
	L.bx1 = (p.crop_xmax - 1) / p.bucket_xsize + 1;
	L.by1 = (p.crop_ymax - 1) / p.bucket_ysize + 1;
	return L;
}

void replayFrame(const AqhFrameParams& p, const ReplayLayout& L, Random& rng, bool jitter,
                 uint8_t* planes, float* dither)
{
	const size_t plane = size_t(L.sw)*L.sh;
	if(!jitter)
		std::fill(planes, planes + 5*plane, uint8_t(0));
	for(int row = L.by0; row < L.by1; ++row)
	{
		const int yPos = row*p.bucket_ysize;
		const int ySize = std::min(p.bucket_ysize, p.yres - yPos);
		for(int col = L.bx0; col < L.bx1; ++col)
		{
			const int xPos = col*p.bucket_xsize;
			const int xSize = std::min(p.bucket_xsize, p.xres - xPos);
			// CqBucketProcessor::preProcess, bucketprocessor.cpp:112-137.  With row-major
			// bucket order the left/top cache segments exist for every bucket after the
			// first column/row of m_bucketRegion; right/bottom never do.
			int sminx = xPos - L.shiftX, sminy = yPos - L.shiftY;
			int smaxx = xPos + xSize + L.shiftX, smaxy = yPos + ySize + L.shiftY;
			sminx = std::max(sminx, p.crop_xmin - L.shiftX);
			sminy = std::max(sminy, p.crop_ymin - L.shiftY);
			smaxx = std::min(smaxx, p.crop_xmax + L.shiftX);
			smaxy = std::min(smaxy, p.crop_ymax + L.shiftY);
			if(col > L.bx0) sminx += 2*L.shiftX;
			if(row > L.by0) sminy += 2*L.shiftY;
			if(jitter)
			{
				for(int y = sminy; y < smaxy; ++y)
				{
					uint8_t* q = planes + size_t(y - L.sy0)*L.sw - L.sx0;
					for(int x = sminx; x < smaxx; ++x)
					{
						// CqImagePixel::setSamples, imagepixel.cpp:338-347
						q[x]           = static_cast<uint8_t>(rng.nextInt(250)); // getShuffledIndices
						q[x + plane]   = static_cast<uint8_t>(rng.nextInt(250)); // positions
						q[x + 2*plane] = static_cast<uint8_t>(rng.nextInt(250)); // dofOffsets
						q[x + 3*plane] = static_cast<uint8_t>(rng.nextInt(250)); // times
						q[x + 4*plane] = static_cast<uint8_t>(rng.nextInt(250)); // lods
					}
				}
			}
			// CqDisplayRequest::FormatBucketForDisplay draws one float per pixel per display
			// (ddmanager.cpp:1046-1050), quantised or not.
			for(int d = 0; d < p.n_displays; ++d)
			{
				for(int y = 0; y < ySize; ++y)
					for(int x = 0; x < xSize; ++x)
					{
						float s = rng.nextFloat();
						if(dither)
							dither[(size_t(d)*p.yres + (yPos + y))*p.xres + xPos + x] = s;
					}
			}
		}
	}
}

// ---------------------------------------------------------------------------
void buildFilterTable(const AqhFrameParams& p, std::vector<float>& table)
{
	const int xs = p.xsamples, ys = p.ysamples, n = xs*ys;
	const int xmax = static_cast<int>(lfloorf_(p.filter_xwidth/2.0f));
	const int ymax = static_cast<int>(lfloorf_(p.filter_ywidth/2.0f));
	const float cxw = std::ceil(p.filter_xwidth), cyw = std::ceil(p.filter_ywidth);
	const float xfwo2 = cxw*0.5f, yfwo2 = cyw*0.5f;
	AqhFilterFunc f = p.filter_func ? p.filter_func : aqh_gaussian_filter;
	table.assign(size_t(2*xmax+1)*(2*ymax+1)*n, 0.f);
	for(int py = -ymax; py <= ymax; ++py)
		for(int px = -xmax; px <= xmax; ++px)
		{
			size_t k = size_t((py + ymax)*(2*xmax+1) + px + xmax)*n;
			for(int sy = 0; sy < ys; ++sy)
				for(int sx = 0; sx < xs; ++sx, ++k)
				{
					float fx = (sx + 0.5f)/xs + px - 0.5f;
					float fy = (sy + 0.5f)/ys + py - 0.5f;
					float w = 0;
					if(fx >= -xfwo2 && fy >= -yfwo2 && fx <= xfwo2 && fy <= yfwo2)
						w = f(fx, fy, cxw, cyw);
					table[k] = w;
				}
		}
}

void projectToCircle(float x, float y, float& ox, float& oy)
{
	// CqVector2D::Magnitude2 shortcuts (include/aqsis/math/vector2d.h:132-138)
	float m2;
	if(y == 0.0f) m2 = x*x;
	else if(x == 0.0f) m2 = y*y;
	else m2 = x*x + y*y;
	float r = std::sqrt(m2);
	if(r == 0.0f) { ox = 0; oy = 0; return; }
	float ax = std::fabs(x), ay = std::fabs(y);
	float adj = ((ax < ay) ? ay : ax) / r;
	ox = adj*x; oy = adj*y;
}

void buildDofBounds(int xs, int ys, std::vector<float>& bounds)
{
	bounds.assign(size_t(xs)*ys*4, 0.f);
	const float dx = static_cast<float>(2.0/xs);
	const float dy = static_cast<float>(2.0/ys);
	float minX = -1.0f, minY = -1.0f;
	int which = 0;
	for(int j = 0; j < ys; ++j)
	{
		for(int i = 0; i < xs; ++i)
		{
			float tlx, tly, trx, try_, blx, bly, brx, bry;
			projectToCircle(minX, minY, tlx, tly);
			projectToCircle(minX + dx, minY, trx, try_);
			projectToCircle(minX, minY + dy, blx, bly);
			projectToCircle(minX + dx, minY + dy, brx, bry);
			if((tly > 0.0f && bly < 0.0f) || (tly < 0.0f && bly > 0.0f))
			{
				tlx = minX; blx = minX; trx = minX + dx; brx = minX + dx;
			}
			if((tlx > 0.0f && trx < 0.0f) || (tlx < 0.0f && trx > 0.0f))
			{
				tly = minY; bly = minY + dy; try_ = minY; bry = minY + dy;
			}
			float mnx = tlx, mny = tly, mxx = tlx, mxy = tly;
			const float px[3] = {trx, blx, brx}, py[3] = {try_, bly, bry};
			for(int k = 0; k < 3; ++k)
			{
				mxx = (mxx < px[k]) ? px[k] : mxx;  mxy = (mxy < py[k]) ? py[k] : mxy;
				mnx = (mnx < px[k]) ? mnx : px[k];  mny = (mny < py[k]) ? mny : py[k];
			}
			bounds[4*which] = mnx; bounds[4*which+1] = mny; bounds[4*which+2] = mxx; bounds[4*which+3] = mxy;
			++which;
			minX += dx;
		}
		minX = -1.0f;
		minY += dy;
	}
}

} // namespace aqh
