// host_filters.cpp -- the standard RenderMan pixel filters (RtFilterFunc values).
//
// Behavioural twins of RiGaussianFilter & co (libs/core/filters.cpp:71-348).  A pixel
// filter is only ever *tabulated* on the host (bucketprocessor.cpp:811-856); the table
// then drives the device filter kernel.  Precision matters for bit parity of the table,
// so each function states where the reference evaluates in double: its C math calls
// resolve to the ::f(double) overloads (probed against the in-place compiled reference,
// tests/test_oracle_leaves.py (filter values against tests/golden/ref_leaves.npz, generated from libs/core/filters.cpp)).
#include "../../include/aqsis_b200_hider.h"

#include <cmath>
#include <cstring>

namespace {
const float kPi = 3.14159265359f;   // RI_PI, include/aqsis/ri/ri.h:40
inline double dmin(double a, double b) { return (a < b) ? a : b; }

// CqMitchellFilter::Evaluate(x), filters.cpp:49-58, all in float.
inline float mitchell1d(float x, float B, float C)
{
	x = std::fabs(2.f * x);
	if(x > 1.f)
		return ((-B - 6*C) * x*x*x + (6*B + 30*C) * x*x + (-12*B - 48*C) * x + (8*B + 24*C)) * (1.f/6.f);
	return ((12 - 9*B - 6*C) * x*x*x + (-18 + 12*B + 6*C) * x*x + (6 - 2*B)) * (1.f/6.f);
}
// One axis of the cosine-windowed sinc (filters.cpp:268-287).
inline float sinc1d(float t, float width)
{
	if(t == 0.0f)
		return 1.0f;
	t *= kPi;
	return static_cast<float>(std::cos(0.5 * t / width) * std::sin(static_cast<double>(t)) / t);
}
} // namespace

extern "C" {

float aqh_box_filter(float x, float y, float xw, float yw)
{
	double fx = (std::fabs(static_cast<double>(x)) <= xw / 2.0) ? 1.0 : 0.0;
	double fy = (std::fabs(static_cast<double>(y)) <= yw / 2.0) ? 1.0 : 0.0;
	return static_cast<float>(dmin(fx, fy));
}

float aqh_triangle_filter(float x, float y, float xw, float yw)
{
	float hxw = static_cast<float>(xw / 2.0);
	float hyw = static_cast<float>(yw / 2.0);
	float absx = std::fabs(x), absy = std::fabs(y);
	double fx = (absx <= hxw) ? static_cast<double>((hxw - absx) / hxw) : 0.0;
	double fy = (absy <= hyw) ? static_cast<double>((hyw - absy) / hyw) : 0.0;
	return static_cast<float>(dmin(fx, fy));
}

float aqh_gaussian_filter(float x, float y, float xw, float yw)
{
	x /= xw;
	y /= yw;
	return static_cast<float>(std::exp(-8.0 * static_cast<double>(x*x + y*y)));
}

float aqh_catmullrom_filter(float x, float y, float /*xw*/, float /*yw*/)
{
	// Radial RI-spec-3.2 form; the widths are ignored (filters.cpp:239-242).
	float r2 = x*x + y*y;
	float r = static_cast<float>(std::sqrt(static_cast<double>(r2)));
	if(r >= 2.0)
		return 0.0f;
	if(r < 1.0)
		return static_cast<float>(3.0*r*r2 - 5.0*r2 + 2.0);
	float mr3 = -r*r2;                       // "-r*r2" is a float product in the reference
	return static_cast<float>(mr3 + 5.0*r2 - 8.0*r + 4.0);
}

float aqh_sinc_filter(float x, float y, float xw, float yw)
{
	return sinc1d(x, xw) * sinc1d(y, yw);
}

float aqh_mitchell_filter(float x, float y, float xw, float yw)
{
	const float B = 1/3.0f, C = 1/3.0f;
	float invX = 1.0f/xw, invY = 1.0f/yw;
	return mitchell1d(x*invX, B, C) * mitchell1d(y*invY, B, C);
}

float aqh_disk_filter(float x, float y, float xw, float yw)
{
	double xx = x*x, yy = y*y;
	xw *= 0.5f; yw *= 0.5f;
	double d = xx / (xw*xw) + yy / (yw*yw);
	return (d < 1.0) ? 1.0f : 0.0f;
}

float aqh_bessel_filter(float x, float y, float xw, float yw)
{
	double xx = x*x, yy = y*y;
	xw *= 0.5f; yw *= 0.5f;
	double w = xx / (xw*xw) + yy / (yw*yw);
	if(w >= 1.0)
		return 0.0f;
	double d = std::sqrt(xx + yy);
	if(d == 0.0)
		return kPi;
	w = std::cos(0.5 * kPi * std::sqrt(w));
	return static_cast<float>(w * 2*j1(kPi * d) / d);
}

AqhFilterFunc aqh_filter_by_name(const char* name)
{
	if(!name) return 0;
	if(!std::strcmp(name, "box")) return aqh_box_filter;
	if(!std::strcmp(name, "triangle")) return aqh_triangle_filter;
	if(!std::strcmp(name, "gaussian")) return aqh_gaussian_filter;
	if(!std::strcmp(name, "catmull-rom") || !std::strcmp(name, "catmullrom")) return aqh_catmullrom_filter;
	if(!std::strcmp(name, "sinc")) return aqh_sinc_filter;
	if(!std::strcmp(name, "mitchell")) return aqh_mitchell_filter;
	if(!std::strcmp(name, "disk")) return aqh_disk_filter;
	if(!std::strcmp(name, "bessel")) return aqh_bessel_filter;
	return 0;
}

} // extern "C"
