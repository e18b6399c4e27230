// ref_pdiff.cpp -- TEST INFRASTRUCTURE.  C entry point over the reference's own perceptual image diff
// (thirdparty/pdiff: Metric.cpp, LPyramid.cpp, CompareArgs.cpp compiled in place from /root/reference -- Yee's
// method, the tool aqsis' regression suite uses and BASELINE.json's "zero pdiff-detected differences" refers to).
// Only the TIFF reader of the reference tool is left out (libtiff is absent here): images come in as RGBA8 arrays.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>

#include "CompareArgs.h"
#include "Metric.h"
#include "RGBAImage.h"

// CompareArgs::Parse_Args (unused here) references the TIFF reader of RGBAImage.cpp
RGBAImage* RGBAImage::ReadTiff(char*) { return 0; }

extern "C" {

// Returns 1 when pdiff passes (binary identical or perceptually indistinguishable), 0 when it reports visible
// differences, -1 on bad arguments.  *pixels_failed receives the count pdiff prints (0 on a pass below threshold).
// Defaults of the reference tool (CompareArgs.cpp): fov 45, gamma 2.2, luminance 100, threshold 100 pixels;
// pass threshold_pixels = 1 to demand ZERO failing pixels.
int ref_pdiff(const uint8_t* rgba_a, const uint8_t* rgba_b, int width, int height,
              float fov, float gamma, float luminance, unsigned threshold_pixels, int* pixels_failed, int* identical)
{
	if(!rgba_a || !rgba_b || width < 1 || height < 1) return -1;
	CompareArgs args;
	args.ImgA = new RGBAImage(width, height);
	args.ImgB = new RGBAImage(width, height);
	for(int y = 0; y < height; ++y)
		for(int x = 0; x < width; ++x)
		{
			const uint8_t* a = rgba_a + 4*(size_t(y)*width + x);
			const uint8_t* b = rgba_b + 4*(size_t(y)*width + x);
			args.ImgA->Set(x, y, a[0] | (a[1] << 8) | (a[2] << 16) | ((unsigned)a[3] << 24));
			args.ImgB->Set(x, y, b[0] | (b[1] << 8) | (b[2] << 16) | ((unsigned)b[3] << 24));
		}
	args.Verbose = false;
	args.FieldOfView = fov; args.Gamma = gamma; args.Luminance = luminance; args.ThresholdPixels = threshold_pixels;
	const bool pass = Yee_Compare(args);
	int failed = 0;
	const std::string::size_type at = args.ErrorStr.find("pixels are different");
	if(at != std::string::npos)
	{
		std::string::size_type b = args.ErrorStr.rfind('\n', at);
		failed = std::atoi(args.ErrorStr.c_str() + (b == std::string::npos ? 0 : b + 1));
	}
	if(pixels_failed) *pixels_failed = failed;
	if(identical) *identical = args.ErrorStr.find("binary identical") != std::string::npos ? 1 : 0;
	return pass ? 1 : 0;            // ~CompareArgs deletes the images
}

} // extern "C"
