// ORBextractor.h — drop-in replacement header for the reference's include/ORBextractor.h.
//
// Same namespace, class name, constructor, call operator, accessors and the public mvImagePyramid
// member (reference include/ORBextractor.h:47-96), so Frame (src/frame.cpp:34-56, 132-133) and
// example/test.cpp keep compiling unchanged.  What is gone is the CPU machinery behind it
// (ExtractorNode, ComputePyramid, ComputeKeyPointsOctTree, DistributeOctTree): the body
// (src/ORBextractor.cpp of this directory) hands the image to lorb_orb_extract
// (include/lorb_cuda.h), which runs the pyramid, FAST, the quadtree selection, blur, orientation and
// descriptors as sm_100a kernels: one CUDA graph per frame.
#ifndef ORBEXTRACTOR_H
#define ORBEXTRACTOR_H

#include <vector>

#include "common.h"

namespace Simple_ORB_SLAM
{

class ORBextractor
{
public:
	enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };

	ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST);
	~ORBextractor() {}

	// keypoints + descriptors of an 8-bit single-channel image; the mask is ignored, as in the reference
	void operator()(cv::InputArray image, cv::InputArray mask, std::vector<cv::KeyPoint>& keypoints,
	                cv::OutputArray descriptors);

	int inline GetLevels() { return nlevels; }
	float inline GetScaleFactor() { return scaleFactor; }
	std::vector<float> inline GetScaleFactors() { return mvScaleFactor; }
	std::vector<float> inline GetInverseScaleFactors() { return mvInvScaleFactor; }
	std::vector<float> inline GetScaleSigmaSquares() { return mvLevelSigma2; }
	std::vector<float> inline GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }

	// the scale pyramid of the last image (Frame::ComputeStereoMatches reads it)
	std::vector<cv::Mat> mvImagePyramid;

protected:
	int nfeatures;
	double scaleFactor;
	int nlevels;
	int iniThFAST;
	int minThFAST;

	std::vector<float> mvScaleFactor;
	std::vector<float> mvInvScaleFactor;
	std::vector<float> mvLevelSigma2;
	std::vector<float> mvInvLevelSigma2;
};

}  // namespace Simple_ORB_SLAM

#endif
