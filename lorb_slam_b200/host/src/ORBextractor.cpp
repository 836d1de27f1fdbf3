// ORBextractor.cpp — body of the drop-in ORBextractor (include/ORBextractor.h of this directory)
// over the C ABI.  Replaces the reference's src/ORBextractor.cpp (1.2 kLoC of CPU image code):
// constructor = the scale tables callers read (:428-445), operator() = one lorb_orb_extract call.
#include "../include/ORBextractor.h"

#include <cstdint>

#include "lorb_host.h"

namespace Simple_ORB_SLAM
{

namespace
{
// the sampling pattern of the reference's constructor (:464-467)
const int kOrbPattern31[1024] = {
#include "orb_pattern_31.inc"
};
}

ORBextractor::ORBextractor(int _nfeatures, float _scaleFactor, int _nlevels, int _iniThFAST, int _minThFAST)
	: nfeatures(_nfeatures), scaleFactor(_scaleFactor), nlevels(_nlevels), iniThFAST(_iniThFAST),
	  minThFAST(_minThFAST)
{
	// float products, level by level, like the reference: callers compare against these values
	mvScaleFactor.assign(nlevels, 1.0f);
	mvLevelSigma2.assign(nlevels, 1.0f);
	for(int l = 1; l < nlevels; l++)
	{
		mvScaleFactor[l] = mvScaleFactor[l-1] * _scaleFactor;
		mvLevelSigma2[l] = mvScaleFactor[l] * mvScaleFactor[l];
	}
	mvInvScaleFactor.resize(nlevels);
	mvInvLevelSigma2.resize(nlevels);
	for(int l = 0; l < nlevels; l++)
	{
		mvInvScaleFactor[l] = 1.0f / mvScaleFactor[l];
		mvInvLevelSigma2[l] = 1.0f / mvLevelSigma2[l];
	}
	mvImagePyramid.resize(nlevels);
}

void ORBextractor::operator()(cv::InputArray _image, cv::InputArray /*mask*/, std::vector<cv::KeyPoint>& keypoints,
                              cv::OutputArray _descriptors)
{
	if(_image.empty())
		return;
	cv::Mat image = _image.getMat();
	if(image.type() != CV_8U)
		lorb_host::die("ORBextractor: image must be CV_8UC1", LORB_ERR_ARG);

	lorb_orb_params prm;
	prm.nfeatures = nfeatures;
	prm.scale_factor = (float)scaleFactor;
	prm.nlevels = nlevels;
	prm.ini_th_fast = iniThFAST;
	prm.min_th_fast = minThFAST;

	std::vector<int> lw(nlevels), lh(nlevels);
	LORB_HOST_CALL(lorb_orb_level_sizes(&prm, image.cols, image.rows, lw.data(), lh.data(), NULL, NULL));
	std::vector<uint8_t*> levels(nlevels);
	for(int l = 0; l < nlevels; l++)
	{
		mvImagePyramid[l].create(lh[l], lw[l], CV_8U);
		levels[l] = mvImagePyramid[l].ptr<uint8_t>();
	}

	int cap = 0;  // usually nfeatures + a few; wide levels with small shares can return more
	LORB_HOST_CALL(lorb_orb_max_keypoints(&prm, image.cols, image.rows, &cap));
	if(cap < 1)
		cap = 1;  // every level below one 30 px cell: no keypoints, but the output arrays must exist
	std::vector<float> x(cap), y(cap), angle(cap), response(cap), size(cap);
	std::vector<int> octave(cap);
	cv::Mat desc(cap, 32, CV_8U);
	int n = 0;
	LORB_HOST_CALL(lorb_orb_extract(lorb_host::ctx(), image.ptr<uint8_t>(), image.cols, image.rows, (int)image.step,
	                                &prm, kOrbPattern31, cap, x.data(), y.data(), octave.data(), angle.data(),
	                                response.data(), size.data(), desc.ptr<uint8_t>(), &n, levels.data()));

	keypoints.clear();
	keypoints.resize(n);
	for(int i = 0; i < n; i++)
	{
		cv::KeyPoint& kp = keypoints[i];
		kp.pt = cv::Point2f(x[i], y[i]);
		kp.size = size[i];
		kp.angle = angle[i];
		kp.response = response[i];
		kp.octave = octave[i];
		kp.class_id = -1;
	}
	if(n == 0)
	{
		_descriptors.release();
		return;
	}
	_descriptors.create(n, 32, CV_8U);
	cv::Mat out = _descriptors.getMat();
	for(int i = 0; i < n; i++)
		std::memcpy(out.ptr<uint8_t>(i), desc.ptr<uint8_t>(i), 32);
}

}  // namespace Simple_ORB_SLAM
