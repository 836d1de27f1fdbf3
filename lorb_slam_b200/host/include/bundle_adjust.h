// bundle_adjust.h — drop-in replacement header for the reference's
// include/bundle_adjust.h (:12-21): class BA with two static entry points.
// The bodies (src/bundle_adjust.cpp of this directory) assemble the same
// problem the reference hands to Ceres and solve it on the GPU through
// include/lorb_cuda.h.  Call sites: reference src/visual_odometry.cpp:135,207
// and (commented out in the reference) src/local_mapping.cpp:32.
#ifndef BUNDLE_ADJUST_H
#define BUNDLE_ADJUST_H

#include "common.h"
#include "frame.h"
#include "camera.h"

namespace Simple_ORB_SLAM
{

class BA
{
public:
	BA();

	// pose-only optimisation of one frame against its matched map points
	void static ProjectPoseOptimization(Frame* curr);

	// local bundle adjustment over the covisibility window of pCurrFrame
	void static LocalPoseOptimization(Frame* pCurrFrame);
};

}

#endif
