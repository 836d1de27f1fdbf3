// matcher.h — drop-in replacement header for the reference's include/matcher.h.
//
// The class below keeps the reference's public surface exactly (reference
// include/matcher.h:15-36): same namespace, same static member functions with
// the same parameter types (including the by-value std::set of
// SearchLocalPoints and the `size_t static` spelling's meaning), same three
// public constants.  Only the bodies (src/matcher.cpp of this directory) are
// new: they flatten Frame / MapPoint state into plain arrays and call the
// sm_100a kernels through include/lorb_cuda.h.  Callers — Frame, VisualOdometry,
// LocalMapping (reference src/visual_odometry.cpp:124,129,201,437,
// src/frame.cpp:130,195,217) — compile against this header unchanged.
#ifndef MATCHER_H
#define MATCHER_H

#include "common.h"
#include "frame.h"
#include "map_point.h"

namespace Simple_ORB_SLAM
{

class Frame;
class MapPoint;

class Matcher
{
public:
	Matcher();

	// brute-force cross-checked Hamming match of curr's descriptors against prev's map points
	size_t static SearchByProjection(Frame* curr, Frame* prev);
	// projection of LastFrame's map points into CurrentFrame, window th * scale
	size_t static SearchByProjection(Frame* CurrentFrame, Frame* LastFrame, const float th);

	// brute-force match against a set of map points (set taken by value, as in the reference)
	size_t static SearchLocalPoints(Frame* curr, std::set<MapPoint*> vpMPs);
	// projection-guided search of already projected local map points
	size_t static SearchByProjection(Frame* F, const std::set<MapPoint*> &vpMapPoints, const float th);

	int static DescriptorDistance(const cv::Mat &a, const cv::Mat &b);
	float static RadiusByViewingCos(const float &viewCos);
	void static ComputeThreeMaxima(vector<int>* histo, const int L, int &ind1, int &ind2, int &ind3);

public:
    static const int TH_LOW;
    static const int TH_HIGH;
    static const int HISTO_LENGTH;
};

}

#endif
