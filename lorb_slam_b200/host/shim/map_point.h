// Stand-in for the reference's MapPoint (include/map_point.h): exactly the
// public members and getters Matcher / BA reach (include/map_point.h:54-69 and
// GetPos / GetDescriptor / IsBad / GetObservations / SetWorldPos).
#ifndef MAP_POINT_H
#define MAP_POINT_H
#include "common.h"
namespace Simple_ORB_SLAM
{
class Frame;
class MapPoint
{
public:
	MapPoint() {}
	void SetWorldPos(cv::Point3f p) { mWorldPos = p; }
	cv::Point3f GetPos() { return mWorldPos; }
	cv::Mat GetDescriptor() { return mDescriptor.clone(); }
	std::map<Frame*, size_t> GetObservations() { return mObservations; }
	void AddObservation(Frame* pF, size_t id) { mObservations[pF] = id; mnObs++; }
	bool IsBad() { return mbBadFlag; }

public:
	size_t mnObs = 0;
	float mTrackProjX = 0, mTrackProjY = 0, mTrackProjXR = 0;
	int mnTrackScaleLevel = 0;
	float mTrackViewCos = 0;
	bool mbTrackInView = false;

	// test-harness access (private in the reference)
	cv::Point3f mWorldPos;
	cv::Mat mDescriptor;
	std::map<Frame*, size_t> mObservations;
	bool mbBadFlag = false;
};
}
#endif
