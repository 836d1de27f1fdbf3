// Stand-in for the reference's Frame (include/frame.h): the public members the
// hot-path operators read or write (include/frame.h:76-110) and the getters /
// setters they call (GetDescriptors, GetKp2d, SetPose, GetCovisibleFrames,
// GetMapPoints, IsBad).  SetPose mirrors reference src/frame.cpp:578-594
// (Rodrigues into mTcw) because the frame-to-frame search reads mTcw.
#ifndef FRAME_H
#define FRAME_H
#include "common.h"
#include "camera.h"
#include "map_point.h"
namespace Simple_ORB_SLAM
{
const int FRAME_GRID_ROWS = 48;
const int FRAME_GRID_COLS = 64;

class Frame
{
public:
	Frame() : mTcw(4, 4, CV_32F), mTvec(3, 1, CV_32F), mRvec(3, 1, CV_32F)
	{
		for(int i=0;i<4;i++) mTcw.at<float>(i,i) = 1.0f;
	}
	cv::Point2f GetKp2d(size_t i) { return cv::Point2f(mvKeysUn[i].pt.x, mvKeysUn[i].pt.y); }
	cv::Mat GetDescriptors() { return mDescriptors.clone(); }
	cv::Mat GetDescriptor(size_t idx) { return mDescriptors.row((int)idx).clone(); }
	std::vector<MapPoint*> GetMapPoints() { return mvpMapPoints; }
	std::vector<Frame*> GetCovisibleFrames() { return mvpOrderedKeyFrames; }
	bool IsBad() { return mbBadFlag; }
	void SetPose(const cv::Mat& Tvec, const cv::Mat& Rvec)
	{
		mTvec = Tvec.clone();
		mRvec = Rvec.clone();
		const float rx = mRvec.at<float>(0), ry = mRvec.at<float>(1), rz = mRvec.at<float>(2);
		const float th = std::sqrt(rx*rx + ry*ry + rz*rz);
		float R[9] = {1,0,0, 0,1,0, 0,0,1};
		if(th > 1e-12f)
		{
			const float c = std::cos(th), s = std::sin(th), kx = rx/th, ky = ry/th, kz = rz/th, o = 1-c;
			const float M[9] = {c+o*kx*kx, o*kx*ky-s*kz, o*kx*kz+s*ky,
			                    o*ky*kx+s*kz, c+o*ky*ky, o*ky*kz-s*kx,
			                    o*kz*kx-s*ky, o*kz*ky+s*kx, c+o*kz*kz};
			for(int i=0;i<9;i++) R[i] = M[i];
		}
		for(int r=0;r<3;r++)
		{
			for(int c=0;c<3;c++) mTcw.at<float>(r,c) = R[3*r+c];
			mTcw.at<float>(r,3) = mTvec.at<float>(r);
		}
	}

public:
	size_t mnMapPoints = 0;
	std::vector<MapPoint*> mvpMapPoints;
	Camera* mpCamera = nullptr;
	cv::Mat mTcw;
	cv::Mat mTvec, mRvec;
	std::vector<float> mvuRight;
	std::vector<bool> mvbOutlier;
	std::vector<cv::KeyPoint> mvKeysUn;
	std::vector<cv::KeyPoint> mvKeys;
	float mbf = 0, mb = 0, cx = 0, cy = 0, fx = 0, fy = 0;
	float mnMinX = 0, mnMaxX = 0, mnMinY = 0, mnMaxY = 0;
	int mnScaleLevels = 0;
	vector<float> mvScaleFactors;

	// test-harness access (private in the reference)
	cv::Mat mDescriptors;
	std::vector<Frame*> mvpOrderedKeyFrames;
	bool mbBadFlag = false;
};
}
#endif
