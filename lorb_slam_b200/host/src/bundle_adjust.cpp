// bundle_adjust.cpp — new bodies for the reference's BA (reference
// src/bundle_adjust.cpp), interface unchanged (include/bundle_adjust.h).
// The problem assembly follows the reference line by line (which residuals,
// which parameter blocks, float storage in and out); the solve itself — what
// the reference hands to Ceres — runs in the sm_100a kernels behind
// lorb_ba_pose_only / lorb_ba_local with Ceres' default options + DENSE_SCHUR
// semantics (lorb_ba_default_options).
#include "../include/bundle_adjust.h"

#include <unordered_map>

#include "lorb_host.h"

namespace Simple_ORB_SLAM
{

BA::BA()
{
}

namespace
{
cv::Mat Vec3(double a, double b, double c)
{
	cv::Mat m(3, 1, CV_32F);
	m.at<float>(0) = (float)a;
	m.at<float>(1) = (float)b;
	m.at<float>(2) = (float)c;
	return m;
}
}

// reference src/bundle_adjust.cpp:158-202
void BA::ProjectPoseOptimization(Frame* pCurrFrame)
{
	double rt[6];
	for(int i=0; i<3; i++)
	{
		rt[i] = pCurrFrame->mRvec.at<float>(i);
		rt[3+i] = pCurrFrame->mTvec.at<float>(i);
	}
	std::vector<float> xw, uv;
	for(size_t i=0; i<pCurrFrame->mnMapPoints; i++)
	{
		MapPoint* pMP = pCurrFrame->mvpMapPoints[i];
		if(pMP == NULL)
			continue;
		const cv::Point2f kp2d = pCurrFrame->GetKp2d(i);
		const cv::Point3f kp3d = pMP->GetPos();
		xw.push_back(kp3d.x); xw.push_back(kp3d.y); xw.push_back(kp3d.z);
		uv.push_back(kp2d.x); uv.push_back(kp2d.y);
	}
	const Camera* cam = pCurrFrame->mpCamera;
	const float K[4] = {cam->fx, cam->fy, cam->cx, cam->cy};
	lorb_ba_options opt;
	lorb_ba_default_options(&opt);
	lorb_ba_summary summary;
	LORB_HOST_CALL(lorb_ba_pose_only(lorb_host::ctx(), (int)(uv.size()/2), xw.data(), uv.data(), K, rt,
	                                 &opt, &summary));
	pCurrFrame->SetPose(Vec3(rt[3], rt[4], rt[5]), Vec3(rt[0], rt[1], rt[2]));
}

// reference src/bundle_adjust.cpp:207-330
void BA::LocalPoseOptimization(Frame* pCurrFrame)
{
	// window: the current frame plus its non-bad covisible frames (:210-220)
	std::vector<Frame*> vpLocalFrames;
	vpLocalFrames.push_back(pCurrFrame);
	std::vector<Frame*> vpCovisibleFrames = pCurrFrame->GetCovisibleFrames();
	for(size_t i=0; i<vpCovisibleFrames.size(); i++)
		if(!vpCovisibleFrames[i]->IsBad())
			vpLocalFrames.push_back(vpCovisibleFrames[i]);
	std::unordered_map<Frame*, int> frameIndex;
	for(size_t i=0; i<vpLocalFrames.size(); i++)
		if(!frameIndex.count(vpLocalFrames[i]))  // std::find returns the first occurrence
			frameIndex[vpLocalFrames[i]] = (int)i;

	// points: first-occurrence order of the window frames' map points (:224-241);
	// a hash set replaces the reference's O(P^2) std::find without changing the order
	std::vector<MapPoint*> vpLocalMapPoints;
	std::unordered_map<MapPoint*, int> pointIndex;
	for(size_t i=0; i<vpLocalFrames.size(); i++)
	{
		std::vector<MapPoint*> vpMPs = vpLocalFrames[i]->GetMapPoints();
		for(size_t j=0; j<vpMPs.size(); j++)
		{
			MapPoint* pMP = vpMPs[j];
			if(pMP == NULL || pMP->IsBad() || pointIndex.count(pMP))
				continue;
			pointIndex[pMP] = (int)vpLocalMapPoints.size();
			vpLocalMapPoints.push_back(pMP);
		}
	}

	// float state widened to double (:244-265)
	std::vector<double> cams(vpLocalFrames.size()*6), pts(vpLocalMapPoints.size()*3);
	for(size_t i=0; i<vpLocalFrames.size(); i++)
		for(int k=0; k<3; k++)
		{
			cams[6*i+k] = vpLocalFrames[i]->mRvec.at<float>(k);
			cams[6*i+3+k] = vpLocalFrames[i]->mTvec.at<float>(k);
		}
	for(size_t i=0; i<vpLocalMapPoints.size(); i++)
	{
		const cv::Point3f p = vpLocalMapPoints[i]->GetPos();
		pts[3*i] = p.x; pts[3*i+1] = p.y; pts[3*i+2] = p.z;
	}

	// residual blocks (:270-303): in-window observer -> (point, pose) block,
	// out-of-window observer -> point block with that frame's fixed float pose.
	// The pixel comes from GetKp2d (mvKeysUn): the reference's GetKps2d() vector
	// is never filled in the live pipeline.
	std::vector<int> obsCam, obsPt, fixPt;
	std::vector<float> obsUv, fixUv, fixRt;
	for(size_t i=0; i<vpLocalMapPoints.size(); i++)
	{
		std::map<Frame*, size_t> observations = vpLocalMapPoints[i]->GetObservations();
		for(std::map<Frame*, size_t>::const_iterator it = observations.begin(); it != observations.end(); it++)
		{
			Frame* pF = it->first;
			if(pF->IsBad())
				continue;
			const cv::Point2f kp2d = pF->GetKp2d(it->second);
			std::unordered_map<Frame*, int>::const_iterator w = frameIndex.find(pF);
			if(w == frameIndex.end())
			{
				fixPt.push_back((int)i);
				fixUv.push_back(kp2d.x); fixUv.push_back(kp2d.y);
				for(int k=0; k<3; k++) fixRt.push_back(pF->mRvec.at<float>(k));
				for(int k=0; k<3; k++) fixRt.push_back(pF->mTvec.at<float>(k));
			}
			else
			{
				obsCam.push_back(w->second);
				obsPt.push_back((int)i);
				obsUv.push_back(kp2d.x); obsUv.push_back(kp2d.y);
			}
		}
	}

	const Camera* cam = pCurrFrame->mpCamera;
	const float K[4] = {cam->fx, cam->fy, cam->cx, cam->cy};
	lorb_ba_options opt;
	lorb_ba_default_options(&opt);
	lorb_ba_summary summary;
	LORB_HOST_CALL(lorb_ba_local(lorb_host::ctx(), (int)vpLocalFrames.size(), cams.data(),
	                             (int)vpLocalMapPoints.size(), pts.data(), (int)obsCam.size(),
	                             obsCam.data(), obsPt.data(), obsUv.data(), (int)fixPt.size(),
	                             fixPt.data(), fixUv.data(), fixRt.data(), K, &opt, &summary));

	// write back as float (:317-329)
	for(size_t i=0; i<vpLocalFrames.size(); i++)
		vpLocalFrames[i]->SetPose(Vec3(cams[6*i+3], cams[6*i+4], cams[6*i+5]),
		                          Vec3(cams[6*i], cams[6*i+1], cams[6*i+2]));
	for(size_t i=0; i<vpLocalMapPoints.size(); i++)
		vpLocalMapPoints[i]->SetWorldPos(cv::Point3f((float)pts[3*i], (float)pts[3*i+1], (float)pts[3*i+2]));
}

}
