// matcher.cpp — new bodies for the reference's Matcher (reference
// src/matcher.cpp), interface unchanged (include/matcher.h).  Each function
// flattens the Frame / MapPoint objects it is handed into the SoA arrays of
// include/lorb_cuda.h through their PUBLIC members, calls one C-ABI entry
// point, and writes the resulting indices back into Frame::mvpMapPoints.
// All matching arithmetic runs in the sm_100a kernels.
#include "../include/matcher.h"

#include <cstring>

#include "lorb_host.h"

namespace Simple_ORB_SLAM
{

const int Matcher::TH_HIGH = LORB_TH_HIGH;
const int Matcher::TH_LOW = LORB_TH_LOW;
const int Matcher::HISTO_LENGTH = LORB_HISTO_LENGTH;

namespace
{

// Flat view of a Frame for the projection searches (lorb_frame_view).
struct FlatFrame
{
	std::vector<float> x, y, angle, uright;
	std::vector<int> octave, claim;
	cv::Mat desc;
	lorb_frame_view view;

	explicit FlatFrame(Frame* F)
	{
		const size_t n = F->mnMapPoints;
		x.resize(n); y.resize(n); angle.resize(n); uright.resize(n); octave.resize(n); claim.resize(n);
		for(size_t i=0; i<n; i++)
		{
			const cv::KeyPoint &kp = F->mvKeysUn[i];
			x[i] = kp.pt.x;
			y[i] = kp.pt.y;
			octave[i] = kp.octave;
			angle[i] = kp.angle;
			uright[i] = F->mvuRight[i];
			MapPoint* held = F->mvpMapPoints[i];
			claim[i] = held ? (int)held->mnObs : -1;
		}
		desc = F->GetDescriptors();
		view.n_kp = (int)n;
		view.kp_x = x.data();
		view.kp_y = y.data();
		view.kp_octave = octave.data();
		view.kp_angle = angle.data();
		view.kp_uright = uright.data();
		view.desc = desc.empty() ? nullptr : desc.ptr<uint8_t>();
		view.kp_claim_obs = claim.data();
		view.min_x = F->mnMinX;
		view.max_x = F->mnMaxX;
		view.min_y = F->mnMinY;
		view.max_y = F->mnMaxY;
		view.n_levels = (int)F->mvScaleFactors.size();
		view.scale_factors = F->mvScaleFactors.data();
	}
};

// Brute-force cross-check of a frame's descriptors against a list of map
// points, then max(2*minDist, 30) rejection; shared by the two BF entry points.
size_t MatchAgainstPoints(Frame* frame, const std::vector<MapPoint*>& candidates)
{
	cv::Mat query = frame->GetDescriptors();
	const int nq = query.rows, nt = (int)candidates.size();
	std::vector<uint8_t> train((size_t)nt * LORB_DESC_BYTES);
	for(int j=0; j<nt; j++)
	{
		const cv::Mat d = candidates[j]->GetDescriptor();
		std::memcpy(&train[(size_t)j * LORB_DESC_BYTES], d.ptr<uint8_t>(), LORB_DESC_BYTES);
	}
	const int cap = std::max(1, std::min(nq, nt));
	std::vector<int> mq(cap), mt(cap), md(cap);
	std::vector<uint8_t> keep(cap);
	int n = 0, kept = 0, minDist = -1;
	LORB_HOST_CALL(lorb_match_bf_crosscheck(lorb_host::ctx(), nq ? query.ptr<uint8_t>() : nullptr, nq,
	                                        train.data(), nt, LORB_CROSSCHECK_MUTUAL, mq.data(),
	                                        mt.data(), md.data(), keep.data(), &n, &kept, &minDist));
	for(int i=0; i<n; i++)
		if(keep[i])
			frame->mvpMapPoints[mq[i]] = candidates[mt[i]];
	return (size_t)kept;
}

}

size_t Matcher::SearchByProjection(Frame* currFrame, Frame* prevFrame)
{
	std::vector<MapPoint*> prevMPs;
	for(size_t i=0; i<prevFrame->mnMapPoints; i++)
		if(prevFrame->mvpMapPoints[i] != NULL)
			prevMPs.push_back(prevFrame->mvpMapPoints[i]);
	return MatchAgainstPoints(currFrame, prevMPs);
}

size_t Matcher::SearchLocalPoints(Frame* currFrame, std::set<MapPoint*> vpMPs)
{
	std::vector<MapPoint*> pts;
	for(std::set<MapPoint*>::iterator it = vpMPs.begin(); it != vpMPs.end(); it++)
		if(*it != NULL)
			pts.push_back(*it);
	return MatchAgainstPoints(currFrame, pts);
}

size_t Matcher::SearchByProjection(Frame* F, const std::set<MapPoint*> &vpMapPoints, const float th)
{
	FlatFrame flat(F);
	// the set's iteration order is the claim order of the reference loop
	std::vector<MapPoint*> order(vpMapPoints.begin(), vpMapPoints.end());
	const int n = (int)order.size();
	std::vector<float> px(n), py(n), pxr(n), vcos(n);
	std::vector<int> level(n), nobs(n);
	std::vector<uint8_t> active(n), desc((size_t)n * LORB_DESC_BYTES);
	for(int k=0; k<n; k++)
	{
		MapPoint* pMP = order[k];
		active[k] = (pMP->mbTrackInView && !pMP->IsBad()) ? 1 : 0;
		px[k] = pMP->mTrackProjX;
		py[k] = pMP->mTrackProjY;
		pxr[k] = pMP->mTrackProjXR;
		level[k] = pMP->mnTrackScaleLevel;
		vcos[k] = pMP->mTrackViewCos;
		nobs[k] = (int)pMP->mnObs;
		if(active[k])
		{
			const cv::Mat d = pMP->GetDescriptor();
			std::memcpy(&desc[(size_t)k * LORB_DESC_BYTES], d.ptr<uint8_t>(), LORB_DESC_BYTES);
		}
	}
	std::vector<int> kpForPoint(std::max(n, 1)), pointForKp(std::max(flat.view.n_kp, 1));
	int nmatches = 0;
	LORB_HOST_CALL(lorb_search_proj_points(lorb_host::ctx(), &flat.view, n, px.data(), py.data(),
	                                       pxr.data(), level.data(), vcos.data(), active.data(),
	                                       desc.data(), nobs.data(), th, kpForPoint.data(),
	                                       pointForKp.data(), &nmatches, NULL));
	for(int i=0; i<flat.view.n_kp; i++)
		if(pointForKp[i] >= 0)
			F->mvpMapPoints[i] = order[pointForKp[i]];
	return (size_t)nmatches;
}

size_t Matcher::SearchByProjection(Frame* CurrentFrame, Frame* LastFrame, const float th)
{
	FlatFrame flat(CurrentFrame);
	const int n = (int)LastFrame->mnMapPoints;
	std::vector<uint8_t> valid(n), desc((size_t)n * LORB_DESC_BYTES);
	std::vector<float> xw((size_t)n * 3), angle(n);
	std::vector<int> octave(n), nobs(n);
	for(int i=0; i<n; i++)
	{
		MapPoint* pMP = LastFrame->mvpMapPoints[i];
		valid[i] = (pMP && !LastFrame->mvbOutlier[i]) ? 1 : 0;
		octave[i] = LastFrame->mvKeys[i].octave;
		angle[i] = LastFrame->mvKeysUn[i].angle;
		if(valid[i])
		{
			const cv::Point3f p = pMP->GetPos();
			xw[3*i] = p.x; xw[3*i+1] = p.y; xw[3*i+2] = p.z;
			nobs[i] = (int)pMP->mnObs;
			const cv::Mat d = pMP->GetDescriptor();
			std::memcpy(&desc[(size_t)i * LORB_DESC_BYTES], d.ptr<uint8_t>(), LORB_DESC_BYTES);
		}
	}
	float tcwCur[16], tcwLast[16];
	for(int r=0; r<4; r++)
		for(int c=0; c<4; c++)
		{
			tcwCur[4*r+c] = CurrentFrame->mTcw.at<float>(r,c);
			tcwLast[4*r+c] = LastFrame->mTcw.at<float>(r,c);
		}
	lorb_intrinsics K;
	K.fx = CurrentFrame->fx; K.fy = CurrentFrame->fy; K.cx = CurrentFrame->cx; K.cy = CurrentFrame->cy;
	K.mbf = CurrentFrame->mbf; K.mb = CurrentFrame->mb;
	std::vector<int> kpForItem(std::max(n, 1)), stateForKp(std::max(flat.view.n_kp, 1));
	int nmatches = 0;
	LORB_HOST_CALL(lorb_search_proj_frame(lorb_host::ctx(), &flat.view, tcwCur, tcwLast, &K, n,
	                                      valid.data(), xw.data(), octave.data(), angle.data(),
	                                      desc.data(), nobs.data(), th, kpForItem.data(),
	                                      stateForKp.data(), &nmatches, NULL));
	for(int i=0; i<flat.view.n_kp; i++)
	{
		if(stateForKp[i] >= 0)
			CurrentFrame->mvpMapPoints[i] = LastFrame->mvpMapPoints[stateForKp[i]];
		else if(stateForKp[i] == -2)
			CurrentFrame->mvpMapPoints[i] = static_cast<MapPoint*>(NULL);
	}
	return (size_t)nmatches;
}

// Scalar helpers kept for API completeness: Frame::ComputeStereoMatches calls
// DescriptorDistance per candidate (reference src/frame.cpp:217).  They are not
// part of the data-parallel path (the kernels carry their own copies).
int Matcher::DescriptorDistance(const cv::Mat &a, const cv::Mat &b)
{
	const uint32_t *pa = a.ptr<uint32_t>();
	const uint32_t *pb = b.ptr<uint32_t>();
	int dist = 0;
	for(int i=0; i<LORB_DESC_BYTES/4; i++)
		dist += __builtin_popcount(pa[i] ^ pb[i]);
	return dist;
}

float Matcher::RadiusByViewingCos(const float &viewCos)
{
	return viewCos > 0.998 ? 2.5f : 4.0f;
}

void Matcher::ComputeThreeMaxima(vector<int>* histo, const int L, int &ind1, int &ind2, int &ind3)
{
	int best[3] = {0, 0, 0};
	int idx[3] = {ind1, ind2, ind3};
	for(int i=0; i<L; i++)
	{
		const int s = (int)histo[i].size();
		// insertion into the sorted top three, strict '>' so the first bin wins ties
		int pos = 3;
		if(s > best[0]) pos = 0;
		else if(s > best[1]) pos = 1;
		else if(s > best[2]) pos = 2;
		for(int k=2; k>pos; k--) { best[k] = best[k-1]; idx[k] = idx[k-1]; }
		if(pos < 3) { best[pos] = s; idx[pos] = i; }
	}
	if(best[1] < 0.1f*(float)best[0]) { idx[1] = -1; idx[2] = -1; }
	else if(best[2] < 0.1f*(float)best[0]) { idx[2] = -1; }
	ind1 = idx[0]; ind2 = idx[1]; ind3 = idx[2];
}

}
