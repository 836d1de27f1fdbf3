// Stand-in for the reference's Camera (include/camera.h:17-52): the hot path
// only consumes the float intrinsics (include/camera.h:26).
#ifndef CAMERA_H
#define CAMERA_H
#include "common.h"
namespace Simple_ORB_SLAM
{
class Camera
{
public:
	float bf = 0, cx = 0, cy = 0, fx = 0, fy = 0;
};
}
#endif
