// Stand-in for the reference's include/common.h (:6-26): umbrella include plus
// the two using-directives the reference headers rely on.
#ifndef COMMON_H
#define COMMON_H
#include "opencv2/shim.hpp"
using namespace cv;

#include <iostream>
#include <list>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <string>
#include <vector>
using namespace std;
#endif
