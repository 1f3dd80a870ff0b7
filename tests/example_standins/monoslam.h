// TEST STAND-IN: the include the reference's example resolves as "monoslam.h" -> the host shim of libsl2b200.
#pragma once
#include <pangolin/pangolin.h>

#include "scenelib2_b200.h"
#if !defined(SL2_USE_REAL_EIGEN_OPENCV)
namespace cv {
inline bool imwrite(const std::string &, const Mat &) { return false; }  // used by the example's "Save Raw Images"
}
#endif
