// TEST STAND-IN: the mouse handlers of scenelib2/support/pangolin_util.cpp are GUI code (out of scope); the link
// test only needs their symbols.
#include "monoslam.h"
#include "support/pangolin_util.h"
namespace SceneLib2 {
void Handler3D::Mouse(pangolin::View &, pangolin::MouseButton, int, int, bool, int) {}
void Handler2D::Mouse(pangolin::View &, pangolin::MouseButton, int, int, bool, int) {}
}  // namespace SceneLib2
