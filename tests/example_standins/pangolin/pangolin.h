// TEST STAND-IN (tests/test_example_links.py only; not part of the product): the slice of the Pangolin /
// GLUT API that examples/MonoSlamSceneLib1.cpp and scenelib2/support/pangolin_util.h touch, as headless no-ops,
// so that the reference's UNCHANGED example can be compiled and linked against the host shim of libsl2b200.
#pragma once
#include <string>

namespace pangolin {
enum AxisDirection { AxisNone, AxisNegX, AxisX, AxisNegY, AxisY, AxisNegZ, AxisZ };
enum MouseButton { MouseButtonLeft = 1, MouseButtonMiddle = 2, MouseButtonRight = 4, MouseWheelUp = 8, MouseWheelDown = 16 };
struct OpenGlMatrix {};
inline OpenGlMatrix ProjectionMatrix(int, int, double, double, double, double, double, double) { return OpenGlMatrix(); }
inline OpenGlMatrix ModelViewLookAt(double, double, double, double, double, double, AxisDirection) { return OpenGlMatrix(); }
struct OpenGlRenderState {
  OpenGlRenderState(const OpenGlMatrix &, const OpenGlMatrix &) {}
};
struct View;
struct Handler {
  virtual ~Handler() {}
  virtual void Mouse(View &, MouseButton, int, int, bool, int) {}
};
struct Handler3D : Handler {
  Handler3D(OpenGlRenderState &, AxisDirection = AxisNone, float = 0.01f) {}
};
struct View {
  View &SetBounds(double, double, double, double) { return *this; }
  View &SetBounds(double, double, double, double, double) { return *this; }
  View &SetHandler(Handler *h) { handler = h; return *this; }
  void ActivateScissorAndClear() {}
  void ActivateScissorAndClear(const OpenGlRenderState &) {}
  void Render() {}
  void SaveOnRender(const std::string &) {}
  Handler *handler = nullptr;
};
inline View &DisplayBase() { static View v; return v; }
inline View &CreatePanel(const std::string &) { static View v[8]; static int i = 0; return v[i++ % 8]; }
inline View &CreateDisplay() { static View v; return v; }
inline View &Display(const std::string &) { static View v; return v; }
inline void CreateWindowAndBind(const std::string &, int, int) {}
// the headless stand-in runs the main loop body once and then asks to quit
inline bool ShouldQuit() { static int calls = 0; return calls++ > 0; }
inline bool HasResized() { return false; }
inline bool HadInput() { return false; }
inline void FinishFrame() {}
template <typename T>
struct Var {
  Var(const std::string &, const T &v) : value(v) {}
  Var(const std::string &, const T &v, bool) : value(v) {}
  operator const T &() const { return value; }
  T value;
};
inline bool Pushed(Var<bool> &b) { const bool v = b.value; b.value = false; return v; }
}  // namespace pangolin

inline void glutInit(int *, char **) {}
