// sl2_headless.cpp — headless analogue of the reference's only executable
// (examples/MonoSlamSceneLib1.cpp:132-142: GetFrame -> GoOneStep) without Pangolin/GLUT:
//   sl2_headless <config.cfg> <frames.raw> <width> <height> <nframes> [out_state.txt]
// frames.raw = nframes * height * width bytes (8-bit gray).  Prints the camera state per frame
// and optionally writes the final total state and covariance for comparison.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <vector>

#include "scenelib2_b200.h"

int main(int argc, char **argv) {
  if (argc < 6) {
    std::fprintf(stderr, "usage: %s config.cfg frames.raw width height nframes [out.txt]\n", argv[0]);
    return 2;
  }
  const int W = std::atoi(argv[3]), H = std::atoi(argv[4]), T = std::atoi(argv[5]);
  std::vector<unsigned char> buf((size_t)W * H * T);
  std::ifstream f(argv[2], std::ios::binary);
  if (!f.read(reinterpret_cast<char *>(buf.data()), (std::streamsize)buf.size())) {
    std::fprintf(stderr, "cannot read %s\n", argv[2]);
    return 2;
  }
  try {
    SceneLib2::MonoSLAM *g_monoslam = new SceneLib2::MonoSLAM();
    g_monoslam->Init(argv[1]);
    for (int t = 0; t < T; ++t) {
      cv::Mat frame(H, W, CV_8UC1, buf.data() + (size_t)t * W * H);
      g_monoslam->GoOneStep(frame, true, false);
      std::printf("frame %d visible %d measured %d features %zu\n", t,
                  g_monoslam->number_of_visible_features_,
                  g_monoslam->successful_measurement_vector_size_ / 2,
                  g_monoslam->feature_list_.size());
    }
    g_monoslam->print_robot_state();
    if (argc > 6) {
      Eigen::VectorXd V;
      Eigen::MatrixXd M;
      g_monoslam->construct_total_state(V);
      g_monoslam->construct_total_covariance(M);
      std::ofstream o(argv[6]);
      o.precision(17);
      o << V.size() << "\n";
      for (int i = 0; i < V.size(); ++i) o << V(i) << "\n";
      for (int j = 0; j < M.cols(); ++j)
        for (int i = 0; i < M.rows(); ++i) o << M(i, j) << "\n";
    }
    delete g_monoslam;
  } catch (const std::exception &e) {
    std::fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  return 0;
}
