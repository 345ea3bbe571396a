/* DPGO/PGOLogger.h -- PGOLogger::loadMeasurements as the dataset publisher uses it
 * (src/PGODatasetPublisherNode.cpp:168-169): per-robot CSV with the header
 * robot_src,pose_src,robot_dst,pose_dst,qx,qy,qz,qw,tx,ty,tz,kappa,tau,is_known_inlier,weight
 * (data/tunnels/robot0/measurements.csv:1). */
#ifndef DPGO_SHIM_PGOLOGGER_H
#define DPGO_SHIM_PGOLOGGER_H
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "DPGO/RelativeSEMeasurement.h"

namespace DPGO {
class PGOLogger {
 public:
  explicit PGOLogger(std::string logDir = "") : logDirectory(std::move(logDir)) {}
  // load_weight = false: weights come back as 1 (the publisher calls it that way, and statically)
  static std::vector<RelativeSEMeasurement> loadMeasurements(const std::string &filename, bool load_weight = false) {
    std::vector<RelativeSEMeasurement> out;
    std::ifstream in(filename);
    if (!in) return out;
    std::string line;
    std::getline(in, line);  // header
    while (std::getline(in, line)) {
      if (line.empty()) continue;
      std::istringstream ss(line);
      std::string tok;
      double v[15];
      int k = 0;
      while (k < 15 && std::getline(ss, tok, ',')) v[k++] = std::stod(tok);
      if (k < 13) continue;
      const double nq = std::sqrt(v[4] * v[4] + v[5] * v[5] + v[6] * v[6] + v[7] * v[7]);
      const double x = v[4] / nq, y = v[5] / nq, z = v[6] / nq, w = v[7] / nq;
      Matrix R(3, 3), t(3, 1);
      R << 1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
           2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
           2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y);
      t << v[8], v[9], v[10];
      RelativeSEMeasurement m((size_t)v[0], (size_t)v[2], (size_t)v[1], (size_t)v[3], R, t, v[11], v[12]);
      if (k >= 14) m.fixedWeight = v[13] != 0.0;   // known inliers keep weight 1
      if (load_weight && k >= 15) m.weight = v[14];
      out.push_back(m);
    }
    return out;
  }

 private:
  std::string logDirectory;
};
}  // namespace DPGO
#endif
