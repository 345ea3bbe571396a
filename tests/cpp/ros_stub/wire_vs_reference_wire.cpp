// Other half of oracle/_ref/dpgo_ros_wire_vs_reference: include/dpgo_ros_wire/wire.h (this repo's ROS-free mirror of the
// wire side) against the reference's own codecs, value for value.  wire.h lives in namespace dpgo_ros like the code it
// mirrors; it is renamed here so that both can sit in one binary.  TEST INFRASTRUCTURE.
#define dpgo_ros dpgo_ros_wire_ns
#include "dpgo_ros_wire/wire.h"
#undef dpgo_ros

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

extern "C" {
void ref_matrix_to_msg(const double *colmajor, int rows, int cols, unsigned *orows, unsigned *ocols, double *values);
void ref_matrix_from_msg(int rows, int cols, const double *values, double *colmajor);
void ref_status_roundtrip(unsigned id, int state, unsigned instance, unsigned iteration, int ready, double rel_change,
                          double *msg_fields, double *back_fields);
void ref_measurement_roundtrip(const double *R_rowmajor, const double *t, double *R_out_rowmajor, double *t_out,
                               double *kappa_tau_fixed, int odometry);
}

#define REQUIRE(c)                                                      \
  do {                                                                  \
    if (!(c)) {                                                         \
      std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #c);        \
      return 1;                                                         \
    }                                                                   \
  } while (0)

int main(int argc, char **argv) {
  namespace wire = dpgo_ros_wire_ns;
  std::mt19937 rng(11);
  std::uniform_real_distribution<double> u(-10, 10);
  // ---- MatrixMsg: rows / cols / ROW-major values, bit for bit
  for (auto shape : {std::make_pair(5, 4), std::make_pair(3, 3), std::make_pair(5, 3), std::make_pair(8, 4), std::make_pair(3, 1)}) {
    const int r = shape.first, c = shape.second;
    DPGO::Matrix M(r, c);
    for (int j = 0; j < c; ++j)
      for (int i = 0; i < r; ++i) M(i, j) = u(rng);
    unsigned rr = 0, cc = 0;
    std::vector<double> ref((size_t)r * c);
    ref_matrix_to_msg(M.data(), r, c, &rr, &cc, ref.data());
    const wire::MatrixMsg mine = wire::MatrixToMsg(M);
    REQUIRE(mine.rows == rr && mine.cols == cc && mine.values.size() == ref.size());
    for (size_t k = 0; k < ref.size(); ++k) REQUIRE(mine.values[k] == ref[k]);
    // the pose-buffer shortcut the device path uses (column-major r x 4 block -> message) must say the same
    const wire::MatrixMsg fromBuffer = wire::PoseBufferToMsg(M.data(), (unsigned)r, (unsigned)c);
    REQUIRE(fromBuffer.values == mine.values);
    // and decoding: the reference reads my message, I read the reference's
    std::vector<double> back((size_t)r * c);
    ref_matrix_from_msg(r, c, mine.values.data(), back.data());
    for (int k = 0; k < r * c; ++k) REQUIRE(back[k] == M.data()[k]);
    wire::MatrixMsg theirs;
    theirs.rows = (uint16_t)rr;
    theirs.cols = (uint16_t)cc;
    theirs.values = ref;
    const DPGO::Matrix Mb = wire::MatrixFromMsg(theirs);
    for (int k = 0; k < r * c; ++k) REQUIRE(Mb.data()[k] == M.data()[k]);
  }
  // ---- Status: float32 relative change on the wire (msg/Status.msg:11)
  for (double rel : {0.5, 0.1, 1.0 / 3.0, 2.718281828459045e-3, 123456.789}) {
    double m[6], b[6];
    ref_status_roundtrip(3, 2, 7, 41, 1, rel, m, b);
    const DPGO::PGOAgentStatus st(3, DPGO::PGOAgentState::INITIALIZED, 7, 41, true, rel);
    const wire::Status msg = wire::statusToMsg(st);
    REQUIRE(msg.robot_id == m[0] && msg.state == m[1] && msg.instance_number == m[2] && msg.iteration_number == m[3] &&
            (double)msg.ready_to_terminate == m[4] && (double)msg.relative_change == m[5]);
    const DPGO::PGOAgentStatus bk = wire::statusFromMsg(msg);
    REQUIRE(bk.agentID == b[0] && (double)bk.state == b[1] && bk.instanceNumber == b[2] && bk.iterationNumber == b[3] &&
            (double)bk.readyToTerminate == b[4] && bk.relativeChange == b[5]);
    REQUIRE((double)(float)rel == b[5]);
  }
  // ---- what a measurement looks like after the PoseGraphEdge message: kappa = 10000, tau = 100, odometry fixed
  {
    const double c = std::cos(0.3), s = std::sin(0.3);
    const double R[9] = {c, -s, 0, s, c, 0, 0, 0, 1}, t[3] = {-1.5, 2.1, 3.9};
    double Ro[9], to[3], ktf[3];
    ref_measurement_roundtrip(R, t, Ro, to, ktf, 0);
    for (int k = 0; k < 9; ++k) REQUIRE(std::fabs(Ro[k] - R[k]) < 1e-12);
    for (int k = 0; k < 3; ++k) REQUIRE(to[k] == t[k]);
    REQUIRE(ktf[0] == 10000 && ktf[1] == 100 && ktf[2] == 0);
    ref_measurement_roundtrip(R, t, Ro, to, ktf, 1);
    REQUIRE(ktf[2] == 1);
  }
  // ---- the CSV header of the per-round log (argv[1] = a log written by the unmodified wrapper)
  if (argc > 1) {
    const std::string mine = std::string(argv[1]) + ".wire";
    wire::IterationLog log;
    REQUIRE(log.open(mine));
    log.logIteration(1, 0, 5, 12, 500, 123456, 0.004, 1.5, 0.25);
    std::ifstream a(argv[1]), b(mine);
    std::string ha, hb, ra, rb;
    std::getline(a, ha);
    std::getline(b, hb);
    REQUIRE(ha == hb);
    std::getline(a, ra);
    std::getline(b, rb);
    REQUIRE(std::count(ra.begin(), ra.end(), ',') == std::count(rb.begin(), rb.end(), ','));
  }
  std::printf("wire vs reference ok\n");
  return 0;
}
