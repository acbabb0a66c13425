// Compiled by the tests with `g++ -std=c++11 -Wall -Werror`: a C++ consumer written the way the reference's wrapper
// consumes its Rust FFI (crates/optik-cpp/src/lib.cpp:5-31 declares the symbols itself, against its own opaque
// optik::detail::robot and its own optik::SolverConfig POD, include/optik.hpp:13-27) -- NOT against
// include/optik_b200.h.  It links against liboptik_b200.so only through those declarations, which is the drop-in
// claim of INTEGRATION.md section 1.  Eigen is absent in this image, so vectors are std::vector and the pose is a
// column-major double[16] (what Eigen::Isometry3d::matrix().data() points at, lib.cpp:105-120).
//   probe <urdf> <base> <ee> cpu        : constructors, limits, random configuration (no GPU needed)
//   probe <urdf> <base> <ee> gpu <N>    : the protocol of examples/example.cpp:19-42 (N random reachable targets)
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace wrap {
namespace detail { struct robot; }
enum class SolutionMode { kQuality = 1, kSpeed = 2 };
struct SolverConfig {  // the wrapper's POD and defaults, include/optik.hpp:18-27 (max_restarts = 0 => no limit)
  SolutionMode solution_mode = SolutionMode::kSpeed;
  double max_time = 0.1;
  unsigned long max_restarts = 0;
  double tol_f = 1e-6;
  double tol_df = -1.0;
  double tol_dx = -1.0;
  double linear_weight[3]{1.0, 1.0, 1.0};
  double angular_weight[3]{1.0, 1.0, 1.0};
};
}  // namespace wrap

extern "C" {
wrap::detail::robot* optik_robot_from_urdf_file(const char*, const char*, const char*);
void optik_robot_free(wrap::detail::robot*);
void optik_robot_set_parallelism(wrap::detail::robot*, unsigned int);
unsigned int optik_robot_num_positions(const wrap::detail::robot*);
double* optik_robot_joint_limits(const wrap::detail::robot*);
double* optik_robot_random_configuration(const wrap::detail::robot*);
double* optik_robot_joint_jacobian(const wrap::detail::robot*, const double*);
double* optik_robot_fk(const wrap::detail::robot*, const double*);
double* optik_robot_ik(const wrap::detail::robot*, const wrap::SolverConfig*, const double*, const double*);
double* optik_robot_diff_ik(const wrap::detail::robot*, const double*, const double*, const double*);
}

namespace wrap {
class Robot final {  // move-only RAII owner, same method set as optik::Robot (include/optik.hpp:29-105)
 public:
  static Robot FromUrdfFile(const std::string& p, const std::string& b, const std::string& e) {
    return Robot(optik_robot_from_urdf_file(p.c_str(), b.c_str(), e.c_str()));
  }
  Robot(Robot&& o) : inner_(o.inner_) { o.inner_ = nullptr; }
  Robot(const Robot&) = delete;
  Robot& operator=(const Robot&) = delete;
  ~Robot() { if (inner_) optik_robot_free(inner_); }
  void SetParallelism(unsigned n) { optik_robot_set_parallelism(inner_, n); }
  unsigned num_positions() const { return optik_robot_num_positions(inner_); }
  std::vector<double> take(double* p, size_t n) const {  // caller frees, lib.cpp:72,87,100,115,129
    std::vector<double> v(p, p + n);
    free(p);
    return v;
  }
  std::vector<double> JointLimits() const { return take(optik_robot_joint_limits(inner_), 2 * num_positions()); }
  std::vector<double> RandomConfiguration() const { return take(optik_robot_random_configuration(inner_), num_positions()); }
  std::vector<double> JointJacobian(const std::vector<double>& q) const {
    if (q.size() != num_positions()) throw std::runtime_error("dof mismatch");
    return take(optik_robot_joint_jacobian(inner_, q.data()), 6 * num_positions());
  }
  std::vector<double> DoFk(const std::vector<double>& q) const {
    if (q.size() != num_positions()) throw std::runtime_error("dof mismatch");
    return take(optik_robot_fk(inner_, q.data()), 16);
  }
  bool DoIk(const SolverConfig& c, const std::vector<double>& target16, const std::vector<double>& x0, std::vector<double>* q) const {
    if (x0.size() != num_positions()) throw std::runtime_error("dof mismatch");
    double* d = optik_robot_ik(inner_, &c, target16.data(), x0.data());
    if (!d) return false;
    *q = take(d, num_positions());
    return true;
  }
  bool DoDiffIk(const std::vector<double>& x0, const double V[6], const std::vector<double>& vmax, std::vector<double>* v) const {
    double* d = optik_robot_diff_ik(inner_, x0.data(), V, vmax.data());
    if (!d) return false;
    *v = take(d, num_positions());
    return true;
  }

 private:
  explicit Robot(detail::robot* r) : inner_(r) {}
  detail::robot* inner_;
};
}  // namespace wrap

int main(int argc, char** argv) {
  if (argc < 5) return 2;
  static_assert(sizeof(wrap::SolverConfig) == 96, "CSolverConfig layout (crates/optik-cpp/src/lib.rs:10-20)");
  wrap::Robot robot = wrap::Robot::FromUrdfFile(argv[1], argv[2], argv[3]);
  const unsigned n = robot.num_positions();
  const std::vector<double> lim = robot.JointLimits();
  robot.SetParallelism(4);
  bool inside = true;
  for (int k = 0; k < 16; k++) {
    const std::vector<double> q = robot.RandomConfiguration();
    for (unsigned i = 0; i < n; i++) inside = inside && q[i] >= lim[i] && q[i] <= lim[n + i];
  }
  std::printf("n=%u inside=%d\n", n, (int)inside);
  if (std::strcmp(argv[4], "gpu") != 0) return inside ? 0 : 5;

  const int N = argc > 5 ? std::atoi(argv[5]) : 200;
  wrap::SolverConfig config;  // defaults: Speed, max_time 0.1 s, unlimited restarts
  int solved = 0, accurate = 0, dik = 0;
  double total_us = 0.0;
  for (int k = 0; k < N; k++) {
    const std::vector<double> x0 = robot.RandomConfiguration(), qstar = robot.RandomConfiguration();
    const std::vector<double> target = robot.DoFk(qstar);
    std::vector<double> q;
    const auto t0 = std::chrono::steady_clock::now();
    const bool ok = robot.DoIk(config, target, x0, &q);
    total_us += std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
    if (!ok) continue;
    solved++;
    const std::vector<double> got = robot.DoFk(q);
    double err = 0.0;
    for (int i = 0; i < 16; i++) err = std::fmax(err, std::fabs(got[i] - target[i]));
    bool in = true;
    for (unsigned i = 0; i < n; i++) in = in && q[i] >= lim[i] && q[i] <= lim[n + i];
    if (err < 5e-3 && in) accurate++;  // f < 1e-6 bounds the pose error by ~1e-3
    if (n == 6 || n == 7) {
      const double V[6] = {0.1, 0.2, 0.05, 0.3, 0.1, 0.2};
      std::vector<double> v, vmax(n, 1.0);
      if (robot.DoDiffIk(q, V, vmax, &v)) {
        bool vin = true;
        for (unsigned i = 0; i < n; i++) vin = vin && std::fabs(v[i]) <= 1.0 + 1e-9;
        dik += vin;
      }
    }
  }
  const std::vector<double> J = robot.JointJacobian(robot.RandomConfiguration());
  std::printf("solved=%d/%d accurate=%d diff_ik_ok=%d jac=%zu avg_us=%.1f\n", solved, N, accurate, dik, J.size(), total_us / N);
  return 0;
}
