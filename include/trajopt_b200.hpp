// trajopt_b200.hpp — C++ host layer over the C ABI (trajopt_b200.h) under the reference's own names.
//
// The reference is a C++ library; what a caller writes against it for this path is
//   trajopt::ProblemConstructionInfo pci(env);  pci.basic_info...;  pci.cost_infos.push_back(term);  ...
//   auto prob = trajopt::ConstructProblem(pci);                      problem_description.hpp:235-259, 661
//   sco::BasicTrustRegionSQP opt(prob);  opt.setParameters(pci.opt_info);  opt.optimize();  opt.results();
// (trajopt/include/trajopt/problem_description.hpp:123-259, 273-659; trajopt_sco/include/trajopt_sco/optimizers.hpp:25-135).
// This header mirrors those types field by field for the terms the device path implements, for a BATCH of problems
// that share the robot, the term structure and the parameters (what differs per trajectory: initial trajectory,
// Cartesian targets, obstacle set), and flattens them into the POD description of the C ABI.  Differences from the
// reference, all forced by the missing tesseract / Eigen / jsoncpp dependencies:
//   * kinematics come as a RobotModel (URDF joint origins / axes / limits + link collision spheres) instead of a
//     tesseract::environment::Environment; frames are named links of that model;
//   * Eigen::VectorXd -> std::vector<double>, Eigen::Isometry3d -> Pose {xyz, wxyz};
//   * errors are std::runtime_error (PRINT_AND_THROW, trajopt_common/macros.h:90-98); solver failures are status codes;
//   * sco::ModelType is resolved BY NAME (the reference's name table is permuted against its enum).
// Header only; link with -ltrajopt_b200.  Everything lives in namespace trajopt_b200 so that a shim inside the
// reference tree can alias it (namespace tb = trajopt_b200).
#pragma once
#include <cmath>
#include <cstdint>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "trajopt_b200.h"

namespace trajopt_b200 {

using DblVec = std::vector<double>;  // sco::DblVec, trajopt_sco/include/trajopt_sco/sco_common.hpp:17
using IntVec = std::vector<int>;

namespace sco {

// trajopt_sco/include/trajopt_sco/optimizers.hpp:25-33
enum OptStatus : int {
  OPT_CONVERGED = TB200_OPT_CONVERGED,
  OPT_SCO_ITERATION_LIMIT = TB200_OPT_SCO_ITERATION_LIMIT,
  OPT_PENALTY_ITERATION_LIMIT = TB200_OPT_PENALTY_ITERATION_LIMIT,
  OPT_TIME_LIMIT = TB200_OPT_TIME_LIMIT,
  OPT_FAILED = TB200_OPT_FAILED,
  INVALID = TB200_OPT_INVALID
};

// optimizers.hpp:92-135 (same names, same defaults)
struct BasicTrustRegionSQPParameters {
  double improve_ratio_threshold = 0.25;
  double min_trust_box_size = 1e-4;
  double min_approx_improve = 1e-4;
  double min_approx_improve_frac = std::numeric_limits<double>::lowest();
  int max_iter = 50;
  double trust_shrink_ratio = 0.1;
  double trust_expand_ratio = 1.5;
  double cnt_tolerance = 1e-4;
  double max_merit_coeff_increases = 5;
  int max_qp_solver_failures = 3;
  double merit_coeff_increase_ratio = 10;
  double initial_merit_error_coeff = 10;
  bool inflate_constraints_individually = true;
  double trust_box_size = 1e-1;
};

// optimizers.hpp:40-59
struct OptResults {
  DblVec x;  // [T*D], trajToDblVec order
  OptStatus status = INVALID;
  double total_cost = 0;
  DblVec cost_vals, cnt_viols;
  int n_func_evals = 0, n_qp_solves = 0;
};

// solver_interface.hpp:229-236.  Only the OSQP-equivalent device solver exists here.
enum class ModelType { GUROBI, OSQP, QPOASES, BPMPD, AUTO_SOLVER };
inline ModelType modelTypeFromName(const std::string& name) {
  if (name == "GUROBI") return ModelType::GUROBI;
  if (name == "OSQP") return ModelType::OSQP;
  if (name == "QPOASES") return ModelType::QPOASES;
  if (name == "BPMPD") return ModelType::BPMPD;
  if (name == "AUTO_SOLVER") return ModelType::AUTO_SOLVER;
  throw std::runtime_error("invalid solver name:\"" + name + "\"");  // solver_interface.cpp:243-258
}

}  // namespace sco

namespace trajopt {

// problem_description.hpp:34-41
enum TermType : int { TT_INVALID = 0, TT_COST = 0x1, TT_CNT = 0x2, TT_USE_TIME = 0x4 };

struct Pose {  // Eigen::Isometry3d stand-in
  double xyz[3] = {0, 0, 0};
  double wxyz[4] = {1, 0, 0, 0};
};

// What ConstructProblem takes from pci.kin / pci.env (problem_description.cpp:410-460, 553-592): the manipulator
// chain, its limits and the collision geometry of its links.
struct RobotModel {
  struct Joint {
    std::string child_link;  // name of the frame this joint creates
    int parent = -1;         // index of the parent joint/frame, -1 = scene root
    int type = TB200_JOINT_FIXED;
    int q_index = -1;        // column of the trajectory
    Pose origin;
    double axis[3] = {0, 0, 1};
  };
  struct Sphere {
    std::string link;
    double center[3] = {0, 0, 0};
    double radius = 0;
  };
  std::vector<Joint> joints;  // topologically ordered
  DblVec lower, upper;        // kin->getLimits()
  std::vector<Sphere> spheres;
  int numJoints() const { return static_cast<int>(lower.size()); }
  int linkIndex(const std::string& name) const {
    for (size_t i = 0; i < joints.size(); ++i)
      if (joints[i].child_link == name) return static_cast<int>(i);
    throw std::runtime_error("link \"" + name + "\" is not part of the manipulator model");
  }
};

// problem_description.hpp:123-157 (use_time / dt limits are not on the device path)
struct BasicInfo {
  int n_steps = -1;
  std::string manip;
  IntVec fixed_timesteps;
  IntVec fixed_dofs;
  sco::ModelType convex_solver = sco::ModelType::AUTO_SOLVER;
  bool use_time = false;
};

// problem_description.hpp:162-185.  data: JOINT_INTERPOLATED -> end states [B][D]; GIVEN_TRAJ -> [B][T][D];
// start: the current joint values of every problem [B][D] (pci.env->getCurrentJointValues in the reference).
struct InitInfo {
  enum Type : std::uint8_t { STATIONARY, JOINT_INTERPOLATED, GIVEN_TRAJ };
  Type type = STATIONARY;
  DblVec data;
  DblVec start;
};

struct ProblemConstructionInfo;
struct Flat;  // accumulates the POD description

// problem_description.hpp:199-230
struct TermInfo {
  using Ptr = std::shared_ptr<TermInfo>;
  std::string name;
  int term_type = TT_INVALID;
  virtual void hatch(Flat& flat, const ProblemConstructionInfo& pci) const = 0;
  virtual ~TermInfo() = default;
};

struct Flat {
  std::vector<tb200_term> terms;
  DblVec cart_targets;  // [B][n_slots][7]
  int n_slots = 0, batch = 0;
  int addTargets(const std::vector<Pose>& per_traj) {  // one static target per trajectory -> a slot
    if (static_cast<int>(per_traj.size()) != batch) throw std::runtime_error("cart_pose: one target pose per trajectory is required");
    DblVec grown(static_cast<size_t>(batch) * (n_slots + 1) * 7);
    for (int b = 0; b < batch; ++b) {
      for (int s = 0; s < n_slots; ++s)
        for (int k = 0; k < 7; ++k) grown[(static_cast<size_t>(b) * (n_slots + 1) + s) * 7 + k] = cart_targets[(static_cast<size_t>(b) * n_slots + s) * 7 + k];
      double* o = &grown[(static_cast<size_t>(b) * (n_slots + 1) + n_slots) * 7];
      for (int k = 0; k < 3; ++k) o[k] = per_traj[b].xyz[k];
      for (int k = 0; k < 4; ++k) o[3 + k] = per_traj[b].wxyz[k];
    }
    cart_targets.swap(grown);
    return n_slots++;
  }
};

namespace detail {
inline tb200_term blankTerm(int kind, int term_type, const std::string& name) {
  if (term_type != TT_COST && term_type != TT_CNT) throw std::runtime_error(name + ": term_type must be TT_COST or TT_CNT");
  tb200_term t{};
  t.kind = kind;
  t.role = (term_type == TT_COST) ? TB200_ROLE_COST : TB200_ROLE_CNT;
  return t;
}
inline void fill(double* dst, const DblVec& src, int n, double def, const std::string& what) {
  if (!src.empty() && static_cast<int>(src.size()) != n) throw std::runtime_error(what + " has the wrong size");
  for (int i = 0; i < n; ++i) dst[i] = src.empty() ? def : src[i];
}
}  // namespace detail

// JointPos/Vel/AccTermInfo, problem_description.hpp:430-560; hatch step clamping problem_description.cpp:1078-1106,
// 1197-1224, 1393-1421 (last_step <= -1: to the end; velocity needs two steps, acceleration three).
struct JointTermInfoBase : TermInfo {
  DblVec coeffs, targets, upper_tols, lower_tols;
  int first_step = 0, last_step = -1;
  int order_ = 0;
  void hatch(Flat& flat, const ProblemConstructionInfo& pci) const override;
};
struct JointPosTermInfo : JointTermInfoBase { JointPosTermInfo() { order_ = 0; name = "joint_pos"; } };
struct JointVelTermInfo : JointTermInfoBase { JointVelTermInfo() { order_ = 1; name = "joint_vel"; } };
struct JointAccTermInfo : JointTermInfoBase { JointAccTermInfo() { order_ = 2; name = "joint_acc"; } };

// CartPoseTermInfo, problem_description.hpp:330-376 (static target; tolerances / error_function not on the device path)
struct CartPoseTermInfo : TermInfo {
  int timestep = 0;
  double pos_coeffs[3] = {1, 1, 1}, rot_coeffs[3] = {1, 1, 1};
  std::string source_frame;
  Pose source_frame_offset;
  std::vector<Pose> target;  // target_frame * target_frame_offset in the scene root, one per trajectory
  CartPoseTermInfo() { name = "cart_pose"; }
  void hatch(Flat& flat, const ProblemConstructionInfo& pci) const override;
};

// CartVelTermInfo, problem_description.hpp:383-398
struct CartVelTermInfo : TermInfo {
  int first_step = -1, last_step = -1;
  std::string link;
  double max_displacement = 0;
  CartVelTermInfo() { name = "cart_vel"; }
  void hatch(Flat& flat, const ProblemConstructionInfo& pci) const override;
};

// CollisionTermInfo, problem_description.hpp:600-659 + trajopt_common TrajOptCollisionConfig (collision_types.h:120-170)
struct CollisionTermInfo : TermInfo {
  int first_step = 0, last_step = -1;
  IntVec fixed_steps;
  int evaluator_type = TB200_COLL_DISCRETE;  // CollisionEvaluatorType
  double collision_margin = 0.025;           // "dist_pen" / safety margin
  double collision_coeff = 20;
  double collision_margin_buffer = 0.01;
  double longest_valid_segment_length = 0.005;
  CollisionTermInfo() { name = "collision"; }
  void hatch(Flat& flat, const ProblemConstructionInfo& pci) const override;
};

// problem_description.hpp:235-259
struct ProblemConstructionInfo {
  BasicInfo basic_info;
  sco::BasicTrustRegionSQPParameters opt_info;
  std::vector<TermInfo::Ptr> cost_infos, cnt_infos;
  InitInfo init_info;
  std::shared_ptr<const RobotModel> kin;  // stands in for pci.kin + pci.env
  int batch = 1;
  DblVec obstacles;  // static world spheres (x, y, z, r): [B][O][4] or [O][4]
  int n_obstacles = 0;
  bool obstacles_per_problem = true;
};

inline void JointTermInfoBase::hatch(Flat& flat, const ProblemConstructionInfo& pci) const {
  const int T = pci.basic_info.n_steps, D = pci.kin->numJoints();
  int first = first_step, last = last_step;
  if (last <= -1) last = T - 1;
  if ((T - 1 - order_) <= first) first = T - 1 - order_;
  if ((T - 1) <= last) last = T - 1;
  if (order_ > 0 && last == first) last += order_;
  if (last < first) std::swap(first, last);
  tb200_term t = detail::blankTerm(TB200_TERM_JOINT_POS + order_, term_type, name);
  t.first_step = first;
  t.last_step = last;
  detail::fill(t.coeffs, coeffs, D, 1.0, name + ": coeffs");
  detail::fill(t.targets, targets, D, 0.0, name + ": targets");
  detail::fill(t.upper_tols, upper_tols, D, 0.0, name + ": upper_tols");
  detail::fill(t.lower_tols, lower_tols, D, 0.0, name + ": lower_tols");
  flat.terms.push_back(t);
}
inline void CartPoseTermInfo::hatch(Flat& flat, const ProblemConstructionInfo& pci) const {
  tb200_term t = detail::blankTerm(TB200_TERM_CART_POSE, term_type, name);
  t.first_step = t.last_step = timestep;
  t.link = pci.kin->linkIndex(source_frame);
  for (int k = 0; k < 3; ++k) {
    t.source_offset[k] = source_frame_offset.xyz[k];
    t.pos_coeffs[k] = pos_coeffs[k];
    t.rot_coeffs[k] = rot_coeffs[k];
  }
  for (int k = 0; k < 4; ++k) t.source_offset[3 + k] = source_frame_offset.wxyz[k];
  t.target_pose[3] = 1.0;  // unused: the target is read from the per-trajectory slot
  t.target_slot = flat.addTargets(target);
  flat.terms.push_back(t);
}
inline void CartVelTermInfo::hatch(Flat& flat, const ProblemConstructionInfo& pci) const {
  tb200_term t = detail::blankTerm(TB200_TERM_CART_VEL, term_type, name);
  const int T = pci.basic_info.n_steps;
  t.first_step = first_step < 0 ? 0 : first_step;
  t.last_step = (last_step < 0 || last_step > T - 2) ? T - 2 : last_step;  // pair t = (t, t+1)
  t.link = pci.kin->linkIndex(link);
  t.target_slot = -1;
  t.max_displacement = max_displacement;
  flat.terms.push_back(t);
}
inline void CollisionTermInfo::hatch(Flat& flat, const ProblemConstructionInfo& pci) const {
  tb200_term t = detail::blankTerm(TB200_TERM_COLLISION, term_type, name);
  const int T = pci.basic_info.n_steps;
  t.first_step = first_step;
  t.last_step = (last_step <= -1 || last_step > T - 1) ? T - 1 : last_step;
  t.evaluator_type = evaluator_type;
  if (fixed_steps.size() > 8) throw std::runtime_error(name + ": more than 8 fixed_steps");
  t.n_fixed_steps = static_cast<int>(fixed_steps.size());
  for (size_t i = 0; i < fixed_steps.size(); ++i) t.fixed_steps[i] = fixed_steps[i];
  t.margin = collision_margin;
  t.coeff = collision_coeff;
  t.margin_buffer = collision_margin_buffer;
  t.longest_valid_segment_length = longest_valid_segment_length;
  flat.terms.push_back(t);
}

// The flattened description with everything it points to (kept alive together).
struct FlatProblem {
  tb200_problem_desc desc{};
  std::vector<tb200_segment> segments;
  std::vector<tb200_sphere> spheres;
  DblVec lower, upper, init_traj, cart_targets, obstacles;
  std::vector<tb200_term> terms;
  std::vector<int32_t> fixed_timesteps, fixed_dofs;
  int n_costs_terms = 0;
};

// The part of ConstructProblem (problem_description.cpp:410-542) that does not need the device: checks, initial
// trajectory (InitInfo, :330-408), term hatching into the POD description (cost_infos first, then cnt_infos).
inline std::shared_ptr<FlatProblem> FlattenProblem(const ProblemConstructionInfo& pci) {
  if (!pci.kin) throw std::runtime_error("ProblemConstructionInfo: no kinematics");
  const RobotModel& kin = *pci.kin;
  const int T = pci.basic_info.n_steps, D = kin.numJoints(), B = pci.batch;
  if (T < 1) throw std::runtime_error("basic_info.n_steps must be positive");
  if (pci.basic_info.use_time) throw std::runtime_error("use_time problems are not on the device path");
  if (pci.basic_info.convex_solver != sco::ModelType::OSQP && pci.basic_info.convex_solver != sco::ModelType::AUTO_SOLVER)
    throw std::runtime_error("the device path implements the OSQP-equivalent solver only");
  auto fp = std::make_shared<FlatProblem>();
  for (const RobotModel::Joint& j : kin.joints) {
    tb200_segment s{};
    s.parent = j.parent;
    s.joint_type = j.type;
    s.q_index = j.q_index;
    for (int k = 0; k < 3; ++k) { s.origin_xyz[k] = j.origin.xyz[k]; s.axis[k] = j.axis[k]; }
    for (int k = 0; k < 4; ++k) s.origin_wxyz[k] = j.origin.wxyz[k];
    fp->segments.push_back(s);
  }
  for (const RobotModel::Sphere& sp : kin.spheres) {
    tb200_sphere s{};
    s.segment = kin.linkIndex(sp.link);
    for (int k = 0; k < 3; ++k) s.center[k] = sp.center[k];
    s.radius = sp.radius;
    fp->spheres.push_back(s);
  }
  fp->lower = kin.lower;
  fp->upper = kin.upper;
  // ---- InitInfo (problem_description.cpp:330-408)
  const InitInfo& ii = pci.init_info;
  fp->init_traj.assign(static_cast<size_t>(B) * T * D, 0.0);
  if (ii.type == InitInfo::GIVEN_TRAJ) {
    if (ii.data.size() != fp->init_traj.size()) throw std::runtime_error("Initial trajectory has the wrong size");
    fp->init_traj = ii.data;
  } else {
    if (ii.start.size() != static_cast<size_t>(B) * D) throw std::runtime_error("InitInfo.start: one joint state per problem is required");
    if (ii.type == InitInfo::JOINT_INTERPOLATED && ii.data.size() != static_cast<size_t>(B) * D)
      throw std::runtime_error("init_info.data has the wrong size for JOINT_INTERPOLATED");
    for (int b = 0; b < B; ++b)
      for (int t = 0; t < T; ++t)
        for (int d = 0; d < D; ++d) {
          const double s = ii.start[static_cast<size_t>(b) * D + d];
          double v = s;
          if (ii.type == InitInfo::JOINT_INTERPOLATED && T > 1) {  // LinSpaced per joint, :351-355
            const double e = ii.data[static_cast<size_t>(b) * D + d];
            v = (t == T - 1) ? e : s + t * ((e - s) / (T - 1));
          }
          fp->init_traj[(static_cast<size_t>(b) * T + t) * D + d] = v;
        }
  }
  // ---- terms: cost_infos first, then cnt_infos (problem_description.cpp:462-484)
  Flat flat;
  flat.batch = B;
  for (const auto& ti : pci.cost_infos) {
    if (ti->term_type != TT_COST) throw std::runtime_error(ti->name + ": a cost_info must have term_type TT_COST");
    ti->hatch(flat, pci);
  }
  fp->n_costs_terms = static_cast<int>(flat.terms.size());
  for (const auto& ti : pci.cnt_infos) {
    if (ti->term_type != TT_CNT) throw std::runtime_error(ti->name + ": a cnt_info must have term_type TT_CNT");
    ti->hatch(flat, pci);
  }
  fp->terms = flat.terms;
  fp->cart_targets = flat.cart_targets;
  {
    const size_t want = static_cast<size_t>(pci.obstacles_per_problem ? B : 1) * static_cast<size_t>(pci.n_obstacles) * 4;
    if (pci.obstacles.size() != want)
      throw std::runtime_error("obstacles has " + std::to_string(pci.obstacles.size()) + " values, expected " + std::to_string(want) +
                               " ([B or 1][n_obstacles][4])");
  }
  fp->obstacles = pci.obstacles;
  fp->fixed_timesteps.assign(pci.basic_info.fixed_timesteps.begin(), pci.basic_info.fixed_timesteps.end());
  fp->fixed_dofs.assign(pci.basic_info.fixed_dofs.begin(), pci.basic_info.fixed_dofs.end());
  tb200_problem_desc& d = fp->desc;
  d.robot.n_dof = D;
  d.robot.n_segments = static_cast<int32_t>(fp->segments.size());
  d.robot.segments = fp->segments.data();
  d.robot.lower = fp->lower.data();
  d.robot.upper = fp->upper.data();
  d.robot.n_spheres = static_cast<int32_t>(fp->spheres.size());
  d.robot.spheres = fp->spheres.data();
  d.n_steps = T;
  d.batch = B;
  d.n_terms = static_cast<int32_t>(fp->terms.size());
  d.terms = fp->terms.data();
  d.n_fixed_timesteps = static_cast<int32_t>(fp->fixed_timesteps.size());
  d.fixed_timesteps = fp->fixed_timesteps.data();
  d.n_fixed_dofs = static_cast<int32_t>(fp->fixed_dofs.size());
  d.fixed_dofs = fp->fixed_dofs.data();
  d.n_cart_targets = flat.n_slots;
  d.init_traj = fp->init_traj.data();
  d.cart_targets = fp->cart_targets.empty() ? nullptr : fp->cart_targets.data();
  d.n_obstacles = pci.n_obstacles;
  d.obstacles_per_traj = pci.obstacles_per_problem ? 1 : 0;
  d.obstacles = fp->obstacles.empty() ? nullptr : fp->obstacles.data();
  tb200_default_qp_settings(&d.qp);  // OSQPModelConfig::setDefaultOSQPSettings, osqp_interface.cpp:78-90
  const sco::BasicTrustRegionSQPParameters& p = pci.opt_info;
  d.sqp.improve_ratio_threshold = p.improve_ratio_threshold;
  d.sqp.min_trust_box_size = p.min_trust_box_size;
  d.sqp.min_approx_improve = p.min_approx_improve;
  d.sqp.min_approx_improve_frac = p.min_approx_improve_frac;
  d.sqp.max_iter = p.max_iter;
  d.sqp.max_qp_solver_failures = p.max_qp_solver_failures;
  d.sqp.trust_shrink_ratio = p.trust_shrink_ratio;
  d.sqp.trust_expand_ratio = p.trust_expand_ratio;
  d.sqp.cnt_tolerance = p.cnt_tolerance;
  d.sqp.max_merit_coeff_increases = p.max_merit_coeff_increases;
  d.sqp.merit_coeff_increase_ratio = p.merit_coeff_increase_ratio;
  d.sqp.initial_merit_error_coeff = p.initial_merit_error_coeff;
  d.sqp.trust_box_size = p.trust_box_size;
  d.sqp.inflate_constraints_individually = p.inflate_constraints_individually ? 1 : 0;
  return fp;
}

// The batched TrajOptProb: owns the device handle (problem_description.hpp:68-107 for one problem).
class TrajOptProb {
public:
  using Ptr = std::shared_ptr<TrajOptProb>;
  TrajOptProb(std::shared_ptr<FlatProblem> flat, int device) : flat_(std::move(flat)) {
    if (tb200_problem_create(&flat_->desc, device, &handle_) != TB200_OK) throw std::runtime_error(tb200_last_error());
    if (tb200_problem_layout(handle_, &layout_) != TB200_OK) throw std::runtime_error(tb200_last_error());
  }
  ~TrajOptProb() { tb200_problem_destroy(handle_); }
  TrajOptProb(const TrajOptProb&) = delete;
  TrajOptProb& operator=(const TrajOptProb&) = delete;
  int GetNumSteps() const { return flat_->desc.n_steps; }
  int GetNumDOF() const { return flat_->desc.robot.n_dof; }
  int GetBatch() const { return flat_->desc.batch; }
  const DblVec& GetInitTraj() const { return flat_->init_traj; }
  const tb200_sqp_params& sqpParams() const { return flat_->desc.sqp; }  // = pci.opt_info
  int getNumCosts() const { return layout_.n_costs; }
  int getNumConstraints() const { return layout_.n_cnts; }
  tb200_problem* handle() const { return handle_; }

private:
  std::shared_ptr<FlatProblem> flat_;
  tb200_problem* handle_ = nullptr;
  tb200_layout layout_{};
};

// trajopt::ConstructProblem(pci), problem_description.hpp:661 / problem_description.cpp:410-542
inline TrajOptProb::Ptr ConstructProblem(const ProblemConstructionInfo& pci, int device = 0) {
  return std::make_shared<TrajOptProb>(FlattenProblem(pci), device);
}

// BasicTrustRegionSQP::optimize() for every problem of the batch (optimizers.cpp:699-991) with the given parameters: what
// `sco::BasicTrustRegionSQP opt(prob); opt.getParameters() = params; opt.initialize(...); opt.optimize(); opt.results()`
// returns, per problem.
inline std::vector<sco::OptResults> OptimizeWithParams(TrajOptProb& prob, const tb200_sqp_params& params) {
  if (tb200_problem_set_sqp_params(prob.handle(), &params) != TB200_OK) throw std::runtime_error(tb200_last_error());
  const size_t B = prob.GetBatch(), N = static_cast<size_t>(prob.GetNumSteps()) * prob.GetNumDOF();
  const size_t nc = prob.getNumCosts(), nk = prob.getNumConstraints();
  DblVec x(B * N), total(B), cv(B * (nc ? nc : 1)), kv(B * (nk ? nk : 1));
  std::vector<int32_t> status(B), nqp(B), nfe(B);
  tb200_results r{};
  r.x = x.data(); r.status = status.data(); r.total_cost = total.data();
  r.cost_vals = nc ? cv.data() : nullptr; r.cnt_viols = nk ? kv.data() : nullptr;
  r.n_qp_solves = nqp.data(); r.n_func_evals = nfe.data();
  if (tb200_solve_batch(prob.handle(), &r) != TB200_OK) throw std::runtime_error(tb200_last_error());
  std::vector<sco::OptResults> out(B);
  for (size_t b = 0; b < B; ++b) {
    out[b].x.assign(x.begin() + b * N, x.begin() + (b + 1) * N);
    out[b].status = static_cast<sco::OptStatus>(status[b]);
    out[b].total_cost = total[b];
    out[b].cost_vals.assign(cv.begin() + b * nc, cv.begin() + (b + 1) * nc);
    out[b].cnt_viols.assign(kv.begin() + b * nk, kv.begin() + (b + 1) * nk);
    out[b].n_qp_solves = nqp[b];
    out[b].n_func_evals = nfe[b];
  }
  return out;
}
// ... with the problem description's own parameters (pci.opt_info)
inline std::vector<sco::OptResults> OptimizeWithParams(TrajOptProb& prob) { return OptimizeWithParams(prob, prob.sqpParams()); }

// trajopt::OptimizeProblem(prob) (problem_description.hpp:665, problem_description.cpp:392-408).  The reference's function
// does NOT run with pci.opt_info: it builds a fresh BasicTrustRegionSQP and overrides four parameters (max_iter 40,
// min_approx_improve_frac 1e-3, improve_ratio_threshold 0.2, initial_merit_error_coeff 20) on top of the optimizer's
// DEFAULTS.  Reproduced here, so that the same description gives the same iterates; OptimizeWithParams is the entry point
// that honours pci.opt_info.
inline std::vector<sco::OptResults> OptimizeProblem(TrajOptProb& prob) {
  tb200_sqp_params p;
  tb200_default_sqp_params(&p);
  p.max_iter = 40;
  p.min_approx_improve_frac = .001;
  p.improve_ratio_threshold = .2;
  p.initial_merit_error_coeff = 20;
  return OptimizeWithParams(prob, p);
}

}  // namespace trajopt
}  // namespace trajopt_b200
