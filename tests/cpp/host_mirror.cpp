// Exercises the C++ host layer (include/trajopt_b200.hpp) the way a caller of the reference writes a problem:
// ProblemConstructionInfo + TermInfo objects -> ConstructProblem -> OptimizeProblem.  Reads the robot and the
// per-problem data from a text file written by tests/test_cpp_host.py, builds the configs[2] description and either
// dumps the flattened POD description (mode "dump", no device needed) or solves it (mode "solve").
#include <cstdio>
#include <fstream>
#include <iostream>

#include <sstream>

#include "trajopt_b200_json.hpp"

namespace tb = trajopt_b200;
using namespace tb::trajopt;

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  std::ifstream in(argv[1]);
  const std::string mode = argv[2];
  int B, T, D, nseg;
  in >> B >> T >> D >> nseg;
  auto kin = std::make_shared<RobotModel>();
  for (int s = 0; s < nseg; ++s) {
    RobotModel::Joint j;
    in >> j.parent >> j.type >> j.q_index >> j.origin.xyz[0] >> j.origin.xyz[1] >> j.origin.xyz[2] >> j.origin.wxyz[0] >>
        j.origin.wxyz[1] >> j.origin.wxyz[2] >> j.origin.wxyz[3] >> j.axis[0] >> j.axis[1] >> j.axis[2] >> j.child_link;
    kin->joints.push_back(j);
  }
  kin->lower.resize(D);
  kin->upper.resize(D);
  for (double& v : kin->lower) in >> v;
  for (double& v : kin->upper) in >> v;
  int nsph;
  in >> nsph;
  for (int s = 0; s < nsph; ++s) {
    RobotModel::Sphere sp;
    in >> sp.link >> sp.center[0] >> sp.center[1] >> sp.center[2] >> sp.radius;
    kin->spheres.push_back(sp);
  }
  std::string tool;
  in >> tool;
  ProblemConstructionInfo pci;
  pci.kin = kin;
  pci.batch = B;
  pci.basic_info.n_steps = T;
  pci.basic_info.manip = "right_arm";
  pci.basic_info.fixed_timesteps = {0};
  pci.basic_info.convex_solver = tb::sco::modelTypeFromName("OSQP");
  pci.init_info.type = InitInfo::JOINT_INTERPOLATED;
  pci.init_info.start.resize(static_cast<size_t>(B) * D);
  pci.init_info.data.resize(static_cast<size_t>(B) * D);
  for (double& v : pci.init_info.start) in >> v;
  for (double& v : pci.init_info.data) in >> v;
  std::vector<Pose> goals(B);
  for (Pose& g : goals) in >> g.xyz[0] >> g.xyz[1] >> g.xyz[2] >> g.wxyz[0] >> g.wxyz[1] >> g.wxyz[2] >> g.wxyz[3];
  in >> pci.n_obstacles;
  pci.obstacles.resize(static_cast<size_t>(B) * pci.n_obstacles * 4);
  for (double& v : pci.obstacles) in >> v;
  if (!in) { std::fprintf(stderr, "bad input file\n"); return 2; }

  if (mode == "json") {  // the description comes from a JSON document in the reference's schema (argv[3])
    try {
      std::ifstream jf(argv[3]);
      std::stringstream ss;
      ss << jf.rdbuf();
      ProblemConstructionInfo jp;
      jp.kin = kin;
      jp.batch = B;
      jp.obstacles = pci.obstacles;
      jp.n_obstacles = pci.n_obstacles;
      fromJson(jp, tb::json::parse(ss.str()));
      jp.init_info.start = pci.init_info.start;  // the environment's current joint values, one state per problem
      auto fp = FlattenProblem(jp);
      std::printf("n_terms %d n_cart_targets %d n_fixed %d max_iter %d trust %.17g\n", fp->desc.n_terms, fp->desc.n_cart_targets,
                  fp->desc.n_fixed_timesteps, fp->desc.sqp.max_iter, fp->desc.sqp.trust_box_size);
      const unsigned char* p = reinterpret_cast<const unsigned char*>(fp->terms.data());
      std::printf("terms ");
      for (size_t i = 0; i < fp->terms.size() * sizeof(tb200_term); ++i) std::printf("%02x", p[i]);
      std::printf("\ninit");
      for (double v : fp->init_traj) std::printf(" %.17g", v);
      std::printf("\ntargets");
      for (double v : fp->cart_targets) std::printf(" %.17g", v);
      std::printf("\n");
      return 0;
    } catch (const std::runtime_error& e) {
      std::fprintf(stderr, "runtime_error: %s\n", e.what());
      return 3;
    }
  }

  auto vel = std::make_shared<JointVelTermInfo>();
  vel->term_type = TT_COST;
  vel->first_step = 0; vel->last_step = T - 1;
  auto acc = std::make_shared<JointAccTermInfo>();
  acc->term_type = TT_COST;
  acc->first_step = 0; acc->last_step = T - 1;
  pci.cost_infos = {vel, acc};
  auto pose = std::make_shared<CartPoseTermInfo>();
  pose->term_type = TT_CNT;
  pose->timestep = T - 1;
  pose->source_frame = tool;
  pose->target = goals;
  auto coll = std::make_shared<CollisionTermInfo>();
  coll->term_type = TT_CNT;
  coll->first_step = 0; coll->last_step = T - 1;
  coll->fixed_steps = {0};
  coll->evaluator_type = TB200_COLL_DISCRETE;
  coll->collision_margin = 0.02; coll->collision_coeff = 20.0; coll->collision_margin_buffer = 0.01;
  coll->longest_valid_segment_length = 0.5;
  pci.cnt_infos = {pose, coll};

  try {
    if (mode == "dump") {
      auto fp = FlattenProblem(pci);
      std::printf("n_terms %d n_cart_targets %d n_fixed %d\n", fp->desc.n_terms, fp->desc.n_cart_targets, fp->desc.n_fixed_timesteps);
      const unsigned char* p = reinterpret_cast<const unsigned char*>(fp->terms.data());
      std::printf("terms ");
      for (size_t i = 0; i < fp->terms.size() * sizeof(tb200_term); ++i) std::printf("%02x", p[i]);
      std::printf("\ninit");
      for (double v : fp->init_traj) std::printf(" %.17g", v);
      std::printf("\ntargets");
      for (double v : fp->cart_targets) std::printf(" %.17g", v);
      std::printf("\n");
      return 0;
    }
    TrajOptProb::Ptr prob = ConstructProblem(pci);
    // "solve": with the description's own parameters (pci.opt_info); "solve_ref": trajopt::OptimizeProblem, which
    // overrides four of them on top of the optimizer's defaults (problem_description.cpp:394-408)
    std::vector<tb::sco::OptResults> res = (mode == "solve_ref") ? OptimizeProblem(*prob) : OptimizeWithParams(*prob);
    for (const auto& r : res) {
      std::printf("result %d %.17g %d %d", static_cast<int>(r.status), r.total_cost, r.n_qp_solves, r.n_func_evals);
      for (double v : r.x) std::printf(" %.17g", v);
      std::printf("\n");
    }
    return 0;
  } catch (const std::runtime_error& e) {  // PRINT_AND_THROW in the reference
    std::fprintf(stderr, "runtime_error: %s\n", e.what());
    return 3;
  }
}
