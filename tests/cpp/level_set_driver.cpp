// level_set_driver.cpp -- exercises the C++ mirror of the level-set operator classes (include/adaflo_b200/level_set_okz.h)
// the way LevelSetOKZSolver drives them: mass diagonal, advance_concentration rhs / vmult, compute_normal rhs,
// compute_curvature vmult, reinitialization vmult.  Input arrays come from a dump written by
// tests/test_cpp_mirror.py, outputs go back the same way and are compared with the CPU oracle there.
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <string>
#include <vector>

#include "adaflo_b200/level_set_okz.h"

using namespace adaflo_b200;

struct Blob
{
  std::vector<char> bytes;
  template <typename T>
  const T *as() const { return reinterpret_cast<const T *>(bytes.data()); }
};

static std::map<std::string, Blob> read_dump(const char *path)
{
  std::map<std::string, Blob> m;
  std::ifstream in(path, std::ios::binary);
  if (!in) throw std::runtime_error(std::string("cannot open ") + path);
  while (true) {
    std::uint32_t len = 0;
    if (!in.read(reinterpret_cast<char *>(&len), 4)) break;
    std::string name(len, ' ');
    in.read(&name[0], len);
    std::uint64_t nbytes = 0;
    in.read(reinterpret_cast<char *>(&nbytes), 8);
    Blob b;
    b.bytes.resize(nbytes);
    in.read(b.bytes.data(), nbytes);
    m[name] = std::move(b);
  }
  return m;
}

static void write_vector(std::ofstream &out, const std::string &name, const Vector &v)
{
  std::vector<double> host(v.locally_owned_size());
  v.copy_to_host(host.data());
  const std::uint32_t len = name.size();
  const std::uint64_t nbytes = host.size() * sizeof(double);
  out.write(reinterpret_cast<const char *>(&len), 4);
  out.write(name.data(), len);
  out.write(reinterpret_cast<const char *>(&nbytes), 8);
  out.write(reinterpret_cast<const char *>(host.data()), nbytes);
}

int main(int argc, char **argv)
{
  if (argc < 3) {
    std::fprintf(stderr, "usage: %s <input dump> <output dump>\n", argv[0]);
    return 2;
  }
  try {
    const auto in = read_dump(argv[1]);
    const std::uint32_t *meta = in.at("meta").as<std::uint32_t>();
    const std::uint32_t n_cells = meta[0], n_u = meta[1], n_p = meta[2], n_ls = meta[3], n_sub = meta[4];
    const double *prm = in.at("params").as<double>();  // epsilon, epsilon_used, damping scale, min edge, weight, w_old, w_old_old, dt

    Context ctx(0);
    adaflo_b200_space_desc spaces[3] = {};
    spaces[0].fe_kind = ADAFLO_B200_FE_Q;
    spaces[0].degree = 2;
    spaces[0].n_components = 3;
    spaces[0].n_owned_nodes = n_u;
    spaces[0].nodes_per_cell = 27;
    spaces[0].cell_nodes = in.at("cell_u").as<std::uint32_t>();
    spaces[1].fe_kind = ADAFLO_B200_FE_Q;
    spaces[1].degree = 1;
    spaces[1].n_components = 1;
    spaces[1].n_owned_nodes = n_p;
    spaces[1].nodes_per_cell = 8;
    spaces[1].cell_nodes = in.at("cell_p").as<std::uint32_t>();
    spaces[2].fe_kind = ADAFLO_B200_FE_Q_ISO_Q1;
    spaces[2].degree = static_cast<int>(n_sub);
    spaces[2].n_components = 1;
    spaces[2].n_owned_nodes = n_ls;
    spaces[2].nodes_per_cell = (n_sub + 1) * (n_sub + 1) * (n_sub + 1);
    spaces[2].cell_nodes = in.at("cell_ls").as<std::uint32_t>();
    adaflo_b200_quad_desc quads[3] = {};
    const int nq[3] = {3, 2, static_cast<int>(2 * n_sub)};
    for (int i = 0; i < 3; ++i) {
      quads[i].n_q_1d = nq[i];
      quads[i].geometry_type = ADAFLO_B200_GEOM_CARTESIAN;
      quads[i].n_table_rows = 1;
      quads[i].table = in.at("geom_table").as<double>();
      quads[i].cell_row = in.at("geom_row").as<std::uint32_t>();
    }
    adaflo_b200_mf_desc desc = {};
    desc.dim = 3;
    desc.n_cells = n_cells;
    desc.n_spaces = 3;
    desc.spaces = spaces;
    desc.n_quads = 3;
    desc.quads = quads;
    MatrixFreeData matrix_free(ctx, desc);

    LevelSetParameters parameters;
    parameters.concentration_subdivisions = n_sub;
    parameters.epsilon = prm[0];
    parameters.epsilon_used = prm[1];
    parameters.damping_scale_factor = prm[2];
    parameters.minimal_edge_length = prm[3];
    const double *cd = in.at("cell_diameters").as<double>();
    LevelSetOperators operators(matrix_free, parameters, std::vector<double>(cd, cd + n_cells), 2, 0, 2);
    TimeStepping time_stepping;
    time_stepping.weight_val = prm[4];
    time_stepping.weight_old_val = prm[5];
    time_stepping.weight_old_old_val = prm[6];
    operators.set_time(time_stepping, prm[7]);

    LevelSetOKZSolverAdvanceConcentration<3> advance(operators);
    LevelSetOKZSolverComputeNormal<3> normal_operator(operators);
    LevelSetOKZSolverComputeCurvature<3> curvature_operator(operators);
    LevelSetOKZSolverReinitialization<3> reinitialization(operators);

    Vector diagonal, c, c_old, c_old_old, src, rhs, dst, vel;
    for (Vector *v : {&diagonal, &c, &c_old, &c_old_old, &src, &rhs, &dst}) matrix_free.initialize_dof_vector(*v, 2);
    matrix_free.initialize_dof_vector(vel, 0);
    c.copy_from_host(in.at("c").as<double>());
    c_old.copy_from_host(in.at("c_old").as<double>());
    c_old_old.copy_from_host(in.at("c_old_old").as<double>());
    src.copy_from_host(in.at("src").as<double>());
    vel.copy_from_host(in.at("vel").as<double>());

    std::ofstream out(argv[2], std::ios::binary);
    operators.initialize_mass_matrix_diagonal(diagonal);
    operators.set_diagonal(diagonal);
    write_vector(out, "mass_diagonal", diagonal);
    rhs = 0.;
    advance.advance_concentration_rhs(rhs, vel, c, c_old, c_old_old);
    write_vector(out, "advance_rhs", rhs);
    advance.advance_concentration_vmult(dst, src);
    write_vector(out, "advance_vmult", dst);
    NormalVector normal_rhs;
    for (Vector &v : normal_rhs) {
      matrix_free.initialize_dof_vector(v, 2);
      v = 0.;
    }
    normal_operator.compute_normal_rhs(normal_rhs, c);
    write_vector(out, "normal_rhs_0", normal_rhs[0]);
    write_vector(out, "normal_rhs_2", normal_rhs[2]);
    curvature_operator.compute_curvature_vmult(dst, src, true);
    write_vector(out, "curvature_vmult", dst);
    reinitialization.reinitialization_vmult(dst, src, true);
    write_vector(out, "reinit_vmult_diffuse", dst);

    // ---- the reference's constructor argument lists and solver-level routines (level_set_okz_*.h), the way
    // LevelSetOKZSolver / SharpInterfaceSolver build and drive the four classes
    {
      const CellDiameters cell_diameters(cd, cd + n_cells);
      const double epsilon_used = prm[1], minimal_edge_length = prm[3], dt = prm[7];
      DiagonalPreconditioner preconditioner(ctx);
      preconditioner.reinit(diagonal);
      AffineConstraints constraints, constraints_curvature, constraints_normals;
      ConditionalOStream pcout(std::cout, true);
      std::shared_ptr<BlockMatrixExtension> projection_matrix;
      std::shared_ptr<BlockILUExtension> ilu_projection_matrix;
      NormalVector normal_vector_field, normal_vector_rhs;
      for (NormalVector *nv : {&normal_vector_field, &normal_vector_rhs})
        for (Vector &v : *nv) {
          matrix_free.initialize_dof_vector(v, 2);
          v = 0.;
        }
      Vector solution, solution_update, system_rhs, solution_curvature;
      for (Vector *v : {&solution, &solution_update, &system_rhs, &solution_curvature}) matrix_free.initialize_dof_vector(*v, 2);
      solution.copy_from_host(in.at("c").as<double>());

      LevelSetOKZSolverComputeNormalParameter normal_parameters;
      normal_parameters.dof_index_ls = normal_parameters.dof_index_normal = 2;
      normal_parameters.quad_index = 2;
      normal_parameters.epsilon = prm[0];
      normal_parameters.damping_scale_factor = prm[2];
      LevelSetOKZSolverComputeNormal<3> normal_solver(normal_vector_field, normal_vector_rhs, solution, cell_diameters, epsilon_used, minimal_edge_length,
                                                      constraints_normals, normal_parameters, matrix_free, preconditioner, projection_matrix,
                                                      ilu_projection_matrix);
      normal_solver.compute_normal(false);
      write_vector(out, "solved_normal_0", normal_vector_field[0]);
      write_vector(out, "solved_normal_1", normal_vector_field[1]);
      write_vector(out, "solved_normal_2", normal_vector_field[2]);

      LevelSetOKZSolverComputeCurvatureParameter curvature_parameters;
      curvature_parameters.dof_index_ls = curvature_parameters.dof_index_curvature = curvature_parameters.dof_index_normal = 2;
      curvature_parameters.quad_index = 2;
      curvature_parameters.epsilon = prm[0];
      curvature_parameters.curvature_correction = false;  // the correction is singular where kappa d = -(dim - 1): tested on equal inputs elsewhere
      LevelSetOKZSolverComputeCurvature<3> curvature_solver(cell_diameters, normal_vector_field, constraints_curvature, constraints, epsilon_used, system_rhs,
                                                            curvature_parameters, solution_curvature, solution, matrix_free, preconditioner,
                                                            projection_matrix, ilu_projection_matrix);
      curvature_solver.compute_curvature();
      write_vector(out, "solved_curvature", solution_curvature);

      LevelSetOKZSolverAdvanceConcentrationParameter advance_parameters;
      advance_parameters.dof_index_ls = 2;
      advance_parameters.dof_index_vel = 0;
      advance_parameters.quad_index = 2;
      advance_parameters.time.time_step_scheme = TimeSteppingParameters::Scheme::bdf_2;
      advance_parameters.time.time_step_size_start = dt;
      advance_parameters.time.end_time = 10.;
      LevelSetOKZSolverAdvanceConcentrationBoundaryDescriptor<3> boundary;
      Vector advected;
      matrix_free.initialize_dof_vector(advected, 2);
      advected.copy_from_host(in.at("c").as<double>());
      LevelSetOKZSolverAdvanceConcentration<3> advance_solver(advected, c_old, c_old_old, solution_update, system_rhs, vel, vel, vel, cell_diameters, constraints,
                                                              pcout, boundary, matrix_free, advance_parameters, preconditioner);
      advance_solver.advance_concentration(dt);  // first step: weights 1/dt, -1/dt, 0
      std::cout << std::endl;
      write_vector(out, "advected", advected);

      LevelSetOKZSolverReinitializationParameter reinit_parameters;
      reinit_parameters.dof_index_ls = reinit_parameters.dof_index_normal = 2;
      reinit_parameters.quad_index = 2;
      reinit_parameters.time.time_step_size_start = dt;
      reinit_parameters.time.end_time = 10.;
      std::pair<double, double> last_concentration_range(-1., 1.);
      bool first_reinit_step = true;
      LevelSetOKZSolverReinitialization<3> reinit_solver(normal_vector_field, cell_diameters, epsilon_used, minimal_edge_length, constraints, solution_update,
                                                         solution, system_rhs, pcout, preconditioner, last_concentration_range, reinit_parameters,
                                                         first_reinit_step, matrix_free);
      reinit_solver.reinitialize(dt, 2, 1);
      write_vector(out, "reinitialized", solution);
      Vector counts;  // iteration counts as doubles: normal, curvature, advection, then the reinitialisation steps
      std::vector<double> cnt = {double(normal_solver.last_iterations()), double(curvature_solver.last_iterations()),
                                 double(advance_solver.last_iterations())};
      for (const int i : reinit_solver.last_iterations()) cnt.push_back(double(i));
      counts.reinit(ctx, cnt.size(), cnt.size());
      counts.copy_from_host(cnt.data());
      write_vector(out, "counts", counts);
    }
    std::printf("level_set_driver: ok\n");
    return 0;
  } catch (const std::exception &e) {
    std::fprintf(stderr, "level_set_driver: %s\n", e.what());
    return 1;
  }
}
