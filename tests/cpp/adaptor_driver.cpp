// adaptor_driver.cpp -- INTEGRATION.md section 2 compiled and run: a mock of dealii::MatrixFree's batch / lane storage
// (cells grouped in batches of n_lanes, plain_dof_indices component-blocked per cell, a velocity numbering scrambled
// like DoFRenumbering::Cuthill_McKee leaves it: the three components of a node are NOT consecutive) goes through
// adaflo_b200::build_matrix_free; the operator is then applied to HOST vectors in the mock's numbering
// (NavierStokesMatrix::vmult_host, velocity_vmult_host) and tests/test_cpp_mirror.py compares with the CPU oracle.
#include <cstdio>
#include <fstream>
#include <map>
#include <string>
#include <vector>

#include "adaflo_b200/matrix_free_adaptor.h"

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
static void write_array(std::ofstream &out, const std::string &name, const std::vector<double> &host)
{
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
    const std::uint32_t *meta = in.at("meta").as<std::uint32_t>();  // n_cells, n_dofs_u, n_dofs_p, n_constrained_u, n_lanes
    const std::uint32_t n_cells = meta[0], n_dofs_u = meta[1], n_dofs_p = meta[2], n_con = meta[3], n_lanes = meta[4];
    const double *prm = in.at("params").as<double>();
    // the mock MatrixFree: plain_dof_indices [cell][3 * 27] / [cell][8] in the HOST numbering, cells in batches of n_lanes
    const unsigned int *dofs_u = in.at("plain_dof_indices_u").as<unsigned int>();
    const unsigned int *dofs_p = in.at("plain_dof_indices_p").as<unsigned int>();
    const double *inv_h = in.at("inverse_jacobian_diagonal").as<double>();  // [cell][3]
    MatrixFreeView view;
    view.n_cell_batches = (n_cells + n_lanes - 1) / n_lanes;
    view.n_active_entries_per_cell_batch = [=](unsigned int b) { return std::min(n_lanes, n_cells - b * n_lanes); };
    DoFHandlerView u, p;
    u.fe_kind = ADAFLO_B200_FE_Q;
    u.degree = 2;
    u.n_components = 3;
    u.nodes_per_cell = 27;
    u.n_owned_dofs = n_dofs_u;
    u.plain_dof_indices = [=](unsigned int b, unsigned int l) { return dofs_u + static_cast<std::size_t>(b * n_lanes + l) * 81; };
    u.constrained_dofs.assign(in.at("constrained_u").as<unsigned int>(), in.at("constrained_u").as<unsigned int>() + n_con);
    p.fe_kind = ADAFLO_B200_FE_Q;
    p.degree = 1;
    p.n_components = 1;
    p.nodes_per_cell = 8;
    p.n_owned_dofs = n_dofs_p;
    p.plain_dof_indices = [=](unsigned int b, unsigned int l) { return dofs_p + static_cast<std::size_t>(b * n_lanes + l) * 8; };
    view.dof_handlers = {u, p};
    for (const int n_q_1d : {3, 2}) {
      QuadratureView q;
      q.n_q_1d = n_q_1d;
      q.cell_type = [](unsigned int) { return int(ADAFLO_B200_GEOM_CARTESIAN); };
      q.geometry = [=](unsigned int b, unsigned int l, unsigned int, double (&k)[9], double &det) {
        const double *d = inv_h + static_cast<std::size_t>(b * n_lanes + l) * 3;
        for (double &e : k) e = 0;
        k[0] = d[0];
        k[4] = d[1];
        k[8] = d[2];
        det = 1. / (d[0] * d[1] * d[2]);
      };
      view.quadratures.push_back(q);
    }

    Context ctx(0);
    AdaptedMatrixFree adapted = build_matrix_free(ctx, view);
    const MatrixFreeData &matrix_free = *adapted.matrix_free;

    FlowParameters parameters;
    parameters.viscosity = prm[0];
    parameters.density = prm[1];
    parameters.damping = prm[2];
    parameters.tau_grad_div = prm[3];
    parameters.beta_convective_term_momentum_balance = prm[4];
    TimeStepping time_stepping;
    time_stepping.weight_val = prm[5];
    time_stepping.weight_old_val = prm[6];
    time_stepping.weight_old_old_val = prm[7];
    time_stepping.tau1_val = prm[8];
    time_stepping.factor_extrapol_old = prm[9];
    time_stepping.factor_extrapol_old_old = prm[10];
    BlockVector solution, solution_old, solution_old_old, system_rhs, user_rhs;
    NavierStokesMatrix<3> navier_stokes_matrix(parameters, 0, 1, 0, 1, solution_old, solution_old_old);
    for (BlockVector *v : {&solution, &solution_old, &solution_old_old, &system_rhs, &user_rhs}) {
      matrix_free.initialize_dof_vector(v->block(0), 0);
      matrix_free.initialize_dof_vector(v->block(1), 1);
    }
    // host vectors arrive in the mock's (scrambled) numbering: the numbering-aware upload
    check(adaflo_b200_mf_upload_vector(matrix_free.handle(), 0, solution.block(0).begin(), in.at("solution_u").as<double>()));
    check(adaflo_b200_mf_upload_vector(matrix_free.handle(), 1, solution.block(1).begin(), in.at("solution_p").as<double>()));
    check(adaflo_b200_mf_upload_vector(matrix_free.handle(), 0, solution_old.block(0).begin(), in.at("old_u").as<double>()));
    check(adaflo_b200_mf_upload_vector(matrix_free.handle(), 0, solution_old_old.block(0).begin(), in.at("old_old_u").as<double>()));
    navier_stokes_matrix.initialize(matrix_free, time_stepping, true);
    user_rhs = 0.;
    navier_stokes_matrix.residual(system_rhs, solution, user_rhs);  // stores the linearisation point

    std::vector<double> dst_u(n_dofs_u), dst_p(n_dofs_p), vel(n_dofs_u), res_u(n_dofs_u);
    check(adaflo_b200_ns_vmult_host(navier_stokes_matrix.handle(), dst_u.data(), dst_p.data(), in.at("src_u").as<double>(), in.at("src_p").as<double>()));
    navier_stokes_matrix.fix_linearization_point();  // the preconditioner's copy of the linearisation point
    check(adaflo_b200_ns_velocity_vmult_host(navier_stokes_matrix.handle(), vel.data(), in.at("src_u").as<double>()));
    check(adaflo_b200_mf_download_vector(matrix_free.handle(), 0, res_u.data(), system_rhs.block(0).begin()));
    std::ofstream out(argv[2], std::ios::binary);
    write_array(out, "vmult_u", dst_u);
    write_array(out, "vmult_p", dst_p);
    write_array(out, "velocity_vmult", vel);
    write_array(out, "residual_u", res_u);
    std::printf("adaptor_driver: ok (%u cells in %u batches of %u lanes)\n", n_cells, view.n_cell_batches, n_lanes);
    return 0;
  } catch (const std::exception &e) {
    std::fprintf(stderr, "adaptor_driver: %s\n", e.what());
    return 1;
  }
}
