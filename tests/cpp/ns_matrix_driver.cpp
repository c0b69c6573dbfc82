// ns_matrix_driver.cpp -- exercises the C++ mirror class (include/adaflo_b200/navier_stokes_matrix.h)
// the way adaflo's NavierStokes class drives NavierStokesMatrix (source/navier_stokes.cc:452,785,599):
// initialize(matrix_free, time_stepping, pressure_average_fix); residual(rhs, solution, user_rhs);
// vmult(dst, src).  Input arrays come from a dump written by tests/test_cpp_mirror.py, outputs go
// back the same way and are compared with the CPU oracle there.
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <string>
#include <vector>

#include "adaflo_b200/navier_stokes_matrix.h"

using namespace adaflo_b200;

struct Blob
{
  std::vector<char> bytes;
  template <typename T>
  const T *as() const { return reinterpret_cast<const T *>(bytes.data()); }
  template <typename T>
  std::size_t count() const { return bytes.size() / sizeof(T); }
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

static void write_array(std::ofstream &out, const std::string &name, const std::vector<double> &v)
{
  const std::uint32_t len = name.size();
  const std::uint64_t nbytes = v.size() * sizeof(double);
  out.write(reinterpret_cast<const char *>(&len), 4);
  out.write(name.data(), len);
  out.write(reinterpret_cast<const char *>(&nbytes), 8);
  out.write(reinterpret_cast<const char *>(v.data()), nbytes);
}

int main(int argc, char **argv)
{
  if (argc < 3) {
    std::fprintf(stderr, "usage: %s <input dump> <output dump>\n", argv[0]);
    return 2;
  }
  try {
    auto in = read_dump(argv[1]);
    const std::uint32_t *meta = in.at("meta").as<std::uint32_t>();  // n_cells, n_owned_u, n_owned_p, n_constrained_u
    const std::uint32_t n_cells = meta[0], n_u = meta[1], n_p = meta[2], n_con = meta[3];
    const double *prm = in.at("params").as<double>();               // viscosity, density, damping, tau_grad_div, beta, weights(6)

    // an uninitialised operator must fail loudly (Assert(matrix_free != 0, ExcNotInitialized()))
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

    Context ctx(0);
    BlockVector solution, solution_old, solution_old_old, system_rhs, user_rhs, src, dst;
    const unsigned int dof_index_u = 0, dof_index_p = 1, quad_index_u = 0, quad_index_p = 1;
    NavierStokesMatrix<3> navier_stokes_matrix(parameters, dof_index_u, dof_index_p, quad_index_u, quad_index_p, solution_old, solution_old_old);
    bool threw = false;
    try {
      navier_stokes_matrix.vmult(dst, src);
    } catch (const ExcNotInitialized &) {
      threw = true;
    }
    if (!threw) throw std::runtime_error("uninitialised vmult did not throw");

    // MatrixFree data (Cartesian box): spaces {u, p}, quadratures {QGauss(3), QGauss(2)}
    adaflo_b200_space_desc spaces[2] = {};
    spaces[0].fe_kind = ADAFLO_B200_FE_Q;
    spaces[0].degree = 2;
    spaces[0].n_components = 3;
    spaces[0].n_owned_nodes = n_u;
    spaces[0].nodes_per_cell = 27;
    spaces[0].cell_nodes = in.at("cell_u").as<std::uint32_t>();
    spaces[0].n_constrained = n_con;
    spaces[0].constrained_dofs = in.at("constrained_u").as<std::uint32_t>();
    spaces[1].fe_kind = ADAFLO_B200_FE_Q;
    spaces[1].degree = 1;
    spaces[1].n_components = 1;
    spaces[1].n_owned_nodes = n_p;
    spaces[1].nodes_per_cell = 8;
    spaces[1].cell_nodes = in.at("cell_p").as<std::uint32_t>();
    adaflo_b200_quad_desc quads[2] = {};
    for (int i = 0; i < 2; ++i) {
      quads[i].n_q_1d = i == 0 ? 3 : 2;
      quads[i].geometry_type = ADAFLO_B200_GEOM_CARTESIAN;
      quads[i].n_table_rows = 1;
      quads[i].table = in.at("geom_table").as<double>();
      quads[i].cell_row = in.at("geom_row").as<std::uint32_t>();
    }
    adaflo_b200_mf_desc desc = {};
    desc.dim = 3;
    desc.n_cells = n_cells;
    desc.n_spaces = 2;
    desc.spaces = spaces;
    desc.n_quads = 2;
    desc.quads = quads;
    MatrixFreeData matrix_free(ctx, desc);

    for (BlockVector *v : {&solution, &solution_old, &solution_old_old, &system_rhs, &user_rhs, &src, &dst}) {
      matrix_free.initialize_dof_vector(v->block(0), dof_index_u);
      matrix_free.initialize_dof_vector(v->block(1), dof_index_p);
    }
    solution.block(0).copy_from_host(in.at("solution_u").as<double>());
    solution.block(1).copy_from_host(in.at("solution_p").as<double>());
    solution_old.block(0).copy_from_host(in.at("old_u").as<double>());
    solution_old_old.block(0).copy_from_host(in.at("old_old_u").as<double>());
    src.block(0).copy_from_host(in.at("src_u").as<double>());
    src.block(1).copy_from_host(in.at("src_p").as<double>());

    navier_stokes_matrix.initialize(matrix_free, time_stepping, true);
    // compute_residual (navier_stokes.cc:781-790): rhs = const_rhs (= 0); residual; mean projection
    system_rhs = 0.;
    navier_stokes_matrix.residual(system_rhs, solution, user_rhs);
    navier_stokes_matrix.apply_pressure_average_projection(system_rhs.block(1));
    navier_stokes_matrix.vmult(dst, src);
    const double dot = dst * src;
    // begin_linearized_velocities(macro_cell) in the reference's storage order (navier_stokes_matrix.h:162-178): batches of 4
    // lanes, [q][12 rows][lane] per batch
    navier_stokes_matrix.set_cell_batches(4);
    const std::uint32_t n_batches = (n_cells + 3) / 4;
    std::vector<double> lin_canonical;
    for (std::uint32_t b = 0; b < n_batches; ++b) {
      const double *batch = static_cast<const NavierStokesMatrix<3> &>(navier_stokes_matrix).begin_linearized_velocities(b);
      lin_canonical.insert(lin_canonical.end(), batch, batch + 27 * 12 * 4);
    }
    const auto stats = navier_stokes_matrix.get_matvec_statistics();

    std::vector<double> ru(3 * (std::size_t)n_u), rp(n_p), du(3 * (std::size_t)n_u), dp(n_p);
    system_rhs.block(0).copy_to_host(ru.data());
    system_rhs.block(1).copy_to_host(rp.data());
    dst.block(0).copy_to_host(du.data());
    dst.block(1).copy_to_host(dp.data());
    std::ofstream out(argv[2], std::ios::binary);
    {
      const std::string name = "lin_canonical";
      const std::uint32_t len = name.size();
      const std::uint64_t nbytes = lin_canonical.size() * sizeof(double);
      out.write(reinterpret_cast<const char *>(&len), 4);
      out.write(name.data(), len);
      out.write(reinterpret_cast<const char *>(&nbytes), 8);
      out.write(reinterpret_cast<const char *>(lin_canonical.data()), nbytes);
    }
    write_array(out, "residual_u", ru);
    write_array(out, "residual_p", rp);
    write_array(out, "vmult_u", du);
    write_array(out, "vmult_p", dp);
    write_array(out, "scalars", {dot, (double)stats.second});
    std::printf("ns_matrix_driver: ok (matvecs counted: %u)\n", stats.second);
    return 0;
  } catch (const std::exception &e) {
    std::fprintf(stderr, "ns_matrix_driver: %s\n", e.what());
    return 1;
  }
}
