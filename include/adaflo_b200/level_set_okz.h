// level_set_okz.h -- C++ mirror of the operator interface of adaflo's level-set solver classes on top of the C ABI
// (include/adaflo_b200.h, "level-set operators").  The classes keep the reference's names and method names so that
// the deal.II solver calls of the reference compile against them with device vectors:
//   LevelSetOKZSolverAdvanceConcentration::advance_concentration_vmult   include/adaflo/level_set_okz_advance_concentration.h:121
//   LevelSetOKZSolverComputeNormal::compute_normal_vmult                 include/adaflo/level_set_okz_compute_normal.h:96
//   LevelSetOKZSolverComputeCurvature::compute_curvature_vmult           include/adaflo/level_set_okz_compute_curvature.h:101
//   LevelSetOKZSolverReinitialization::reinitialization_vmult            include/adaflo/level_set_okz_reinitialization.h:115
// plus the right-hand-side kernels (local_*_rhs), LevelSetOKZSolver::compute_heaviside / compute_force
// (source/level_set_okz.cc:317-540) and the mass-matrix diagonal (include/adaflo/level_set_okz_preconditioner.h:31-76).
// In the reference the four classes share one MatrixFree, the cell diameters and the time-stepping state; here they
// share one LevelSetOperators handle.  Header only; link with libadaflo_b200.so.
#pragma once

#include <array>
#include <vector>

#include "adaflo_b200/navier_stokes_matrix.h"

namespace adaflo_b200
{
  /// the parameters the operator kernels read from LevelSetOKZSolver*Parameter / FlowParameters
  struct LevelSetParameters
  {
    int          dim                         = 3;
    unsigned int concentration_subdivisions  = 3;
    bool         convection_stabilization    = false;
    double       epsilon                     = 1.5;
    double       epsilon_used                = 0.;   ///< epsilon scaled with the mesh size (LevelSetOKZSolver::initialize_data_structures)
    double       damping_scale_factor        = 4.;
    double       minimal_edge_length         = 0.;
    adaflo_b200_ls_params to_abi() const
    {
      adaflo_b200_ls_params p = {};
      p.dim                        = dim;
      p.concentration_subdivisions = static_cast<int>(concentration_subdivisions);
      p.convection_stabilization   = convection_stabilization ? 1 : 0;
      p.epsilon                    = epsilon;
      p.epsilon_used               = epsilon_used;
      p.damping_scale_factor       = damping_scale_factor;
      p.minimal_edge_length        = minimal_edge_length;
      return p;
    }
  };

  using NormalVector = std::array<Vector, 3>;   ///< LinearAlgebra::distributed::BlockVector<double> with dim blocks

  /// state shared by the level-set operator classes: MatrixFree data, parameters, cell diameters, time weights
  class LevelSetOperators
  {
  public:
    LevelSetOperators(const MatrixFreeData      &matrix_free,
                      const LevelSetParameters  &parameters,
                      const std::vector<double> &cell_diameters,
                      const unsigned int         dof_index_ls,
                      const int                  dof_index_vel,
                      const unsigned int         quad_index)
      : matrix_free(matrix_free)
    {
      const adaflo_b200_ls_params p = parameters.to_abi();
      check(adaflo_b200_ls_create(matrix_free.context().handle(), matrix_free.handle(), &p, static_cast<int>(dof_index_ls), dof_index_vel,
                                  static_cast<int>(quad_index), cell_diameters.data(), &h));
    }
    ~LevelSetOperators() { adaflo_b200_ls_destroy(h); }
    LevelSetOperators(const LevelSetOperators &)            = delete;
    LevelSetOperators &operator=(const LevelSetOperators &) = delete;

    adaflo_b200_ls_t      handle() const { return h; }
    const MatrixFreeData &get_matrix_free() const { return matrix_free; }

    /// TimeStepping weights of the concentration equation and the step size (pseudo time step of the reinitialisation)
    void set_time(const TimeStepping &ts, const double step_size) { check(adaflo_b200_ls_set_time(h, ts.weight(), ts.weight_old(), ts.weight_old_old(), step_size)); }
    /// diagonal used on constrained rows (DiagonalPreconditioner::get_vector in the reference)
    void set_diagonal(const Vector &diagonal) { check(adaflo_b200_ls_set_diagonal(h, diagonal.begin())); }
    /// initialize_mass_matrix_diagonal
    void initialize_mass_matrix_diagonal(Vector &diagonal) const { check(adaflo_b200_ls_mass_diagonal(h, diagonal.begin())); }

    /// LevelSetOKZSolver::compute_heaviside
    void compute_heaviside(Vector &heaviside, const Vector &level_set) const { check(adaflo_b200_ls_compute_heaviside(h, heaviside.begin(), level_set.begin())); }
    /// LevelSetOKZSolver::compute_force (local_compute_force); writes rho / mu into the Navier-Stokes operator if variable
    template <int dim>
    void compute_force(NavierStokesMatrix<dim> &navier_stokes_matrix,
                       Vector                  &force,
                       const Vector            &heaviside,
                       const Vector            &curvature,
                       const double             surface_tension,
                       const double             gravity,
                       const double            *interpolation_concentration_pressure = nullptr) const
    {
      check(adaflo_b200_ls_compute_force(h, navier_stokes_matrix.handle(), force.begin(), heaviside.begin(), curvature.begin(), surface_tension, gravity,
                                         interpolation_concentration_pressure));
    }

  private:
    const MatrixFreeData &matrix_free;
    adaflo_b200_ls_t      h = nullptr;
  };

  template <int dim>
  class LevelSetOKZSolverAdvanceConcentration
  {
  public:
    explicit LevelSetOKZSolverAdvanceConcentration(const LevelSetOperators &op) : op(op) {}
    void advance_concentration_vmult(Vector &dst, const Vector &src) const { check(adaflo_b200_ls_advance_concentration_vmult(op.handle(), dst.begin(), src.begin())); }
    /// local_advance_concentration_rhs: BDF residual with the Q2 velocity evaluated at the level-set quadrature points
    void advance_concentration_rhs(Vector &rhs, const Vector &vel_solution, const Vector &solution, const Vector &solution_old, const Vector &solution_old_old) const
    {
      check(adaflo_b200_ls_advance_concentration_rhs(op.handle(), rhs.begin(), vel_solution.begin(), solution.begin(), solution_old.begin(),
                                                     solution_old_old.begin()));
    }

  private:
    const LevelSetOperators &op;
  };

  template <int dim>
  class LevelSetOKZSolverComputeNormal
  {
  public:
    explicit LevelSetOKZSolverComputeNormal(const LevelSetOperators &op) : op(op) {}
    void compute_normal_vmult(NormalVector &dst, const NormalVector &src) const
    {
      double       *d[3] = {dst[0].begin(), dst[1].begin(), dst[2].begin()};
      const double *s[3] = {src[0].begin(), src[1].begin(), src[2].begin()};
      check(adaflo_b200_ls_compute_normal_vmult(op.handle(), d, s));
    }
    void compute_normal_rhs(NormalVector &rhs, const Vector &level_set) const
    {
      double *d[3] = {rhs[0].begin(), rhs[1].begin(), rhs[2].begin()};
      check(adaflo_b200_ls_compute_normal_rhs(op.handle(), d, level_set.begin()));
    }

  private:
    const LevelSetOperators &op;
  };

  template <int dim>
  class LevelSetOKZSolverComputeCurvature
  {
  public:
    explicit LevelSetOKZSolverComputeCurvature(const LevelSetOperators &op) : op(op) {}
    void compute_curvature_vmult(Vector &dst, const Vector &src, const bool apply_diffusion) const
    {
      check(adaflo_b200_ls_compute_curvature_vmult(op.handle(), dst.begin(), src.begin(), apply_diffusion ? 1 : 0));
    }
    void compute_curvature_rhs(Vector &rhs, const NormalVector &normal) const
    {
      const double *n[3] = {normal[0].begin(), normal[1].begin(), normal[2].begin()};
      check(adaflo_b200_ls_compute_curvature_rhs(op.handle(), rhs.begin(), n));
    }
    /// the correction of the curvature by the distance to the interface (compute_curvature.cc:362-377)
    void curvature_correction(Vector &curvature, const Vector &level_set) const { check(adaflo_b200_ls_curvature_correction(op.handle(), curvature.begin(), level_set.begin())); }

  private:
    const LevelSetOperators &op;
  };

  template <int dim>
  class LevelSetOKZSolverReinitialization
  {
  public:
    explicit LevelSetOKZSolverReinitialization(const LevelSetOperators &op) : op(op) {}
    void reinitialization_vmult(Vector &dst, const Vector &src, const bool diffuse_only) const
    {
      check(adaflo_b200_ls_reinitialization_vmult(op.handle(), dst.begin(), src.begin(), diffuse_only ? 1 : 0));
    }
    /// local_reinitialize_rhs; normal == nullptr with diffuse_only
    void reinitialization_rhs(Vector &rhs, const Vector &solution, const NormalVector *normal, const bool diffuse_only, const bool first_reinit_step) const
    {
      const double *n[3] = {nullptr, nullptr, nullptr};
      if (normal)
        for (int d = 0; d < 3; ++d) n[d] = (*normal)[d].begin();
      check(adaflo_b200_ls_reinitialization_rhs(op.handle(), rhs.begin(), solution.begin(), normal ? n : nullptr, diffuse_only ? 1 : 0, first_reinit_step ? 1 : 0));
    }

  private:
    const LevelSetOperators &op;
  };
} // namespace adaflo_b200
