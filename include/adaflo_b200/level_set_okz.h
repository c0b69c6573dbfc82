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
#include <functional>
#include <map>
#include <memory>
#include <ostream>
#include <set>
#include <utility>
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
    void set_time(const TimeStepping &ts, const double step_size) const { check(adaflo_b200_ls_set_time(h, ts.weight(), ts.weight_old(), ts.weight_old_old(), step_size)); }
    /// diagonal used on constrained rows (DiagonalPreconditioner::get_vector in the reference)
    void set_diagonal(const Vector &diagonal) const { check(adaflo_b200_ls_set_diagonal(h, diagonal.begin())); }
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

  // ------------------------------------------------------------------------------------------------------------
  // Stand-ins for the deal.II / adaflo types in the constructor lists of the reference's level-set classes.  The
  // constraints live in the MatrixFree upload (flag bits / hanging-node masks), the output stream is optional, the
  // assembled projection matrix and its ILU are replaced by matrix-free CG with the diagonal preconditioner.
  struct AffineConstraints
  {
    void distribute(Vector &) const {}  ///< hanging nodes are not stored, constrained rows are kept by the operators
  };
  struct ConditionalOStream
  {
    std::ostream *stream = nullptr;
    bool          active = false;
    ConditionalOStream() = default;
    ConditionalOStream(std::ostream &s, const bool a) : stream(&s), active(a) {}
    template <typename T>
    const ConditionalOStream &operator<<(const T &t) const
    {
      if (stream && active)
        (*stream) << t;
      return *this;
    }
  };
  struct BlockMatrixExtension
  {};
  struct BlockILUExtension
  {};
  /// AlignedVector<VectorizedArray<double>> cell_diameters of the reference: here one value per cell in cell order
  using CellDiameters = std::vector<double>;

  struct LevelSetOKZSolverAdvanceConcentrationParameter  // level_set_okz_advance_concentration.h:30-66
  {
    unsigned int           dof_index_ls             = 2;
    unsigned int           dof_index_vel            = 0;
    unsigned int           quad_index               = 2;
    bool                   convection_stabilization = false;
    bool                   do_iteration             = false;
    double                 tol_nl_iteration         = 1e-8;
    TimeSteppingParameters time;
  };
  template <int dim>
  struct LevelSetOKZSolverAdvanceConcentrationBoundaryDescriptor  // :72-84; Dirichlet values are applied by the host (deal.II) side
  {
    std::map<unsigned int, std::function<double(const std::array<double, dim> &)>> dirichlet;
    std::set<unsigned int>                                                         symmetry;
  };
  struct LevelSetOKZSolverComputeNormalParameter  // level_set_okz_compute_normal.h:34-67
  {
    unsigned int dof_index_ls            = 2;
    unsigned int dof_index_normal        = 2;
    unsigned int quad_index              = 2;
    double       epsilon                 = 1.5;
    bool         approximate_projections = false;
    double       damping_scale_factor    = 4.0;
  };
  struct LevelSetOKZSolverComputeCurvatureParameter  // level_set_okz_compute_curvature.h:32-75
  {
    unsigned int dof_index_ls            = 2;
    unsigned int dof_index_curvature     = 2;
    unsigned int dof_index_normal        = 2;
    unsigned int quad_index              = 2;
    double       epsilon                 = 1.5;
    bool         approximate_projections = false;
    bool         curvature_correction    = true;
  };
  struct LevelSetOKZSolverReinitializationParameter  // level_set_okz_reinitialization.h:32-62
  {
    unsigned int           dof_index_ls     = 2;
    unsigned int           dof_index_normal = 2;
    unsigned int           quad_index       = 2;
    bool                   do_iteration     = false;
    TimeSteppingParameters time;
  };

  namespace internal
  {
    /// handle owned by one of the classes below when it is built from the reference's constructor arguments
    inline std::unique_ptr<LevelSetOperators> make_operators(const MatrixFreeData &matrix_free,
                                                             const CellDiameters  &cell_diameters,
                                                             const int             dim,
                                                             const unsigned int    dof_index_ls,
                                                             const int             dof_index_vel,
                                                             const unsigned int    quad_index,
                                                             const double          epsilon,
                                                             const double          epsilon_used,
                                                             const double          minimal_edge_length,
                                                             const double          damping_scale_factor,
                                                             const bool            convection_stabilization)
    {
      LevelSetParameters p;
      int                degree = 0;
      check(adaflo_b200_mf_space_info(matrix_free.handle(), static_cast<int>(dof_index_ls), nullptr, &degree, nullptr));
      p.dim                        = dim;
      p.concentration_subdivisions = static_cast<unsigned int>(degree);  // FE_Q_iso_Q1(n): the degree is the subdivision count
      p.convection_stabilization   = convection_stabilization;
      p.epsilon                    = epsilon;
      p.epsilon_used               = epsilon_used;
      p.damping_scale_factor       = damping_scale_factor;
      p.minimal_edge_length        = minimal_edge_length;
      return std::unique_ptr<LevelSetOperators>(new LevelSetOperators(matrix_free, p, cell_diameters, dof_index_ls, dof_index_vel, quad_index));
    }
    /// hands the diagonal of the reference's DiagonalPreconditioner to the library whenever it has been re-initialised
    struct DiagonalLink
    {
      const DiagonalPreconditioner *preconditioner = nullptr;
      mutable unsigned int          seen           = 0;
      void sync(const LevelSetOperators &op) const
      {
        if (preconditioner && preconditioner->generation() != seen)
          {
            check(adaflo_b200_ls_set_diagonal(op.handle(), preconditioner->get_vector().begin()));
            seen = preconditioner->generation();
          }
      }
    };
  } // namespace internal

  /// include/adaflo/level_set_okz_advance_concentration.h:86-226
  template <int dim>
  class LevelSetOKZSolverAdvanceConcentration
  {
  public:
    using VectorType = Vector;
    /// the reference's constructor (:93-110); vel_solution_old / _old_old are unused by the reference's kernels as well
    LevelSetOKZSolverAdvanceConcentration(VectorType                                                         &solution,
                                          const VectorType                                                   &solution_old,
                                          const VectorType                                                   &solution_old_old,
                                          VectorType                                                         &increment,
                                          VectorType                                                         &rhs,
                                          const VectorType                                                   &vel_solution,
                                          const VectorType                                                   &vel_solution_old,
                                          const VectorType                                                   &vel_solution_old_old,
                                          const CellDiameters                                                &cell_diameters,
                                          const AffineConstraints                                            &constraints,
                                          const ConditionalOStream                                           &pcout,
                                          const LevelSetOKZSolverAdvanceConcentrationBoundaryDescriptor<dim> &boundary,
                                          const MatrixFreeData                                               &matrix_free,
                                          const LevelSetOKZSolverAdvanceConcentrationParameter               &parameters,
                                          const DiagonalPreconditioner                                       &preconditioner)
      : owned(internal::make_operators(matrix_free, cell_diameters, dim, parameters.dof_index_ls, static_cast<int>(parameters.dof_index_vel),
                                       parameters.quad_index, 1., 1., 0., 4., parameters.convection_stabilization))
      , op(*owned)
      , parameters(parameters)
      , solution(&solution)
      , solution_old(&solution_old)
      , solution_old_old(&solution_old_old)
      , increment(&increment)
      , rhs(&rhs)
      , vel_solution(&vel_solution)
      , constraints(&constraints)
      , pcout(pcout)
      , time_stepping(parameters.time)
    {
      (void)vel_solution_old;
      (void)vel_solution_old_old;
      (void)boundary;
      diagonal.preconditioner = &preconditioner;
    }
    /// operator-only view on a shared handle
    explicit LevelSetOKZSolverAdvanceConcentration(const LevelSetOperators &op) : op(op) {}
    virtual ~LevelSetOKZSolverAdvanceConcentration() = default;

    /// :503-693: advance the time weights, right-hand side, BiCGStab (GMRES fallback), solution += increment
    virtual void advance_concentration(const double dt)
    {
      time_stepping.set_desired_time_step(dt);
      time_stepping.next();
      op.set_time(time_stepping, time_stepping.step_size());
      diagonal.sync(op);
      check(adaflo_b200_ls_advance_concentration(op.handle(), solution->begin(), solution_old->begin(), solution_old_old->begin(), vel_solution->begin(),
                                                 parameters.tol_nl_iteration, &initial_residual, &n_iterations));
      if (!parameters.do_iteration)
        pcout << "  Concentration advance: advect [" << initial_residual << "/" << n_iterations << "]";
    }
    void advance_concentration_vmult(VectorType &dst, const VectorType &src) const
    {
      diagonal.sync(op);
      check(adaflo_b200_ls_advance_concentration_vmult(op.handle(), dst.begin(), src.begin()));
    }
    /// local_advance_concentration_rhs: BDF residual with the Q2 velocity evaluated at the level-set quadrature points
    void advance_concentration_rhs(Vector &rhs, const Vector &vel_solution, const Vector &solution, const Vector &solution_old, const Vector &solution_old_old) const
    {
      check(adaflo_b200_ls_advance_concentration_rhs(op.handle(), rhs.begin(), vel_solution.begin(), solution.begin(), solution_old.begin(),
                                                     solution_old_old.begin()));
    }
    const TimeStepping &get_time_stepping() const { return time_stepping; }
    double              last_initial_residual() const { return initial_residual; }
    int                 last_iterations() const { return n_iterations; }

  private:
    std::unique_ptr<LevelSetOperators>                   owned;
    const LevelSetOperators                             &op;
    const LevelSetOKZSolverAdvanceConcentrationParameter parameters;
    VectorType                                          *solution = nullptr;
    const VectorType                                    *solution_old = nullptr, *solution_old_old = nullptr;
    VectorType                                          *increment = nullptr, *rhs = nullptr;
    const VectorType                                    *vel_solution = nullptr;
    const AffineConstraints                             *constraints  = nullptr;
    ConditionalOStream                                   pcout;
    TimeStepping                                         time_stepping;
    internal::DiagonalLink                               diagonal;
    double                                               initial_residual = 0.;
    int                                                  n_iterations     = 0;
  };

  /// include/adaflo/level_set_okz_compute_normal.h:69-170
  template <int dim>
  class LevelSetOKZSolverComputeNormal
  {
  public:
    using VectorType      = Vector;
    using BlockVectorType = NormalVector;
    /// the reference's constructor (:76-88)
    LevelSetOKZSolverComputeNormal(BlockVectorType                               &normal_vector_field,
                                   BlockVectorType                               &normal_vector_rhs,
                                   const VectorType                              &level_set_field,
                                   const CellDiameters                           &cell_diameters,
                                   const double                                  &epsilon_used,
                                   const double                                  &minimal_edge_length,
                                   const AffineConstraints                       &constraints_normals,
                                   const LevelSetOKZSolverComputeNormalParameter &parameters,
                                   const MatrixFreeData                          &matrix_free,
                                   const DiagonalPreconditioner                  &preconditioner,
                                   const std::shared_ptr<BlockMatrixExtension>   &projection_matrix,
                                   const std::shared_ptr<BlockILUExtension>      &ilu_projection_matrix)
      : owned(internal::make_operators(matrix_free, cell_diameters, dim, parameters.dof_index_ls, -1, parameters.quad_index, parameters.epsilon,
                                       epsilon_used, minimal_edge_length, parameters.damping_scale_factor, false))
      , op(*owned)
      , parameters(parameters)
      , normal_vector_field(&normal_vector_field)
      , normal_vector_rhs(&normal_vector_rhs)
      , level_set_solution(&level_set_field)
    {
      if (parameters.approximate_projections)
        throw ExcNotImplemented("approximate_projections (the reference throws as well, compute_normal.cc:224-226)");
      (void)constraints_normals;
      (void)projection_matrix;
      (void)ilu_projection_matrix;
      diagonal.preconditioner = &preconditioner;
    }
    explicit LevelSetOKZSolverComputeNormal(const LevelSetOperators &op) : op(op) {}
    virtual ~LevelSetOKZSolverComputeNormal() = default;

    /// :211-289: right-hand side, CG(4000, 1e-50, fast ? 1e-5 : 1e-7) with the diagonal preconditioner, ghost update
    virtual void compute_normal(const bool fast_computation)
    {
      diagonal.sync(op);
      double *n[3] = {(*normal_vector_field)[0].begin(), (*normal_vector_field)[1].begin(), (*normal_vector_field)[2].begin()};
      check(adaflo_b200_ls_compute_normal(op.handle(), n, level_set_solution->begin(), fast_computation ? 1 : 0, &n_iterations));
    }
    void compute_normal_vmult(NormalVector &dst, const NormalVector &src) const
    {
      diagonal.sync(op);
      double       *d[3] = {dst[0].begin(), dst[1].begin(), dst[2].begin()};
      const double *s[3] = {src[0].begin(), src[1].begin(), src[2].begin()};
      check(adaflo_b200_ls_compute_normal_vmult(op.handle(), d, s));
    }
    void compute_normal_rhs(NormalVector &rhs, const Vector &level_set) const
    {
      double *d[3] = {rhs[0].begin(), rhs[1].begin(), rhs[2].begin()};
      check(adaflo_b200_ls_compute_normal_rhs(op.handle(), d, level_set.begin()));
    }
    int last_iterations() const { return n_iterations; }

  private:
    std::unique_ptr<LevelSetOperators>            owned;
    const LevelSetOperators                      &op;
    const LevelSetOKZSolverComputeNormalParameter parameters;
    BlockVectorType                              *normal_vector_field = nullptr, *normal_vector_rhs = nullptr;
    const VectorType                             *level_set_solution  = nullptr;
    internal::DiagonalLink                        diagonal;
    int                                           n_iterations = 0;
  };

  /// include/adaflo/level_set_okz_compute_curvature.h:77-190
  template <int dim>
  class LevelSetOKZSolverComputeCurvature
  {
  public:
    /// the reference's constructor (:80-93)
    LevelSetOKZSolverComputeCurvature(const CellDiameters                              &cell_diameters,
                                      const NormalVector                               &normal_vector_field,
                                      const AffineConstraints                          &constraints_curvature,
                                      const AffineConstraints                          &constraints,
                                      const double                                     &epsilon_used,
                                      Vector                                           &system_rhs,
                                      const LevelSetOKZSolverComputeCurvatureParameter &parameters,
                                      Vector                                           &solution_curvature,
                                      const Vector                                     &solution_ls,
                                      const MatrixFreeData                             &matrix_free,
                                      const DiagonalPreconditioner                     &preconditioner,
                                      std::shared_ptr<BlockMatrixExtension>            &projection_matrix,
                                      std::shared_ptr<BlockILUExtension>               &ilu_projection_matrix)
      : owned(internal::make_operators(matrix_free, cell_diameters, dim, parameters.dof_index_ls, -1, parameters.quad_index, parameters.epsilon,
                                       epsilon_used, 0., 4., false))
      , op(*owned)
      , parameters(parameters)
      , normal_vector_field(&normal_vector_field)
      , rhs(&system_rhs)
      , solution_curvature(&solution_curvature)
      , solution_ls(&solution_ls)
    {
      (void)constraints_curvature;
      (void)constraints;
      (void)projection_matrix;
      (void)ilu_projection_matrix;
      diagonal.preconditioner = &preconditioner;
    }
    explicit LevelSetOKZSolverComputeCurvature(const LevelSetOperators &op) : op(op) {}
    virtual ~LevelSetOKZSolverComputeCurvature() = default;

    /// :305-392: right-hand side from the normal field, CG(2000, 1e-50, 1e-8) on the projection operator, distance correction
    virtual void compute_curvature(const bool /*diffuse_large_values*/ = false)
    {
      diagonal.sync(op);
      if (parameters.approximate_projections)
        throw ExcNotImplemented("approximate_projections");
      const double *n[3] = {(*normal_vector_field)[0].begin(), (*normal_vector_field)[1].begin(), (*normal_vector_field)[2].begin()};
      check(adaflo_b200_ls_compute_curvature(op.handle(), solution_curvature->begin(), solution_ls->begin(), n, parameters.curvature_correction ? 1 : 0,
                                             use_projection_matrix_operator ? 1 : 0, &n_iterations));
    }
    void compute_curvature_vmult(Vector &dst, const Vector &src, const bool apply_diffusion) const
    {
      diagonal.sync(op);
      check(adaflo_b200_ls_compute_curvature_vmult(op.handle(), dst.begin(), src.begin(), apply_diffusion ? 1 : 0));
    }
    void compute_curvature_rhs(Vector &rhs, const NormalVector &normal) const
    {
      const double *n[3] = {normal[0].begin(), normal[1].begin(), normal[2].begin()};
      check(adaflo_b200_ls_compute_curvature_rhs(op.handle(), rhs.begin(), n));
    }
    /// the correction of the curvature by the distance to the interface (compute_curvature.cc:362-377)
    void curvature_correction(Vector &curvature, const Vector &level_set) const { check(adaflo_b200_ls_curvature_correction(op.handle(), curvature.begin(), level_set.begin())); }
    /// true (default): solve with the operator of the reference's assembled projection_matrix (mass + 4 delta^2 Laplace),
    /// false: with compute_curvature_vmult(apply_diffusion = true), the matrix-free operator of ComputeCurvatureMatrix
    bool use_projection_matrix_operator = true;
    int  last_iterations() const { return n_iterations; }

  private:
    std::unique_ptr<LevelSetOperators>               owned;
    const LevelSetOperators                         &op;
    const LevelSetOKZSolverComputeCurvatureParameter parameters;
    const NormalVector                              *normal_vector_field = nullptr;
    Vector                                          *rhs = nullptr, *solution_curvature = nullptr;
    const Vector                                    *solution_ls = nullptr;
    internal::DiagonalLink                           diagonal;
    int                                              n_iterations = 0;
  };

  /// include/adaflo/level_set_okz_reinitialization.h:64-200
  template <int dim>
  class LevelSetOKZSolverReinitialization
  {
  public:
    using VectorType      = Vector;
    using BlockVectorType = NormalVector;
    /// the reference's constructor (:74-88); the normal field is written by reinitialize (the reference passes a callback
    /// that fills it), so it is taken by non-const reference here
    LevelSetOKZSolverReinitialization(BlockVectorType                                  &normal_vector_field,
                                      const CellDiameters                              &cell_diameters,
                                      const double                                     &epsilon_used,
                                      const double                                     &minimal_edge_length,
                                      const AffineConstraints                          &constraints,
                                      VectorType                                       &solution_update,
                                      VectorType                                       &solution,
                                      VectorType                                       &system_rhs,
                                      const ConditionalOStream                         &pcout,
                                      const DiagonalPreconditioner                     &preconditioner,
                                      const std::pair<double, double>                  &last_concentration_range,
                                      const LevelSetOKZSolverReinitializationParameter &parameters,
                                      bool                                             &first_reinit_step,
                                      const MatrixFreeData                             &matrix_free)
      : owned(internal::make_operators(matrix_free, cell_diameters, dim, parameters.dof_index_ls, -1, parameters.quad_index, 1., epsilon_used,
                                       minimal_edge_length, 4., false))
      , op(*owned)
      , parameters(parameters)
      , normal_vector_field(&normal_vector_field)
      , solution(&solution)
      , last_concentration_range(&last_concentration_range)
      , first_reinit_step(&first_reinit_step)
      , pcout(pcout)
      , time_stepping(parameters.time)
    {
      (void)constraints;
      (void)solution_update;
      (void)system_rhs;
      diagonal.preconditioner = &preconditioner;
    }
    explicit LevelSetOKZSolverReinitialization(const LevelSetOperators &op) : op(op) {}

    /// :256-367.  compute_normal: the reference's callback; when empty the library recomputes the normal itself
    /// (fast tolerance) at the first compressive step, which is what the reference's callers pass
    void reinitialize(const double dt, const unsigned int stab_steps, const unsigned int diff_steps = 0,
                      const std::function<void(bool)> &compute_normal = std::function<void(bool)>())
    {
      diagonal.sync(op);
      time_stepping.set_desired_time_step(dt);
      op.set_time(time_stepping, dt);
      (void)compute_normal;
      const double range[2] = {last_concentration_range->first, last_concentration_range->second};
      std::vector<int> its(stab_steps + diff_steps + 3, 0);
      int              done = 0;
      double          *n[3] = {(*normal_vector_field)[0].begin(), (*normal_vector_field)[1].begin(), (*normal_vector_field)[2].begin()};
      check(adaflo_b200_ls_reinitialize(op.handle(), solution->begin(), n, dt, stab_steps, diff_steps, range, its.data(), &done));
      unsigned int actual_diff_steps = diff_steps + ((range[0] < -1.02 || range[1] > 1.02) ? 3 : 0);
      iterations.assign(its.begin(), its.begin() + done);
      *first_reinit_step = false;
      if (!parameters.do_iteration)
        {
          pcout << (time_stepping.now() == time_stepping.start() ? "  " : " and ") << "reinitialize (";
          for (int i = 0; i < done; ++i)
            {
              if (static_cast<unsigned int>(i) < actual_diff_steps)
                pcout << "d";
              pcout << iterations[i];
              if (i + 1 < done)
                pcout << " + ";
            }
          pcout << ")\n";
        }
      time_stepping.next();
    }
    void reinitialization_vmult(Vector &dst, const Vector &src, const bool diffuse_only) const
    {
      diagonal.sync(op);
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
    const std::vector<int> &last_iterations() const { return iterations; }

  private:
    std::unique_ptr<LevelSetOperators>               owned;
    const LevelSetOperators                         &op;
    const LevelSetOKZSolverReinitializationParameter parameters;
    BlockVectorType                                 *normal_vector_field = nullptr;
    VectorType                                      *solution            = nullptr;
    const std::pair<double, double>                 *last_concentration_range = nullptr;
    bool                                            *first_reinit_step        = nullptr;
    ConditionalOStream                               pcout;
    TimeStepping                                     time_stepping;
    internal::DiagonalLink                           diagonal;
    std::vector<int>                                 iterations;
  };
} // namespace adaflo_b200
