// navier_stokes_matrix.h -- C++ host-side mirror of adaflo::NavierStokesMatrix<dim> on top of
// the C ABI (include/adaflo_b200.h).
//
// Same class shape, method names, argument meaning and error behaviour as the reference
// (include/adaflo/navier_stokes_matrix.h:57-204): solver templates that only call
//   void vmult(BlockVector &dst, const BlockVector &src) const
// can take this class in place of adaflo::NavierStokesMatrix.  deal.II is not needed here;
// where the reference takes deal.II types (MatrixFree, BlockVector, TimeStepping,
// FlowParameters) this header provides device-backed stand-ins with the same accessors.  The
// adaptor that fills MatrixFreeData from a dealii::MatrixFree is sketched in INTEGRATION.md.
//
// Errors: reference Assert/AssertThrow become exceptions (ExcNotInitialized,
// ExcNotImplemented), every failing ABI call throws ExcAdafloB200 with the library message.
#ifndef ADAFLO_B200_NAVIER_STOKES_MATRIX_H
#define ADAFLO_B200_NAVIER_STOKES_MATRIX_H

#include <array>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../adaflo_b200.h"

namespace adaflo_b200
{
  struct ExcAdafloB200 : std::runtime_error
  {
    using std::runtime_error::runtime_error;
  };
  struct ExcNotInitialized : ExcAdafloB200
  {
    ExcNotInitialized() : ExcAdafloB200("ExcNotInitialized: operator has no MatrixFree data") {}
  };
  struct ExcNotImplemented : ExcAdafloB200
  {
    using ExcAdafloB200::ExcAdafloB200;
  };

  inline void check(const int rc)
  {
    if (rc == ADAFLO_B200_OK)
      return;
    const std::string msg = adaflo_b200_last_error();
    if (rc == ADAFLO_B200_ERR_UNSUPPORTED)
      throw ExcNotImplemented(msg);
    if (rc == ADAFLO_B200_ERR_STATE)
      throw ExcNotInitialized();
    throw ExcAdafloB200(msg);
  }

  /// One rank / one GPU.
  class Context
  {
  public:
    explicit Context(const int device = 0, void *cuda_stream = nullptr) { check(adaflo_b200_ctx_create(&h, device, cuda_stream)); }
    ~Context() { adaflo_b200_ctx_destroy(h); }
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;
    adaflo_b200_ctx_t handle() const { return h; }
    void synchronize() const { check(adaflo_b200_ctx_synchronize(h)); }

  private:
    adaflo_b200_ctx_t h = nullptr;
  };

  /// LinearAlgebra::distributed::Vector<double> stand-in: owned entries, then ghosts, on the device.
  class Vector
  {
  public:
    Vector() = default;
    Vector(const Context &ctx, const std::uint64_t n_owned, const std::uint64_t n_local) { reinit(ctx, n_owned, n_local); }
    ~Vector() { clear(); }
    Vector(const Vector &) = delete;
    Vector &operator=(const Vector &) = delete;
    Vector(Vector &&o) noexcept { swap(o); }
    Vector &operator=(Vector &&o) noexcept
    {
      swap(o);
      return *this;
    }
    void reinit(const Context &c, const std::uint64_t n_owned_in, const std::uint64_t n_local_in)
    {
      clear();
      ctx = &c;
      n_owned = n_owned_in;
      n_local = n_local_in;
      check(adaflo_b200_vec_alloc(c.handle(), n_local, &data));
    }
    void clear()
    {
      if (data)
        adaflo_b200_vec_free(ctx->handle(), data);
      data = nullptr;
    }
    std::uint64_t locally_owned_size() const { return n_owned; }
    std::uint64_t size_with_ghosts() const { return n_local; }
    double *begin() { return data; }
    const double *begin() const { return data; }
    void copy_from_host(const double *host) { check(adaflo_b200_vec_upload(ctx->handle(), data, host, n_owned)); }
    void copy_to_host(double *host) const { check(adaflo_b200_vec_download(ctx->handle(), host, data, n_owned)); }
    Vector &operator=(const double v)
    {
      check(adaflo_b200_vec_set(ctx->handle(), data, v, n_local));
      return *this;
    }
    void add(const double a, const Vector &x) { check(adaflo_b200_vec_axpy(ctx->handle(), data, a, x.data, n_owned)); }
    void sadd(const double s, const double a, const Vector &x) { check(adaflo_b200_vec_sadd(ctx->handle(), data, s, a, x.data, n_owned)); }
    void scale(const Vector &factors) { check(adaflo_b200_vec_scale(ctx->handle(), data, factors.data, n_owned)); }
    double operator*(const Vector &x) const
    {
      double r = 0;
      check(adaflo_b200_vec_dot(ctx->handle(), data, x.data, n_owned, &r));
      return r;
    }
    double add_and_dot(const double a, const Vector &x, const Vector &z)
    {
      double r = 0;
      check(adaflo_b200_vec_add_and_dot(ctx->handle(), data, a, x.data, z.data, n_owned, &r));
      return r;
    }
    void swap(Vector &o)
    {
      std::swap(ctx, o.ctx);
      std::swap(data, o.data);
      std::swap(n_owned, o.n_owned);
      std::swap(n_local, o.n_local);
    }

  private:
    const Context *ctx     = nullptr;
    double        *data    = nullptr;
    std::uint64_t  n_owned = 0, n_local = 0;
  };

  /// LinearAlgebra::distributed::BlockVector<double> stand-in with the blocks (velocity, pressure).
  class BlockVector
  {
  public:
    Vector       &block(const unsigned int i) { return blocks[i]; }
    const Vector &block(const unsigned int i) const { return blocks[i]; }
    unsigned int  n_blocks() const { return 2; }
    BlockVector  &operator=(const double v)
    {
      blocks[0] = v;
      blocks[1] = v;
      return *this;
    }
    double operator*(const BlockVector &x) const { return blocks[0] * x.blocks[0] + blocks[1] * x.blocks[1]; }

  private:
    std::array<Vector, 2> blocks;
  };

  /// The slice of adaflo::FlowParameters the operator reads (include/adaflo/parameters.h:54-86).
  struct FlowParameters
  {
    enum PhysicalType
    {
      incompressible,
      incompressible_stationary,
      stokes
    } physical_type = incompressible;
    enum Linearization
    {
      coupled_implicit_newton,
      coupled_implicit_picard,
      coupled_velocity_semi_implicit,
      coupled_velocity_explicit,
      projection
    } linearization                                = coupled_implicit_newton;
    double beta_convective_term_momentum_balance = 0.5;
    bool   augmented_taylor_hood                 = false;
    double viscosity = 1., density = 1., damping = 0., tau_grad_div = 0.;
    double density_diff = 0., viscosity_diff = 0.;

    adaflo_b200_flow_params to_abi() const
    {
      adaflo_b200_flow_params p{};
      p.physical_type                         = static_cast<int>(physical_type);
      p.linearization                         = static_cast<int>(linearization);
      p.augmented_taylor_hood                 = augmented_taylor_hood;
      p.density                               = density;
      p.viscosity                             = viscosity;
      p.density_diff                          = density_diff;
      p.viscosity_diff                        = viscosity_diff;
      p.damping                               = damping;
      p.tau_grad_div                          = tau_grad_div;
      p.beta_convective_term_momentum_balance = beta_convective_term_momentum_balance;
      return p;
    }
  };

  /// adaflo::TimeSteppingParameters (include/adaflo/time_stepping.h:34-55)
  struct TimeSteppingParameters
  {
    enum class Scheme
    {
      implicit_euler,
      explicit_euler,
      crank_nicolson,
      bdf_2
    };
    Scheme time_step_scheme     = Scheme::bdf_2;
    double start_time           = 0.;
    double end_time             = 1.;
    double time_step_size_start = 1.;
    double time_stepping_cfl    = 0.8;
    double time_stepping_coef2  = 10.;
    double time_step_tolerance  = 1e-2;
    double time_step_size_max   = 1.;
    double time_step_size_min   = 1.;
  };

  /// adaflo::TimeStepping: the accessors the operators call (include/adaflo/time_stepping.h:240-295) and the step
  /// control of source/time_stepping.cc:113-192 (next(): BDF-2 weights for variable step sizes, extrapolation factors).
  struct TimeStepping
  {
    double weight_val = 1., weight_old_val = -1., weight_old_old_val = 0., tau1_val = 1.;
    double factor_extrapol_old = 1., factor_extrapol_old_old = 0.;
    TimeStepping() = default;
    explicit TimeStepping(const TimeSteppingParameters &p)
      : weight_val(1. / p.time_step_size_start)
      , weight_old_val(-1.)
      , factor_extrapol_old(0.)
      , scheme(p.time_step_scheme)
      , start_val(p.start_time)
      , final_val(p.end_time)
      , start_step_val(p.time_step_size_start)
      , max_step_val(p.time_step_size_max)
      , current_step_val(p.time_step_size_start)
      , step_val(p.time_step_size_start)
      , now_val(p.start_time)
      , prev_val(p.start_time)
    {
      tau1_val = scheme == TimeSteppingParameters::Scheme::explicit_euler ? 0. : (scheme == TimeSteppingParameters::Scheme::crank_nicolson ? .5 : 1.);
    }
    double weight() const { return weight_val; }
    double weight_old() const { return weight_old_val; }
    double weight_old_old() const { return weight_old_old_val; }
    double tau1() const { return tau1_val; }
    double tau2() const { return 1. - tau1_val; }
    double start() const { return start_val; }
    double now() const { return now_val; }
    double previous() const { return prev_val; }
    double step_size() const { return current_step_val; }
    unsigned int step_no() const { return step_no_val; }
    bool at_end() const { return at_end_val; }
    void set_desired_time_step(const double dt) { step_val = dt; current_step_val = dt; }
    /// time_stepping.cc:113-192
    double next()
    {
      double s = current_step_val;
      if (now_val != start_val)
        {
          last_step_val = current_step_val;
          if (scheme == TimeSteppingParameters::Scheme::bdf_2 && step_no_val == 1)
            s = step_val;
          if (s > max_step_val)
            s = max_step_val;
        }
      double h         = now_val + s;
      current_step_val = s;
      if (!at_end_val && h > final_val - .01 * s)
        {
          current_step_val = final_val - now_val;
          h                = final_val;
          at_end_val       = true;
        }
      if (scheme == TimeSteppingParameters::Scheme::bdf_2 && now_val != start_val)
        {
          weight_val         = (2. * current_step_val + last_step_val) / (current_step_val * (current_step_val + last_step_val));
          weight_old_val     = -((current_step_val + last_step_val) / (current_step_val * last_step_val));
          weight_old_old_val = current_step_val / (last_step_val * (current_step_val + last_step_val));
        }
      else
        {
          weight_val     = 1. / current_step_val;
          weight_old_val = -1. / current_step_val;
        }
      if (step_no_val > 1)
        {
          factor_extrapol_old     = (current_step_val + last_step_val) / last_step_val;
          factor_extrapol_old_old = -current_step_val / last_step_val;
        }
      else
        {
          factor_extrapol_old     = 1.;
          factor_extrapol_old_old = 0.;
        }
      prev_val = now_val;
      now_val  = h;
      ++step_no_val;
      return now_val;
    }
    adaflo_b200_time_weights to_abi() const
    {
      return adaflo_b200_time_weights{weight_val, weight_old_val, weight_old_old_val, tau1_val, factor_extrapol_old, factor_extrapol_old_old};
    }

  private:
    TimeSteppingParameters::Scheme scheme = TimeSteppingParameters::Scheme::bdf_2;
    double       start_val = 0., final_val = 1e300, start_step_val = 1., max_step_val = 1e300, current_step_val = 1., last_step_val = 0., step_val = 1.;
    double       now_val = 0., prev_val = 0.;
    unsigned int step_no_val = 0;
    bool         at_end_val  = false;
  };

  /// Device mirror of dealii::MatrixFree<dim,double> (what MatrixFree::reinit produced on the host).
  class MatrixFreeData
  {
  public:
    MatrixFreeData(const Context &ctx, const adaflo_b200_mf_desc &desc) : ctx(ctx) { check(adaflo_b200_mf_create(ctx.handle(), &desc, &h)); }
    ~MatrixFreeData() { adaflo_b200_mf_destroy(h); }
    MatrixFreeData(const MatrixFreeData &) = delete;
    MatrixFreeData &operator=(const MatrixFreeData &) = delete;
    adaflo_b200_mf_t handle() const { return h; }
    const Context   &context() const { return ctx; }
    /// MatrixFree::initialize_dof_vector
    void initialize_dof_vector(Vector &v, const unsigned int dof_index) const
    {
      std::uint64_t n_owned = 0, n_local = 0;
      check(adaflo_b200_mf_vector_size(h, static_cast<int>(dof_index), &n_owned, &n_local));
      v.reinit(ctx, n_owned, n_local);
    }

  private:
    const Context   &ctx;
    adaflo_b200_mf_t h = nullptr;
  };

  template <int dim>
  class NavierStokesMatrix
  {
  public:
    static_assert(dim == 2 || dim == 3, "dim 2 and 3 are implemented (the reference also instantiates dim 1, source/navier_stokes_matrix.cc:1210)");

    NavierStokesMatrix(const FlowParameters &parameters,
                       const unsigned int   &dof_index_u,
                       const unsigned int   &dof_index_p,
                       const unsigned int   &quad_index_u,
                       const unsigned int   &quad_index_p,
                       const BlockVector    &solution_old,
                       const BlockVector    &solution_old_old)
      : parameters(parameters)
      , dof_index_u(dof_index_u)
      , dof_index_p(dof_index_p)
      , quad_index_u(quad_index_u)
      , quad_index_p(quad_index_p)
      , solution_old(solution_old)
      , solution_old_old(solution_old_old)
    {}
    ~NavierStokesMatrix() { clear(); }

    void initialize(const MatrixFreeData &matrix_free_in, const TimeStepping &time_stepping_in, const bool pressure_average_fix)
    {
      clear();
      matrix_free   = &matrix_free_in;
      time_stepping = &time_stepping_in;
      const adaflo_b200_flow_params p = parameters.to_abi();
      check(adaflo_b200_ns_create(matrix_free_in.context().handle(), matrix_free_in.handle(), &p, dof_index_u, dof_index_p,
                                  quad_index_u, quad_index_p, pressure_average_fix, &h));
      check(adaflo_b200_ns_set_old_solutions(h, solution_old.block(0).begin(), solution_old_old.block(0).begin()));
    }

    void clear()
    {
      if (h)
        adaflo_b200_ns_destroy(h);
      h             = nullptr;
      matrix_free   = nullptr;
      time_stepping = nullptr;
    }

    void vmult(BlockVector &dst, const BlockVector &src) const
    {
      sync_parameters();
      check(adaflo_b200_ns_vmult(h, dst.block(0).begin(), dst.block(1).begin(), src.block(0).begin(), src.block(1).begin()));
    }
    void Tvmult(BlockVector &dst, const BlockVector &src) const
    {
      sync_parameters();
      check(adaflo_b200_ns_Tvmult(h, dst.block(0).begin(), dst.block(1).begin(), src.block(0).begin(), src.block(1).begin()));
    }
    void residual(BlockVector &system_rhs, const BlockVector &src, const BlockVector &user_rhs) const
    {
      sync_parameters();
      check(adaflo_b200_ns_set_old_solutions(h, solution_old.block(0).begin(), solution_old_old.block(0).begin()));
      check(adaflo_b200_ns_residual(h, system_rhs.block(0).begin(), system_rhs.block(1).begin(), src.block(0).begin(),
                                    src.block(1).begin(), user_rhs.block(0).begin(), user_rhs.block(1).begin()));
    }
    void divergence_vmult_add(Vector &dst, const Vector &src, const bool weight_by_viscosity = false) const
    {
      sync_parameters();
      check(adaflo_b200_ns_divergence_vmult_add(h, dst.begin(), src.begin(), weight_by_viscosity));
    }
    void velocity_vmult(Vector &dst, const Vector &src) const
    {
      sync_parameters();
      check(adaflo_b200_ns_velocity_vmult(h, dst.begin(), src.begin()));
    }
    void pressure_poisson_vmult(Vector &dst, const Vector &src) const
    {
      sync_parameters();
      check(adaflo_b200_ns_pressure_poisson_vmult(h, dst.begin(), src.begin()));
    }
    void pressure_mass_vmult(Vector &dst, const Vector &src) const
    {
      sync_parameters();
      check(adaflo_b200_ns_pressure_mass_vmult(h, dst.begin(), src.begin()));
    }
    void pressure_convdiff_vmult(Vector &dst, const Vector &src) const
    {
      sync_parameters();
      check(adaflo_b200_ns_pressure_convdiff_vmult(h, dst.begin(), src.begin()));
    }
    void fix_linearization_point() const
    {
      require();
      check(adaflo_b200_ns_fix_linearization_point(h));
    }
    void apply_pressure_average_projection(Vector &vector) const
    {
      sync_parameters();
      check(adaflo_b200_ns_apply_pressure_average_projection(h, vector.begin()));
    }
    void apply_pressure_shift(const double shift, Vector &pressure) const
    {
      require();
      check(adaflo_b200_ns_apply_pressure_shift(h, shift, pressure.begin()));
    }

    /// Raw device pointers into the decade-blocked coefficient arrays (layout: include/adaflo_b200.h).
    double *begin_densities() const { return coefficient(ADAFLO_B200_COEF_DENSITY); }
    double *begin_viscosities() const { return coefficient(ADAFLO_B200_COEF_VISCOSITY); }
    double *begin_damping_coeff() const { return coefficient(ADAFLO_B200_COEF_DAMPING); }
    double *begin_linearized_velocities() const { return coefficient(ADAFLO_B200_COEF_LINEARIZED_VELOCITY); }
    bool    use_variable_coefficients() const { return parameters.density_diff != 0 || parameters.viscosity_diff != 0; }

    /// The reference's accessors begin_densities(macro_cell) & co. (include/adaflo/navier_stokes_matrix.h:162-178): a pointer to
    /// the n_q VectorizedArray<double> of one cell batch, i.e. n_q * n_lanes doubles [q][lane] (the linearisation point:
    /// [q][12 rows][lane]).  They address a HOST mirror in the reference's storage order; external writers
    /// (LevelSetOKZSolver::local_compute_force level_set_okz.cc:351-364, compute_initial_stokes_field navier_stokes.cc:
    /// 1178-1196) fill it batch by batch and call update_coefficients_on_device() once.  Reading accessors download first.
    /// n_lanes = VectorizedArray<double>::size() of the host build; slot_cell maps slot batch * n_lanes + lane to the
    /// library's cell (empty: identity; 0xffffffff: unused lane), e.g. AdaptedMatrixFree::cell_of.
    void set_cell_batches(const unsigned int n_lanes_in, const std::vector<std::uint32_t> &slot_cell_in = std::vector<std::uint32_t>())
    {
      require();
      n_lanes   = n_lanes_in;
      slot_cell = slot_cell_in;
      std::uint32_t n_cells = 0;
      check(adaflo_b200_mf_n_cells(matrix_free->handle(), &n_cells));
      n_batches = slot_cell.empty() ? (n_cells + n_lanes - 1) / n_lanes : static_cast<std::uint32_t>(slot_cell.size() / n_lanes);
      for (auto &m : host_coefficients) m.clear();
    }
    double       *begin_densities(const unsigned int macro_cell) { return host_batch(ADAFLO_B200_COEF_DENSITY, macro_cell, true); }
    const double *begin_densities(const unsigned int macro_cell) const { return host_batch(ADAFLO_B200_COEF_DENSITY, macro_cell, false); }
    double       *begin_viscosities(const unsigned int macro_cell) { return host_batch(ADAFLO_B200_COEF_VISCOSITY, macro_cell, true); }
    const double *begin_viscosities(const unsigned int macro_cell) const { return host_batch(ADAFLO_B200_COEF_VISCOSITY, macro_cell, false); }
    double       *begin_damping_coeff(const unsigned int macro_cell) { return host_batch(ADAFLO_B200_COEF_DAMPING, macro_cell, true); }
    const double *begin_damping_coeff(const unsigned int macro_cell) const { return host_batch(ADAFLO_B200_COEF_DAMPING, macro_cell, false); }
    const double *begin_linearized_velocities(const unsigned int macro_cell) const
    {
      return host_batch(ADAFLO_B200_COEF_LINEARIZED_VELOCITY, macro_cell, false);
    }
    /// upload the batches handed out for writing since the last call
    void update_coefficients_on_device()
    {
      for (int which = 0; which < 4; ++which)
        if (host_dirty[which])
          {
            check(adaflo_b200_ns_set_coefficient_canonical(h, which, host_coefficients[which].data(), 1, n_batches, n_lanes,
                                                           slot_cell.empty() ? nullptr : slot_cell.data()));
            host_dirty[which] = false;
          }
    }
    /// forget the host mirrors (the device arrays have been rewritten, e.g. by residual() or the level-set force kernel)
    void invalidate_host_coefficients() const
    {
      for (auto &m : host_coefficients) m.clear();
    }

    /// (accumulated wall time, number of calls) since the last query; resets the counters.
    /// the C-ABI handle (for the level-set force, which writes rho / mu into this operator)
    adaflo_b200_ns_t handle() const
    {
      if (!h)
        throw ExcNotInitialized();
      return h;
    }

    std::pair<double, unsigned int> get_matvec_statistics() const
    {
      require();
      std::uint64_t n = 0;
      double        s = 0;
      check(adaflo_b200_ns_get_matvec_statistics(h, &n, &s));
      return {s, static_cast<unsigned int>(n)};
    }

    const MatrixFreeData &get_matrix_free() const
    {
      require();
      return *matrix_free;
    }
    const TimeStepping &get_time_stepping() const
    {
      require();
      return *time_stepping;
    }
    void initialize_u_vector(Vector &vec) const { get_matrix_free().initialize_dof_vector(vec, dof_index_u); }
    void initialize_p_vector(Vector &vec) const { get_matrix_free().initialize_dof_vector(vec, dof_index_p); }
    unsigned int pressure_degree() const { return 1; }

  private:
    void require() const
    {
      if (!h)
        throw ExcNotInitialized();
    }
    // parameters and time weights are held by reference in the reference class and may change
    // between calls (time step changes, compute_initial_stokes_field): forward them every time
    void sync_parameters() const
    {
      require();
      const adaflo_b200_flow_params  p = parameters.to_abi();
      const adaflo_b200_time_weights w = time_stepping->to_abi();
      check(adaflo_b200_ns_set_flow_params(h, &p));
      check(adaflo_b200_ns_set_time_weights(h, &w));
    }
    double *coefficient(const int which) const
    {
      require();
      double       *p = nullptr;
      std::uint64_t n = 0;
      check(adaflo_b200_ns_coefficient_ptr(h, which, &p, &n));
      return p;
    }

    const FlowParameters &parameters;
    const unsigned int    dof_index_u, dof_index_p, quad_index_u, quad_index_p;
    const BlockVector    &solution_old;
    const BlockVector    &solution_old_old;
    const MatrixFreeData *matrix_free   = nullptr;
    const TimeStepping   *time_stepping = nullptr;
    adaflo_b200_ns_t      h             = nullptr;
    // host mirrors of the per-point coefficient arrays in the reference's [batch][q][(row)][lane] order
    unsigned int                       n_lanes   = 8;
    std::uint32_t                      n_batches = 0;
    std::vector<std::uint32_t>         slot_cell;
    mutable std::vector<double>        host_coefficients[4];
    mutable bool                       host_dirty[4] = {false, false, false, false};
    double *host_batch(const int which, const unsigned int macro_cell, const bool for_writing) const
    {
      require();
      if (n_batches == 0)
        throw ExcAdafloB200("begin_*(macro_cell): call set_cell_batches(n_lanes) first");
      const std::size_t n_rows = which == ADAFLO_B200_COEF_LINEARIZED_VELOCITY ? 12 : 1, per_batch = 27 * n_rows * n_lanes;
      std::vector<double> &m = host_coefficients[which];
      if (m.empty())
        {  // first access since the device arrays changed: fetch them
          m.assign(per_batch * n_batches, 0.);
          check(adaflo_b200_ns_get_coefficient_canonical(h, which, m.data(), 1, n_batches, n_lanes, slot_cell.empty() ? nullptr : slot_cell.data()));
        }
      if (macro_cell >= n_batches)
        throw ExcAdafloB200("begin_*(macro_cell): cell batch out of range");
      if (for_writing)
        host_dirty[which] = true;
      return m.data() + per_batch * macro_cell;
    }
  };

  /// adaflo::DiagonalPreconditioner<double> (source/diagonal_preconditioner.cc:38-124).
  class DiagonalPreconditioner
  {
  public:
    explicit DiagonalPreconditioner(const Context &ctx) : ctx(ctx) {}
    DiagonalPreconditioner(const Context &ctx, const Vector &diagonal_vector_in) : ctx(ctx) { reinit(diagonal_vector_in); }
    void reinit(const Vector &diagonal_vector_in)
    {
      const std::uint64_t n_owned = diagonal_vector_in.locally_owned_size(), n_local = diagonal_vector_in.size_with_ghosts();
      diagonal_vector.reinit(ctx, n_owned, n_local);
      check(adaflo_b200_vec_axpy(ctx.handle(), diagonal_vector.begin(), 1.0, diagonal_vector_in.begin(), n_owned));  // copy (vectors are zero initialised)
      inverse_diagonal_vector.reinit(ctx, n_owned, n_local);
      check(adaflo_b200_jacobi_reinit(ctx.handle(), inverse_diagonal_vector.begin(), diagonal_vector_in.begin(), n_owned));
      ++generation_val;
    }
    void vmult(Vector &dst, const Vector &src) const
    {
      check(adaflo_b200_jacobi_vmult(ctx.handle(), dst.begin(), src.begin(), inverse_diagonal_vector.begin(),
                                     inverse_diagonal_vector.locally_owned_size(), 1));
    }
    const Vector &get_inverse_vector() const { return inverse_diagonal_vector; }
    const Vector &get_vector() const { return diagonal_vector; }
    /// incremented by every reinit: users that hand the diagonal to the library re-send it when this changes
    unsigned int generation() const { return generation_val; }

  private:
    const Context &ctx;
    Vector         diagonal_vector, inverse_diagonal_vector;
    unsigned int   generation_val = 0;
  };
} // namespace adaflo_b200

#endif
