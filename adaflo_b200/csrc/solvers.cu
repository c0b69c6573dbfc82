// solvers.cu -- Krylov drivers on device block vectors behind the C ABI: the loops deal.II's SolverCG /
// SolverBicgstab / SolverFGMRES run around the operators of this library.  Reference call sites:
//   source/level_set_okz_compute_normal.cc:187-217, level_set_okz_compute_curvature.cc:300-330   (CG, projections)
//   source/level_set_okz_reinitialization.cc:334-336                                           (CG)
//   source/level_set_okz_advance_concentration.cc:626-644                                      (BiCGStab / GMRES)
//   source/navier_stokes.cc:571-609                                                            (FGMRES, block preconditioner)
// The host drives the iteration; every vector operation is a fused kernel of vector_ops.cu and every reduction is one
// small D2H copy (all ranks reduce through the context's communicator).  Stopping rule = deal.II ReductionControl:
// stop when ||r|| <= max(abs_tol, rel_tol ||r_0||).  Work vectors come from a per-context pool (no cudaMalloc per solve).
#include <cmath>
#include <map>
#include <mutex>
#include <vector>

#include "common.cuh"

using namespace ab200;

namespace {
struct PoolEntry {
  double *ptr;
  size_t doubles;
  bool used;
};
std::mutex g_pool_mutex;
std::map<adaflo_b200_ctx_t, std::vector<PoolEntry>> g_pool;

}  // namespace
namespace ab200 {
double *pool_get(adaflo_b200_ctx_t ctx, size_t doubles)
{
  std::lock_guard<std::mutex> lock(g_pool_mutex);
  auto &v = g_pool[ctx];
  PoolEntry *best = nullptr;
  for (auto &e : v)
    if (!e.used && e.doubles >= doubles && (!best || e.doubles < best->doubles)) best = &e;
  if (best) {
    best->used = true;
    return best->ptr;
  }
  double *p = nullptr;
  if (cudaMalloc((void **)&p, std::max<size_t>(1, doubles) * sizeof(double)) != cudaSuccess) return nullptr;
  v.push_back({p, doubles, true});
  return p;
}
void pool_put(adaflo_b200_ctx_t ctx, double *p)
{
  std::lock_guard<std::mutex> lock(g_pool_mutex);
  for (auto &e : g_pool[ctx])
    if (e.ptr == p) e.used = false;
}
}  // namespace ab200
namespace {

// a block vector of the solver: n_blocks device arrays (owned + ghost entries each)
struct BVec {
  adaflo_b200_ctx_t ctx = nullptr;
  int nb = 0;
  double *b[4] = {nullptr, nullptr, nullptr, nullptr};
  bool ok = true;
  BVec() = default;
  BVec(adaflo_b200_ctx_t c, int n_blocks, const uint64_t *n_local) : ctx(c), nb(n_blocks)
  {
    for (int i = 0; i < nb; ++i) {
      b[i] = pool_get(ctx, n_local[i]);
      ok = ok && b[i];
    }
  }
  BVec(const BVec &) = delete;
  BVec &operator=(const BVec &) = delete;
  BVec(BVec &&o) noexcept { *this = std::move(o); }
  BVec &operator=(BVec &&o) noexcept
  {
    release();
    ctx = o.ctx; nb = o.nb; ok = o.ok;
    for (int i = 0; i < 4; ++i) { b[i] = o.b[i]; o.b[i] = nullptr; }
    return *this;
  }
  void release()
  {
    for (int i = 0; i < 4; ++i)
      if (b[i]) { pool_put(ctx, b[i]); b[i] = nullptr; }
  }
  ~BVec() { release(); }
};

struct Sys {
  adaflo_b200_ctx_t ctx;
  int nb;
  const uint64_t *n_owned, *n_local;
  adaflo_b200_apply_fn A, P;
  void *A_user, *P_user;

  int apply_A(double *const *dst, const double *const *src) const { return A(A_user, dst, src); }
  int apply_P(double *const *dst, const double *const *src) const
  {
    if (P) return P(P_user, dst, src);
    for (int i = 0; i < nb; ++i) AB_CUDA(cudaMemcpyAsync(dst[i], src[i], n_owned[i] * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    return 0;
  }
  int copy(double *const *dst, const double *const *src) const
  {
    for (int i = 0; i < nb; ++i) AB_CUDA(cudaMemcpyAsync(dst[i], src[i], n_owned[i] * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    return 0;
  }
  int zero(double *const *v) const
  {
    for (int i = 0; i < nb; ++i) AB_CUDA(cudaMemsetAsync(v[i], 0, n_local[i] * sizeof(double), ctx->stream));
    return 0;
  }
  int axpy(double *const *y, double a, const double *const *x) const
  {
    for (int i = 0; i < nb; ++i) AB_CHECK(adaflo_b200_vec_axpy(ctx, y[i], a, x[i], n_owned[i]));
    return 0;
  }
  int sadd(double *const *y, double s, double a, const double *const *x) const
  {
    for (int i = 0; i < nb; ++i) AB_CHECK(adaflo_b200_vec_sadd(ctx, y[i], s, a, x[i], n_owned[i]));
    return 0;
  }
  // out[j] = <x, ys[j]> summed over the blocks (and ranks), 64 dots per reduction pass
  int dots(const double *const *x, int k, const double *const *const *ys, double *out) const
  {
    for (int j = 0; j < k; ++j) out[j] = 0;
    for (int blk = 0; blk < nb; ++blk)
      for (int j0 = 0; j0 < k; j0 += 64) {
        const int kk = std::min(64, k - j0);
        const double *ptr[64];
        double part[64];
        for (int j = 0; j < kk; ++j) ptr[j] = ys[j0 + j][blk];
        AB_CHECK(adaflo_b200_vec_multi_dot(ctx, x[blk], ptr, kk, n_owned[blk], part));
        for (int j = 0; j < kk; ++j) out[j0 + j] += part[j];
      }
    return 0;
  }
  int dot(const double *const *x, const double *const *y, double *out) const
  {
    const double *const *ys[1] = {y};
    return dots(x, 1, ys, out);
  }
  // y += sum_j a[j] xs[j]
  int maxpy(double *const *y, int k, const double *a, const double *const *const *xs) const
  {
    for (int blk = 0; blk < nb; ++blk)
      for (int j0 = 0; j0 < k; j0 += 64) {
        const int kk = std::min(64, k - j0);
        const double *ptr[64];
        for (int j = 0; j < kk; ++j) ptr[j] = xs[j0 + j][blk];
        AB_CHECK(adaflo_b200_vec_multi_axpy(ctx, y[blk], a + j0, ptr, kk, n_owned[blk]));
      }
    return 0;
  }
};

int check_args(adaflo_b200_ctx_t ctx, int nb, const uint64_t *n_owned, const uint64_t *n_local, adaflo_b200_apply_fn A, double *const *x,
               const double *const *b, adaflo_b200_solver_control *c)
{
  if (!ctx || !n_owned || !n_local || !A || !x || !b || !c) return fail(ADAFLO_B200_ERR_ARG, "solver: null argument");
  if (nb < 1 || nb > 4) return fail(ADAFLO_B200_ERR_ARG, "solver: 1..4 blocks");
  for (int i = 0; i < nb; ++i)
    if (!x[i] || !b[i] || n_local[i] < n_owned[i]) return fail(ADAFLO_B200_ERR_ARG, "solver: bad block");
  return 0;
}
inline double target(const adaflo_b200_solver_control *c, double r0) { return std::max(c->abs_tol, c->rel_tol * r0); }
inline void record(adaflo_b200_solver_control *c, double res)
{
  if (c->history && c->history_size < c->history_capacity) c->history[c->history_size++] = res;
}
int finish(adaflo_b200_solver_control *c, int it, double res, bool converged, const char *name)
{
  c->last_step = it;
  c->last_value = res;
  if (!converged) return fail(ADAFLO_B200_ERR_NO_CONVERGENCE, std::string(name) + ": no convergence after " + std::to_string(it) + " iterations");
  return 0;
}
}  // namespace

namespace ab200 {
void solver_pool_release(adaflo_b200_ctx_t ctx)
{
  std::lock_guard<std::mutex> lock(g_pool_mutex);
  auto it = g_pool.find(ctx);
  if (it == g_pool.end()) return;
  for (auto &e : it->second) cudaFree(e.ptr);
  g_pool.erase(it);
}
}  // namespace ab200

extern "C" {

// preconditioned conjugate gradients (deal.II SolverCG)
int adaflo_b200_solve_cg(adaflo_b200_ctx_t ctx, int n_blocks, const uint64_t *n_owned, const uint64_t *n_local, adaflo_b200_apply_fn A, void *A_user,
                         adaflo_b200_apply_fn P, void *P_user, double *const *x, const double *const *b, adaflo_b200_solver_control *c)
{
  AB_CHECK(check_args(ctx, n_blocks, n_owned, n_local, A, x, b, c));
  AB_CUDA(cudaSetDevice(ctx->device));
  const Sys S{ctx, n_blocks, n_owned, n_local, A, P, A_user, P_user};
  BVec r(ctx, n_blocks, n_local), z(ctx, n_blocks, n_local), p(ctx, n_blocks, n_local), Ap(ctx, n_blocks, n_local);
  if (!r.ok || !z.ok || !p.ok || !Ap.ok) return fail(ADAFLO_B200_ERR_CUDA, "solve_cg: out of device memory");
  AB_CHECK(S.zero(p.b));
  AB_CHECK(S.zero(z.b));
  AB_CHECK(S.apply_A(Ap.b, x));
  AB_CHECK(S.copy(r.b, b));
  AB_CHECK(S.axpy(r.b, -1.0, Ap.b));  // r = b - A x
  double rr;
  AB_CHECK(S.dot(r.b, r.b, &rr));
  double res = std::sqrt(rr);
  c->initial_value = res;
  c->history_size = 0;
  record(c, res);
  const double tol = target(c, res);
  int it = 0;
  if (res > tol) {
    AB_CHECK(S.apply_P(z.b, r.b));
    AB_CHECK(S.copy(p.b, z.b));
    double rz;
    AB_CHECK(S.dot(r.b, z.b, &rz));
    for (;;) {
      ++it;
      AB_CHECK(S.apply_A(Ap.b, p.b));
      double pAp;
      AB_CHECK(S.dot(p.b, Ap.b, &pAp));
      if (pAp == 0.0) return finish(c, it, res, false, "solve_cg (breakdown)");
      const double alpha = rz / pAp;
      AB_CHECK(S.axpy(x, alpha, p.b));
      AB_CHECK(S.axpy(r.b, -alpha, Ap.b));
      AB_CHECK(S.dot(r.b, r.b, &rr));
      res = std::sqrt(rr);
      record(c, res);
      if (res <= tol) break;
      if (it >= c->max_iter) return finish(c, it, res, false, "solve_cg");
      AB_CHECK(S.apply_P(z.b, r.b));
      double rz_new;
      AB_CHECK(S.dot(r.b, z.b, &rz_new));
      AB_CHECK(S.sadd(p.b, rz_new / rz, 1.0, z.b));  // p = z + beta p
      rz = rz_new;
    }
  }
  return finish(c, it, res, true, "solve_cg");
}

// BiCGStab with preconditioned directions: the control flow of deal.II's SolverBicgstab::iterate with
// exact_residual = false (recurrence residual, a convergence check after the first half step), which is what the
// reference's advection solve uses (level_set_okz_advance_concentration.cc:626-644) -- iteration counts are comparable
int adaflo_b200_solve_bicgstab(adaflo_b200_ctx_t ctx, int n_blocks, const uint64_t *n_owned, const uint64_t *n_local, adaflo_b200_apply_fn A,
                               void *A_user, adaflo_b200_apply_fn P, void *P_user, double *const *x, const double *const *b,
                               adaflo_b200_solver_control *c)
{
  AB_CHECK(check_args(ctx, n_blocks, n_owned, n_local, A, x, b, c));
  AB_CUDA(cudaSetDevice(ctx->device));
  const Sys S{ctx, n_blocks, n_owned, n_local, A, P, A_user, P_user};
  BVec r(ctx, n_blocks, n_local), rbar(ctx, n_blocks, n_local), p(ctx, n_blocks, n_local), v(ctx, n_blocks, n_local), y(ctx, n_blocks, n_local),
      z(ctx, n_blocks, n_local), t(ctx, n_blocks, n_local);
  if (!r.ok || !rbar.ok || !p.ok || !v.ok || !y.ok || !z.ok || !t.ok) return fail(ADAFLO_B200_ERR_CUDA, "solve_bicgstab: out of device memory");
  for (BVec *w : {&p, &y, &z, &t, &v}) AB_CHECK(S.zero(w->b));
  AB_CHECK(S.apply_A(r.b, x));
  AB_CHECK(S.sadd(r.b, -1.0, 1.0, b));  // r = b - A x
  AB_CHECK(S.copy(rbar.b, r.b));
  double rr;
  AB_CHECK(S.dot(r.b, r.b, &rr));
  double res = std::sqrt(rr);
  c->initial_value = res;
  c->history_size = 0;
  record(c, res);
  const double tol = target(c, res);
  double rho = 1, alpha = 1, omega = 1;
  bool startup = true;
  int it = 0;
  if (res <= tol) return finish(c, 0, res, true, "solve_bicgstab");
  for (;;) {
    ++it;
    double rhobar;
    AB_CHECK(S.dot(r.b, rbar.b, &rhobar));
    if (rho == 0.0 || omega == 0.0) return finish(c, it, res, false, "solve_bicgstab (breakdown)");
    const double beta = rhobar * alpha / (rho * omega);
    rho = rhobar;
    if (startup) {
      AB_CHECK(S.copy(p.b, r.b));
      startup = false;
    } else {
      AB_CHECK(S.sadd(p.b, beta, 1.0, r.b));
      AB_CHECK(S.axpy(p.b, -beta * omega, v.b));
    }
    AB_CHECK(S.apply_P(y.b, p.b));
    AB_CHECK(S.apply_A(v.b, y.b));
    AB_CHECK(S.dot(rbar.b, v.b, &rhobar));
    if (rhobar == 0.0) return finish(c, it, res, false, "solve_bicgstab (breakdown)");
    alpha = rho / rhobar;
    AB_CHECK(S.axpy(r.b, -alpha, v.b));
    AB_CHECK(S.dot(r.b, r.b, &rr));
    res = std::sqrt(rr);
    if (res <= tol) {  // early success: the second half step is not needed
      AB_CHECK(S.axpy(x, alpha, y.b));
      record(c, res);
      break;
    }
    AB_CHECK(S.apply_P(z.b, r.b));
    AB_CHECK(S.apply_A(t.b, z.b));
    double tt_tr[2];
    const double *const *ys[2] = {t.b, r.b};
    AB_CHECK(S.dots(t.b, 2, ys, tt_tr));
    if (tt_tr[0] == 0.0) return finish(c, it, res, false, "solve_bicgstab (breakdown)");
    omega = tt_tr[1] / tt_tr[0];
    AB_CHECK(S.axpy(x, alpha, y.b));
    AB_CHECK(S.axpy(x, omega, z.b));
    AB_CHECK(S.axpy(r.b, -omega, t.b));
    AB_CHECK(S.dot(r.b, r.b, &rr));
    res = std::sqrt(rr);
    record(c, res);
    if (res <= tol) break;
    if (it >= c->max_iter) return finish(c, it, res, false, "solve_bicgstab");
  }
  return finish(c, it, res, true, "solve_bicgstab");
}

// restarted flexible GMRES(max_basis), right preconditioned, classical Gram-Schmidt with one fused multi-dot and one
// fused multi-axpy per step, re-orthogonalisation on severe cancellation (deal.II SolverFGMRES / SolverGMRES, right)
int adaflo_b200_solve_gmres(adaflo_b200_ctx_t ctx, int n_blocks, const uint64_t *n_owned, const uint64_t *n_local, adaflo_b200_apply_fn A, void *A_user,
                            adaflo_b200_apply_fn P, void *P_user, double *const *x, const double *const *b, int max_basis,
                            adaflo_b200_solver_control *c)
{
  AB_CHECK(check_args(ctx, n_blocks, n_owned, n_local, A, x, b, c));
  if (max_basis < 1) return fail(ADAFLO_B200_ERR_ARG, "solve_gmres: max_basis < 1");
  AB_CUDA(cudaSetDevice(ctx->device));
  const Sys S{ctx, n_blocks, n_owned, n_local, A, P, A_user, P_user};
  const int m = max_basis;
  BVec w(ctx, n_blocks, n_local);
  if (!w.ok) return fail(ADAFLO_B200_ERR_CUDA, "solve_gmres: out of device memory");
  std::vector<BVec> V, Z;  // grown on demand (an early exit never allocates the full basis); reserved: the elements never move
  V.reserve(m + 1);
  Z.reserve(m + 1);
  auto basis = [&](std::vector<BVec> &B, size_t j) -> BVec * {
    while (B.size() <= j) {
      B.emplace_back(ctx, n_blocks, n_local);
      if (!B.back().ok) return nullptr;
      if (S.zero(B.back().b)) return nullptr;
    }
    return &B[j];
  };
  std::vector<double> H((size_t)(m + 1) * m, 0.0), g(m + 1), cs(m), sn(m), y(m), hw(m + 2), h(m + 1), h2(m + 1);
  auto Hat = [&](int i, int col) -> double & { return H[(size_t)col * (m + 1) + i]; };
  int it_total = 0;
  double tol = -1, res = 0;
  bool first = true;
  for (;;) {
    BVec *r = basis(V, 0);
    if (!r) return fail(ADAFLO_B200_ERR_CUDA, "solve_gmres: out of device memory");
    AB_CHECK(S.apply_A(w.b, x));
    AB_CHECK(S.copy(r->b, b));
    AB_CHECK(S.axpy(r->b, -1.0, w.b));
    double rr;
    AB_CHECK(S.dot(r->b, r->b, &rr));
    const double beta = std::sqrt(rr);
    if (first) {
      c->initial_value = beta;
      c->history_size = 0;
      record(c, beta);
      tol = target(c, beta);
      first = false;
    }
    res = beta;
    if (beta <= tol) break;
    AB_CHECK(S.sadd(r->b, 1.0 / beta, 0.0, r->b));
    g[0] = beta;
    int k_dim = 0;
    for (int j = 0; j < m; ++j) {
      ++it_total;
      if (P) {
        BVec *zj = basis(Z, j);
        if (!zj) return fail(ADAFLO_B200_ERR_CUDA, "solve_gmres: out of device memory");
        AB_CHECK(S.apply_P(zj->b, V[j].b));
        AB_CHECK(S.apply_A(w.b, zj->b));
      } else
        AB_CHECK(S.apply_A(w.b, V[j].b));
      std::vector<const double *const *> vs(j + 2);
      for (int i = 0; i <= j; ++i) vs[i] = V[i].b;
      vs[j + 1] = w.b;
      AB_CHECK(S.dots(w.b, j + 2, vs.data(), hw.data()));
      const double ww = hw[j + 1];
      for (int i = 0; i <= j; ++i) h[i] = -hw[i];
      AB_CHECK(S.maxpy(w.b, j + 1, h.data(), vs.data()));
      for (int i = 0; i <= j; ++i) h[i] = hw[i];
      double hn2;
      AB_CHECK(S.dot(w.b, w.b, &hn2));
      if (hn2 < 1e-2 * ww) {  // severe cancellation: second Gram-Schmidt pass
        AB_CHECK(S.dots(w.b, j + 1, vs.data(), h2.data()));
        for (int i = 0; i <= j; ++i) {
          h[i] += h2[i];
          h2[i] = -h2[i];
        }
        AB_CHECK(S.maxpy(w.b, j + 1, h2.data(), vs.data()));
        AB_CHECK(S.dot(w.b, w.b, &hn2));
      }
      const double hn = std::sqrt(std::max(hn2, 0.0));
      for (int i = 0; i <= j; ++i) Hat(i, j) = h[i];
      Hat(j + 1, j) = hn;
      for (int i = 0; i < j; ++i) {  // previous Givens rotations
        const double t0 = cs[i] * Hat(i, j) + sn[i] * Hat(i + 1, j);
        Hat(i + 1, j) = -sn[i] * Hat(i, j) + cs[i] * Hat(i + 1, j);
        Hat(i, j) = t0;
      }
      const double dnm = std::hypot(Hat(j, j), Hat(j + 1, j));
      k_dim = j + 1;
      if (dnm == 0.0) {  // w = 0 and no component along the basis: the Krylov space is exhausted
        k_dim = j;
        break;
      }
      cs[j] = Hat(j, j) / dnm;
      sn[j] = Hat(j + 1, j) / dnm;
      Hat(j, j) = dnm;
      Hat(j + 1, j) = 0.0;
      g[j + 1] = -sn[j] * g[j];
      g[j] = cs[j] * g[j];
      res = std::fabs(g[j + 1]);
      record(c, res);
      if (hn > 0 && j + 1 < m && res > tol && it_total < c->max_iter) {
        BVec *vn = basis(V, j + 1);
        if (!vn) return fail(ADAFLO_B200_ERR_CUDA, "solve_gmres: out of device memory");
        AB_CHECK(S.copy(vn->b, w.b));
        AB_CHECK(S.sadd(vn->b, 1.0 / hn, 0.0, vn->b));
      }
      if (res <= tol || it_total >= c->max_iter || hn == 0) break;
    }
    for (int i = k_dim - 1; i >= 0; --i) {
      double s = g[i];
      for (int cc = i + 1; cc < k_dim; ++cc) s -= Hat(i, cc) * y[cc];
      y[i] = s / Hat(i, i);
    }
    if (k_dim > 0) {
      std::vector<const double *const *> xs(k_dim);
      for (int i = 0; i < k_dim; ++i) xs[i] = P ? Z[i].b : V[i].b;
      AB_CHECK(S.maxpy(x, k_dim, y.data(), xs.data()));
    }
    if (res <= tol) break;
    if (it_total >= c->max_iter || k_dim == 0) return finish(c, it_total, res, false, "solve_gmres");
  }
  return finish(c, it_total, res, true, "solve_gmres");
}

// ---------------------------------------------------------------------------------------------------------------
// Ready-made adaflo_b200_apply_fn for the library's own operators: user = const adaflo_b200_operator_ref *
int adaflo_b200_apply_operator(void *user, double *const *dst, const double *const *src)
{
  const adaflo_b200_operator_ref *op = static_cast<const adaflo_b200_operator_ref *>(user);
  if (!op || !dst || !src) return fail(ADAFLO_B200_ERR_ARG, "apply_operator: null argument");
  switch (op->kind) {
    case ADAFLO_B200_OP_NS_VMULT: return adaflo_b200_ns_vmult((adaflo_b200_ns_t)op->handle, dst[0], dst[1], src[0], src[1]);
    case ADAFLO_B200_OP_NS_VELOCITY: return adaflo_b200_ns_velocity_vmult((adaflo_b200_ns_t)op->handle, dst[0], src[0]);
    case ADAFLO_B200_OP_NS_PRESSURE_POISSON: return adaflo_b200_ns_pressure_poisson_vmult((adaflo_b200_ns_t)op->handle, dst[0], src[0]);
    case ADAFLO_B200_OP_NS_PRESSURE_MASS: return adaflo_b200_ns_pressure_mass_vmult((adaflo_b200_ns_t)op->handle, dst[0], src[0]);
    case ADAFLO_B200_OP_LS_ADVECT: return adaflo_b200_ls_advance_concentration_vmult((adaflo_b200_ls_t)op->handle, dst[0], src[0]);
    case ADAFLO_B200_OP_LS_NORMAL: return adaflo_b200_ls_compute_normal_vmult((adaflo_b200_ls_t)op->handle, dst, src);
    case ADAFLO_B200_OP_LS_NORMAL_SCALAR: return adaflo_b200_ls_projection_vmult((adaflo_b200_ls_t)op->handle, dst[0], src[0]);
    case ADAFLO_B200_OP_LS_CURVATURE: return adaflo_b200_ls_compute_curvature_vmult((adaflo_b200_ls_t)op->handle, dst[0], src[0], op->flag);
    case ADAFLO_B200_OP_LS_REINIT: return adaflo_b200_ls_reinitialization_vmult((adaflo_b200_ls_t)op->handle, dst[0], src[0], op->flag);
    case ADAFLO_B200_OP_JACOBI: {  // DiagonalPreconditioner::vmult: the same inverse diagonal on every block
      for (int i = 0; i < op->n_blocks; ++i)
        AB_CHECK(adaflo_b200_jacobi_vmult((adaflo_b200_ctx_t)op->handle, dst[i], src[i], op->inv_diag, op->n, 1));
      return 0;
    }
    default: return fail(ADAFLO_B200_ERR_ARG, "apply_operator: unknown operator kind");
  }
}

}  // extern "C"
