// precond.cu -- device-resident block preconditioner of the linearised Navier-Stokes system:
// NavierStokesPreconditioner::vmult with do_inner_solves = true (source/navier_stokes_preconditioner.cc:593-737),
// lower block-triangular with the Cahouet-Chabard Schur complement approximation:
//   z_u  = A_u^-1 r_u                                   BiCGStab(100, 3e-2 ||r_u||) on velocity_vmult           (:634-668)
//   t    = -r_p + divergence_vmult_add(z_u)                                                                       (:673-675)
//   z_p  = M_p^-1 t + L_p^-1 t                          CG(100, 1e-2 rel) on pressure_mass_vmult (:712, :745-760),
//                                                       CG(30, 3e-2 ||t||) on pressure_poisson_vmult if density > 0 (:715-734)
// The reference preconditions the inner solves with Trilinos ILU / AMG built on the host; here they are matrix-free with
// Jacobi preconditioners (exact diagonal of the velocity block from compute_diagonal, lumped pressure mass), so the
// iteration counts of the reference's logs (which belong to its ILU / AMG) are not reproduced and the inner counts
// grow with 1 / h.  The inner solves are inexact: the outer solver must be flexible (adaflo_b200_solve_gmres is).
#include <cmath>
#include <new>

#include "common.cuh"

using namespace ab200;

struct adaflo_b200_precond_s {
  adaflo_b200_ns_t ns;
  int device;  // kept here: the preconditioner may outlive the operator it was built for (destruction order of a host language)
  double *inv_diag_u, *inv_lump_p, *ones_p, *temp, *temp2;
  uint64_t nu_owned, nu_local, np_owned, np_local;
  int inner_iterations[3];  // velocity, pressure mass, pressure Poisson (accumulated)
  int n_applications;
};

namespace {
int alloc(double **p, uint64_t n)
{
  AB_CUDA(cudaMalloc((void **)p, std::max<uint64_t>(1, n) * sizeof(double)));
  AB_CUDA(cudaMemset(*p, 0, std::max<uint64_t>(1, n) * sizeof(double)));
  return 0;
}
}  // namespace

extern "C" {

int adaflo_b200_ns_preconditioner_create(adaflo_b200_ns_t ns, adaflo_b200_precond_t *out)
{
  if (!ns || !out) return fail(ADAFLO_B200_ERR_ARG, "preconditioner_create: null argument");
  adaflo_b200_ctx_t ctx = ns->ctx;
  AB_CUDA(cudaSetDevice(ctx->device));
  adaflo_b200_precond_t P = new (std::nothrow) adaflo_b200_precond_s();
  if (!P) return fail(ADAFLO_B200_ERR_ARG, "out of host memory");
  P->ns = ns;
  P->device = ctx->device;
  const Space &su = ns->mf->spaces[ns->dof_u], &sp = ns->mf->spaces[ns->dof_p];
  P->nu_owned = su.n_owned_entries(); P->nu_local = su.n_local_entries();
  P->np_owned = sp.n_owned_entries(); P->np_local = sp.n_local_entries();
  for (int i = 0; i < 3; ++i) P->inner_iterations[i] = 0;
  P->n_applications = 0;
  double *diag_u = nullptr, *diag_p = nullptr;
  AB_CHECK(alloc(&P->inv_diag_u, P->nu_local));
  AB_CHECK(alloc(&P->inv_lump_p, P->np_local));
  AB_CHECK(alloc(&P->ones_p, P->np_local));
  AB_CHECK(alloc(&P->temp, P->np_local));
  AB_CHECK(alloc(&P->temp2, P->np_local));
  AB_CHECK(alloc(&diag_u, P->nu_local));
  AB_CHECK(alloc(&diag_p, P->np_local));
  AB_CHECK(adaflo_b200_ns_compute_diagonal(ns, diag_u, diag_p));
  AB_CHECK(adaflo_b200_jacobi_reinit(ctx, P->inv_diag_u, diag_u, P->nu_owned));
  // lumped pressure mass (row sums): spectrally equivalent to the diagonal on uniform meshes and positive everywhere
  AB_CHECK(adaflo_b200_vec_set(ctx, P->ones_p, 1.0, P->np_owned));
  AB_CHECK(adaflo_b200_ns_pressure_mass_vmult(ns, diag_p, P->ones_p));
  AB_CHECK(adaflo_b200_jacobi_reinit(ctx, P->inv_lump_p, diag_p, P->np_owned));
  AB_CUDA(cudaStreamSynchronize(ctx->stream));
  cudaFree(diag_u);
  cudaFree(diag_p);
  *out = P;
  return 0;
}

int adaflo_b200_ns_preconditioner_destroy(adaflo_b200_precond_t P)
{
  if (!P) return 0;
  cudaSetDevice(P->device);
  cudaDeviceSynchronize();
  cudaFree(P->inv_diag_u);
  cudaFree(P->inv_lump_p);
  cudaFree(P->ones_p);
  cudaFree(P->temp);
  cudaFree(P->temp2);
  delete P;
  return 0;
}

int adaflo_b200_ns_preconditioner_vmult(adaflo_b200_precond_t P, double *dst_u, double *dst_p, const double *src_u, const double *src_p)
{
  if (!P || !dst_u || !dst_p || !src_u || !src_p) return fail(ADAFLO_B200_ERR_ARG, "preconditioner_vmult: null argument");
  adaflo_b200_ns_t ns = P->ns;
  adaflo_b200_ctx_t ctx = ns->ctx;
  cudaStream_t s = ctx->stream;
  ++P->n_applications;
  // velocity block, :634-668
  {
    double nn = 0;
    AB_CHECK(adaflo_b200_vec_dot(ctx, src_u, src_u, P->nu_owned, &nn));
    AB_CUDA(cudaMemsetAsync(dst_u, 0, P->nu_local * sizeof(double), s));
    adaflo_b200_operator_ref A = {ADAFLO_B200_OP_NS_VELOCITY, ns, 0, nullptr, 0, 1};
    adaflo_b200_operator_ref J = {ADAFLO_B200_OP_JACOBI, ctx, 0, P->inv_diag_u, P->nu_owned, 1};
    adaflo_b200_solver_control c = {100, 3e-2 * std::sqrt(nn), 0.0, 0, 0, 0, nullptr, 0, 0};
    double *x[1] = {dst_u};
    const double *b[1] = {src_u};
    const int rc = adaflo_b200_solve_bicgstab(ctx, 1, &P->nu_owned, &P->nu_local, adaflo_b200_apply_operator, &A, adaflo_b200_apply_operator, &J, x, b, &c);
    P->inner_iterations[0] += c.last_step;
    if (rc && rc != ADAFLO_B200_ERR_NO_CONVERGENCE) return rc;  // the reference only asserts on non-convergence
  }
  // block-triangular part: temp = -r_p + B z_u, :673-675
  AB_CUDA(cudaMemsetAsync(P->temp, 0, P->np_local * sizeof(double), s));
  AB_CHECK(adaflo_b200_vec_axpy(ctx, P->temp, -1.0, src_p, P->np_owned));
  AB_CHECK(adaflo_b200_ns_divergence_vmult_add(ns, P->temp, dst_u, 0));
  // pressure mass, solve_pressure_mass :745-775: ReductionControl(100, 1e-50, 1e-2)
  {
    AB_CUDA(cudaMemsetAsync(dst_p, 0, P->np_local * sizeof(double), s));
    adaflo_b200_operator_ref A = {ADAFLO_B200_OP_NS_PRESSURE_MASS, ns, 0, nullptr, 0, 1};
    adaflo_b200_operator_ref J = {ADAFLO_B200_OP_JACOBI, ctx, 0, P->inv_lump_p, P->np_owned, 1};
    adaflo_b200_solver_control c = {100, 1e-50, 1e-2, 0, 0, 0, nullptr, 0, 0};
    double *x[1] = {dst_p};
    const double *b[1] = {P->temp};
    const int rc = adaflo_b200_solve_cg(ctx, 1, &P->np_owned, &P->np_local, adaflo_b200_apply_operator, &A, adaflo_b200_apply_operator, &J, x, b, &c);
    P->inner_iterations[1] += c.last_step;
    if (rc && rc != ADAFLO_B200_ERR_NO_CONVERGENCE) return rc;
  }
  // pressure Poisson, :715-734 (skipped for density == 0 like the reference)
  if (ns->prm.density > 0) {
    if (ns->has_mode[0]) {  // pure Neumann problem: make the right-hand side consistent (the reference pins one dof instead, :386-415)
      double sum = 0;
      AB_CHECK(adaflo_b200_vec_dot(ctx, P->temp, P->ones_p, P->np_owned, &sum));   // reduced over all ranks
      double n_glob = 0;
      AB_CHECK(adaflo_b200_vec_dot(ctx, P->ones_p, P->ones_p, P->np_owned, &n_glob));
      if (n_glob > 0) AB_CHECK(adaflo_b200_vec_axpy(ctx, P->temp, -sum / n_glob, P->ones_p, P->np_owned));
    }
    double nn = 0;
    AB_CHECK(adaflo_b200_vec_dot(ctx, P->temp, P->temp, P->np_owned, &nn));
    AB_CUDA(cudaMemsetAsync(P->temp2, 0, P->np_local * sizeof(double), s));
    adaflo_b200_operator_ref A = {ADAFLO_B200_OP_NS_PRESSURE_POISSON, ns, 0, nullptr, 0, 1};
    adaflo_b200_solver_control c = {30, 3e-2 * std::sqrt(nn), 0.0, 0, 0, 0, nullptr, 0, 0};
    double *x[1] = {P->temp2};
    const double *b[1] = {P->temp};
    const int rc = adaflo_b200_solve_cg(ctx, 1, &P->np_owned, &P->np_local, adaflo_b200_apply_operator, &A, nullptr, nullptr, x, b, &c);
    P->inner_iterations[2] += c.last_step;
    if (rc && rc != ADAFLO_B200_ERR_NO_CONVERGENCE) return rc;
    AB_CHECK(adaflo_b200_vec_axpy(ctx, dst_p, 1.0, P->temp2, P->np_owned));
  }
  return 0;
}

// adaflo_b200_apply_fn for the Krylov drivers: user = adaflo_b200_precond_t, blocks (u, p)
int adaflo_b200_ns_preconditioner_apply(void *user, double *const *dst, const double *const *src)
{
  if (!user || !dst || !src) return fail(ADAFLO_B200_ERR_ARG, "preconditioner_apply: null argument");
  return adaflo_b200_ns_preconditioner_vmult(static_cast<adaflo_b200_precond_t>(user), dst[0], dst[1], src[0], src[1]);
}

int adaflo_b200_ns_preconditioner_statistics(adaflo_b200_precond_t P, int *n_applications, int inner_iterations[3])
{
  if (!P) return fail(ADAFLO_B200_ERR_ARG, "null argument");
  if (n_applications) *n_applications = P->n_applications;
  if (inner_iterations)
    for (int i = 0; i < 3; ++i) inner_iterations[i] = P->inner_iterations[i];
  return 0;
}

}  // extern "C"
