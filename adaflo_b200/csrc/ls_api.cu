// ls_api.cu -- C ABI of the level-set operators (include/adaflo_b200.h, "level-set operators"):
// operator-level logic of LevelSetOKZSolver{AdvanceConcentration, ComputeNormal, ComputeCurvature,
// Reinitialization}::*_vmult / local_*_rhs around the sub-cell kernels of ls_ops.cu.
#include <algorithm>
#include <cmath>
#include <map>
#include <new>
#include <vector>

#include "ls_ops.cuh"

struct adaflo_b200_ls_s {
  adaflo_b200_ctx_t ctx;
  adaflo_b200_mf_t mf;
  adaflo_b200_ls_params prm;
  int dof_ls, dof_vel, quad_ls;
  double weight, weight_old, weight_old_old, step_size;
  double *cell_diameters;      // device
  double *conv, *normal_q;     // device [cell][3][8 points][sub-cell]
  const double *diag;          // not owned
  double *inv_diag;            // owned: DiagonalPreconditioner::inverse_diagonal_vector of diag (CG / BiCGStab preconditioner)
  uint32_t *last_cell;         // device [n_dofs]: last cell of the cell loop containing the dof (Heaviside)
  signed char *cell_state;     // device [n_cells]
  // hanging-dof extension of the level-set space (see LsHanging), all device pointers
  uint32_t n_hang;
  uint32_t *cell_ext, *hang_idx, *hang_n;
  double *c_hang, *heav_hang, *curv_hang;
  bool heav_hang_valid;
  double *interp;              // device [8][N] level set -> pressure interpolation
};

using namespace ab200;

namespace {
LSArgs base_args(adaflo_b200_ls_t ls)
{
  adaflo_b200_mf_t mf = ls->mf;
  const Space &sp = mf->spaces[ls->dof_ls];
  const Quad &q = mf->quads[ls->quad_ls];
  LSArgs a{};
  a.n_cells = mf->n_cells;
  a.n_sub = ls->prm.concentration_subdivisions;
  a.cell_ls = sp.cell_nodes;
  a.cell_u = ls->dof_vel >= 0 ? mf->spaces[ls->dof_vel].cell_nodes : nullptr;
  a.hn_u = ls->dof_vel >= 0 ? mf->spaces[ls->dof_vel].hanging_masks : nullptr;
  a.hn = sp.hanging_masks;
  a.geom_row = q.cell_row;
  a.geom_table = q.table;
  a.geom = q.geometry_type;
  a.inv_jac = q.inv_jac;
  a.jxw_q = q.jxw;
  a.cell_diameters = ls->cell_diameters;
  a.gamma = ls->weight;
  a.w_old = ls->weight_old;
  a.w_oo = ls->weight_old_old;
  a.eps_ratio = ls->prm.epsilon_used / ls->prm.epsilon;
  a.epsilon_used = ls->prm.epsilon_used;
  const int dim = ls->prm.dim;
  // reinitialization.cc:65-67
  a.dtau_inv = std::max(0.95 / (1. / (dim * dim) * ls->prm.minimal_edge_length / ls->prm.concentration_subdivisions), 1. / (5. * ls->step_size));
  a.n_comp = 1;
  return a;
}

int ensure_pointdata(adaflo_b200_ls_t ls, double **which)
{
  if (*which) return 0;
  const int S = ls->prm.concentration_subdivisions;
  const size_t npd = (size_t)ls->mf->n_cells * S * S * S * 24;
  AB_CUDA(cudaMalloc((void **)which, std::max<size_t>(1, npd) * sizeof(double)));
  AB_CUDA(cudaMemsetAsync(*which, 0, npd * sizeof(double), ls->ctx->stream));
  return 0;
}

int finish_vmult(adaflo_b200_ls_t ls, double *dst, const double *src)
{
  const Space &sp = ls->mf->spaces[ls->dof_ls];
  if (sp.n_constrained) {
    if (!ls->diag) return fail(ADAFLO_B200_ERR_STATE, "level-set vmult: constrained rows need the preconditioner diagonal (ls_set_diagonal)");
    AB_CHECK(launch_scaled_copy_indexed(dst, src, ls->diag, sp.constrained, sp.n_constrained, ls->ctx->stream));
  }
  return 0;
}

// MatrixFree::cell_loop semantics on a partitioned mesh (level_set_okz_*.cc call cell_loop for every operator): import the
// ghost values of the sources before the cells, add the ghost contributions of the destinations to their owners after.
int import_ghosts(adaflo_b200_ls_t ls, int space, int n, const double *const *vecs)
{
  if (ls->ctx->nranks <= 1) return 0;
  int spaces[4];
  double *v[4];
  int m = 0;
  for (int i = 0; i < n && m < 4; ++i)
    if (vecs[i]) {
      spaces[m] = space;
      v[m++] = const_cast<double *>(vecs[i]);  // the ghost section of a source is scratch (as in deal.II)
    }
  return m ? comm_update_ghosts_n(ls->mf, m, spaces, v) : 0;
}
int import_ghosts(adaflo_b200_ls_t ls, int space, const double *v)
{
  return import_ghosts(ls, space, 1, &v);
}
int compress(adaflo_b200_ls_t ls, int space, int n, double *const *vecs)
{
  if (ls->ctx->nranks <= 1) return 0;
  int spaces[4] = {space, space, space, space};
  return comm_compress_add_n(ls->mf, n, spaces, vecs);
}
int compress(adaflo_b200_ls_t ls, int space, double *v) { return compress(ls, space, 1, &v); }
// right-hand sides are ADDED into the caller's vector; its ghost section must not carry stale data into compress(add)
int zero_ghosts(adaflo_b200_ls_t ls, int space, double *v)
{
  const Space &sp = ls->mf->spaces[space];
  if (ls->ctx->nranks > 1 && sp.n_ghost_nodes)
    AB_CUDA(cudaMemsetAsync(v + sp.n_owned_entries(), 0, sizeof(double) * (size_t)sp.n_ghost_nodes * sp.n_comp, ls->ctx->stream));
  return 0;
}

int zero(adaflo_b200_ls_t ls, double *v)
{
  AB_CUDA(cudaMemsetAsync(v, 0, ls->mf->spaces[ls->dof_ls].n_local_entries() * sizeof(double), ls->ctx->stream));
  return 0;
}
}  // namespace

extern "C" {

int adaflo_b200_ls_create(adaflo_b200_ctx_t ctx, adaflo_b200_mf_t mf, const adaflo_b200_ls_params *prm, int dof_ls, int dof_vel,
                          int quad_ls, const double *cell_diameters_host, adaflo_b200_ls_t *out)
{
  if (!ctx || !mf || !prm || !out || !cell_diameters_host) return fail(ADAFLO_B200_ERR_ARG, "ls_create: null argument");
  if (dof_ls < 0 || dof_ls >= (int)mf->spaces.size() || quad_ls < 0 || quad_ls >= (int)mf->quads.size() || dof_vel >= (int)mf->spaces.size())
    return fail(ADAFLO_B200_ERR_ARG, "ls_create: dof/quad index out of range");
  const Space &sp = mf->spaces[dof_ls];
  const int S = prm->concentration_subdivisions;
  // dim == 2: a 2-D problem on its one-cell-thick extrusion (the spaces and kernels are three-dimensional either way;
  // dim enters the pseudo time step of the reinitialisation and the curvature correction)
  if ((prm->dim != 3 && prm->dim != 2) || S < 1 || S > 4 || sp.fe_kind != ADAFLO_B200_FE_Q_ISO_Q1 || sp.degree != S || sp.n_comp != 1 ||
      sp.nodes_per_cell != (uint32_t)((S + 1) * (S + 1) * (S + 1)) || mf->quads[quad_ls].n_q_1d != 2 * S)
    return fail(ADAFLO_B200_ERR_UNSUPPORTED, "ls_create: needs a 3-D FE_Q_iso_Q1(n_sub <= 4) space and the QIterated(QGauss(2), n_sub) quadrature");
  if (prm->convection_stabilization) return fail(ADAFLO_B200_ERR_UNSUPPORTED, "ls_create: convection_stabilization is not implemented");
  if (dof_vel >= 0 && (mf->spaces[dof_vel].nodes_per_cell != 27 || mf->spaces[dof_vel].n_comp != 3))
    return fail(ADAFLO_B200_ERR_UNSUPPORTED, "ls_create: the velocity space must be Q2^3");
  AB_CUDA(cudaSetDevice(ctx->device));
  adaflo_b200_ls_t ls = new (std::nothrow) adaflo_b200_ls_s();
  if (!ls) return fail(ADAFLO_B200_ERR_ARG, "out of host memory");
  ls->ctx = ctx;
  ls->mf = mf;
  ls->prm = *prm;
  ls->dof_ls = dof_ls;
  ls->dof_vel = dof_vel;
  ls->quad_ls = quad_ls;
  ls->weight = 1;
  ls->weight_old = -1;
  ls->weight_old_old = 0;
  ls->step_size = 1;
  ls->diag = nullptr;
  ls->inv_diag = nullptr;
  const size_t npd = (size_t)mf->n_cells * S * S * S * 24;
  AB_CUDA(cudaMalloc((void **)&ls->cell_diameters, std::max<size_t>(1, mf->n_cells) * sizeof(double)));
  AB_CUDA(cudaMemcpyAsync(ls->cell_diameters, cell_diameters_host, mf->n_cells * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  (void)npd;  // the per-point streams (24 B per point each) are allocated on first use: the four level-set classes of the
              // reference each own a handle (include/adaflo_b200/level_set_okz.h) and most never touch them
  ls->conv = nullptr;
  ls->normal_q = nullptr;
  {  // Heaviside bookkeeping: the reference loops over cells and the last writer of a shared dof wins
    const size_t N = sp.nodes_per_cell, n_dofs = sp.n_local_entries();
    std::vector<uint32_t> cells((size_t)mf->n_cells * N), masks;
    if (!cells.empty()) AB_CUDA(cudaMemcpy(cells.data(), sp.cell_nodes, cells.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    std::vector<uint32_t> ext, hang_idx, hang_n;
    ls->n_hang = 0;
    if (sp.hanging_masks && mf->n_cells) {
      // Hanging dofs are not stored in this library's vectors, but compute_heaviside / compute_force read them
      // plainly in the reference.  Identify them: a slot on a hanging line either coincides with one real dof or
      // is the mean of 2/4/8 parents; two slots are the same hanging dof iff they have the same parents.
      masks.resize(mf->n_cells);
      AB_CUDA(cudaMemcpy(masks.data(), sp.hanging_masks, masks.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost));
      ext.resize(cells.size());
      std::map<std::vector<uint32_t>, uint32_t> ids;
      const int n1 = S + 1;
      for (uint32_t c = 0; c < mf->n_cells; ++c) {
        const HnInfo hn(masks[c]);
        for (int k = 0; k < n1; ++k)
          for (int j = 0; j < n1; ++j)
            for (int i = 0; i < n1; ++i) {
              const size_t slot = c * N + i + n1 * (j + n1 * k);
              if (!hn.any()) {
                ext[slot] = cells[slot] & ADAFLO_B200_NODE_MASK;
                continue;
              }
              LsStencil st;
              ls_stencil(hn, cells.data() + c * N, S, i, j, k, st);
              if (st.n == 1) {
                ext[slot] = st.slot[0] & ADAFLO_B200_NODE_MASK;
                continue;
              }
              std::vector<uint32_t> key(st.n);
              for (int q = 0; q < st.n; ++q) key[q] = st.slot[q] & ADAFLO_B200_NODE_MASK;
              std::sort(key.begin(), key.end());
              auto it = ids.find(key);
              if (it == ids.end()) {
                it = ids.emplace(key, (uint32_t)hang_n.size()).first;
                hang_n.push_back((uint32_t)st.n);
                for (int q = 0; q < 8; ++q) hang_idx.push_back(q < st.n ? key[q] : 0u);
              }
              ext[slot] = (uint32_t)n_dofs + it->second;
            }
      }
      ls->n_hang = (uint32_t)hang_n.size();
    }
    const std::vector<uint32_t> &lists = ext.empty() ? cells : ext;
    std::vector<uint32_t> last(std::max<size_t>(1, n_dofs + ls->n_hang), 0u);
    for (uint32_t c = 0; c < mf->n_cells; ++c)
      for (size_t i = 0; i < N; ++i) last[ext.empty() ? (lists[c * N + i] & ADAFLO_B200_NODE_MASK) : lists[c * N + i]] = c;
    AB_CUDA(cudaMalloc((void **)&ls->last_cell, last.size() * sizeof(uint32_t)));
    AB_CUDA(cudaMemcpy(ls->last_cell, last.data(), last.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    AB_CUDA(cudaMalloc((void **)&ls->cell_state, std::max<size_t>(1, mf->n_cells)));
    AB_CUDA(cudaMalloc((void **)&ls->interp, 8 * N * sizeof(double)));
    if (!ext.empty()) {
      AB_CUDA(cudaMalloc((void **)&ls->cell_ext, ext.size() * sizeof(uint32_t)));
      AB_CUDA(cudaMemcpy(ls->cell_ext, ext.data(), ext.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
      const size_t nh = std::max<size_t>(1, ls->n_hang);
      AB_CUDA(cudaMalloc((void **)&ls->hang_idx, nh * 8 * sizeof(uint32_t)));
      AB_CUDA(cudaMalloc((void **)&ls->hang_n, nh * sizeof(uint32_t)));
      if (ls->n_hang) {
        AB_CUDA(cudaMemcpy(ls->hang_idx, hang_idx.data(), hang_idx.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
        AB_CUDA(cudaMemcpy(ls->hang_n, hang_n.data(), hang_n.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
      }
      AB_CUDA(cudaMalloc((void **)&ls->c_hang, nh * sizeof(double)));
      AB_CUDA(cudaMalloc((void **)&ls->heav_hang, nh * sizeof(double)));
      AB_CUDA(cudaMalloc((void **)&ls->curv_hang, nh * sizeof(double)));
    }
  }
  AB_CUDA(cudaStreamSynchronize(ctx->stream));
  *out = ls;
  return 0;
}

int adaflo_b200_ls_destroy(adaflo_b200_ls_t ls)
{
  if (!ls) return 0;
  cudaSetDevice(ls->ctx->device);
  cudaStreamSynchronize(ls->ctx->stream);
  cudaFree(ls->cell_diameters);
  cudaFree(ls->inv_diag);
  cudaFree(ls->conv);
  cudaFree(ls->normal_q);
  cudaFree(ls->last_cell);
  cudaFree(ls->cell_ext);
  cudaFree(ls->hang_idx);
  cudaFree(ls->hang_n);
  cudaFree(ls->c_hang);
  cudaFree(ls->heav_hang);
  cudaFree(ls->curv_hang);
  cudaFree(ls->cell_state);
  cudaFree(ls->interp);
  delete ls;
  return 0;
}

int adaflo_b200_ls_set_time(adaflo_b200_ls_t ls, double weight, double weight_old, double weight_old_old, double step_size)
{
  if (!ls) return fail(ADAFLO_B200_ERR_ARG, "null argument");
  ls->weight = weight;
  ls->weight_old = weight_old;
  ls->weight_old_old = weight_old_old;
  ls->step_size = step_size;
  return 0;
}

int adaflo_b200_ls_set_diagonal(adaflo_b200_ls_t ls, const double *diag)
{
  if (!ls) return fail(ADAFLO_B200_ERR_ARG, "null argument");
  ls->diag = diag;
  if (!diag) return 0;
  const uint64_t n = ls->mf->spaces[ls->dof_ls].n_owned_entries();
  if (!ls->inv_diag) AB_CUDA(cudaMalloc((void **)&ls->inv_diag, std::max<uint64_t>(1, n) * sizeof(double)));
  return adaflo_b200_jacobi_reinit(ls->ctx, ls->inv_diag, diag, n);  // diagonal_preconditioner.cc:38-60
}

int adaflo_b200_ls_mass_diagonal(adaflo_b200_ls_t ls, double *diag)
{
  if (!ls || !diag) return fail(ADAFLO_B200_ERR_ARG, "null argument");
  AB_CHECK(zero(ls, diag));
  LSArgs a = base_args(ls);
  a.dst[0] = diag;
  AB_CHECK(launch_ls(a, LS_MASS_DIAG, ls->ctx->stream));
  return compress(ls, ls->dof_ls, diag);
}

int adaflo_b200_ls_advance_concentration_vmult(adaflo_b200_ls_t ls, double *dst, const double *src)
{
  if (!ls || !dst || !src) return fail(ADAFLO_B200_ERR_ARG, "null argument");
  AB_CHECK(zero(ls, dst));
  AB_CHECK(import_ghosts(ls, ls->dof_ls, src));
  LSArgs a = base_args(ls);
  a.src[0] = src;
  a.dst[0] = dst;
  AB_CHECK(ensure_pointdata(ls, &ls->conv));
  a.pointdata = ls->conv;
  AB_CHECK(launch_ls(a, LS_ADVECT, ls->ctx->stream));
  AB_CHECK(compress(ls, ls->dof_ls, dst));
  return finish_vmult(ls, dst, src);
}

int adaflo_b200_ls_advance_concentration_rhs(adaflo_b200_ls_t ls, double *rhs, const double *vel, const double *c, const double *c_old,
                                             const double *c_old_old)
{
  if (!ls || !rhs || !vel || !c || !c_old || !c_old_old) return fail(ADAFLO_B200_ERR_ARG, "null argument");
  if (ls->dof_vel < 0) return fail(ADAFLO_B200_ERR_STATE, "advance_concentration_rhs: no velocity space given at create");
  LSArgs a = base_args(ls);
  a.src[0] = c;
  a.aux[0] = c_old;
  a.aux[1] = c_old_old;
  a.vel = vel;
  a.dst[0] = rhs;
  AB_CHECK(ensure_pointdata(ls, &ls->conv));
  a.pointdata = ls->conv;
  {
    const double *in[3] = {c, c_old, c_old_old};
    AB_CHECK(import_ghosts(ls, ls->dof_ls, 3, in));
    AB_CHECK(import_ghosts(ls, ls->dof_vel, vel));
    AB_CHECK(zero_ghosts(ls, ls->dof_ls, rhs));
  }
  AB_CHECK(launch_ls(a, LS_ADVECT_RHS, ls->ctx->stream));
  return compress(ls, ls->dof_ls, rhs);
}

int adaflo_b200_ls_compute_normal_vmult(adaflo_b200_ls_t ls, double *const dst[3], const double *const src[3])
{
  if (!ls || !dst || !src) return fail(ADAFLO_B200_ERR_ARG, "null argument");
  LSArgs a = base_args(ls);
  for (int d = 0; d < 3; ++d) {
    AB_CHECK(zero(ls, dst[d]));
    a.src[d] = src[d];
    a.dst[d] = dst[d];
  }
  a.n_comp = 3;
  a.delta_scale = ls->prm.damping_scale_factor;
  AB_CHECK(import_ghosts(ls, ls->dof_ls, 3, src));
  AB_CHECK(launch_ls(a, LS_DAMPED, ls->ctx->stream));
  AB_CHECK(compress(ls, ls->dof_ls, 3, dst));
  for (int d = 0; d < 3; ++d) AB_CHECK(finish_vmult(ls, dst[d], src[d]));
  return 0;
}

int adaflo_b200_ls_projection_vmult(adaflo_b200_ls_t ls, double *dst, const double *src)
{
  if (!ls || !dst || !src) return fail(ADAFLO_B200_ERR_ARG, "null argument");
  AB_CHECK(zero(ls, dst));
  LSArgs a = base_args(ls);
  a.src[0] = src;
  a.dst[0] = dst;
  a.n_comp = 1;
  a.delta_scale = ls->prm.damping_scale_factor;
  AB_CHECK(import_ghosts(ls, ls->dof_ls, src));
  AB_CHECK(launch_ls(a, LS_DAMPED, ls->ctx->stream));
  AB_CHECK(compress(ls, ls->dof_ls, dst));
  return finish_vmult(ls, dst, src);
}

int adaflo_b200_ls_compute_normal_rhs(adaflo_b200_ls_t ls, double *const rhs[3], const double *level_set)
{
  if (!ls || !rhs || !level_set) return fail(ADAFLO_B200_ERR_ARG, "null argument");
  LSArgs a = base_args(ls);
  a.src[0] = level_set;
  for (int d = 0; d < 3; ++d) a.dst[d] = rhs[d];
  AB_CHECK(import_ghosts(ls, ls->dof_ls, level_set));
  for (int d = 0; d < 3; ++d) AB_CHECK(zero_ghosts(ls, ls->dof_ls, rhs[d]));
  AB_CHECK(launch_ls(a, LS_NORMAL_RHS, ls->ctx->stream));
  return compress(ls, ls->dof_ls, 3, rhs);
}

int adaflo_b200_ls_compute_curvature_vmult(adaflo_b200_ls_t ls, double *dst, const double *src, int apply_diffusion)
{
  if (!ls || !dst || !src) return fail(ADAFLO_B200_ERR_ARG, "null argument");
  AB_CHECK(zero(ls, dst));
  LSArgs a = base_args(ls);
  a.src[0] = src;
  a.dst[0] = dst;
  a.delta_scale = apply_diffusion ? 1.0 : 0.0;
  AB_CHECK(import_ghosts(ls, ls->dof_ls, src));
  AB_CHECK(launch_ls(a, LS_DAMPED, ls->ctx->stream));
  AB_CHECK(compress(ls, ls->dof_ls, dst));
  return finish_vmult(ls, dst, src);
}

int adaflo_b200_ls_compute_curvature_rhs(adaflo_b200_ls_t ls, double *rhs, const double *const normal[3])
{
  if (!ls || !rhs || !normal) return fail(ADAFLO_B200_ERR_ARG, "null argument");
  LSArgs a = base_args(ls);
  for (int d = 0; d < 3; ++d) a.aux[d] = normal[d];
  a.dst[0] = rhs;
  AB_CHECK(import_ghosts(ls, ls->dof_ls, 3, normal));
  AB_CHECK(zero_ghosts(ls, ls->dof_ls, rhs));
  AB_CHECK(launch_ls(a, LS_CURV_RHS, ls->ctx->stream));
  return compress(ls, ls->dof_ls, rhs);
}

int adaflo_b200_ls_curvature_correction(adaflo_b200_ls_t ls, double *curvature, const double *level_set)
{
  if (!ls || !curvature || !level_set) return fail(ADAFLO_B200_ERR_ARG, "null argument");
  return launch_curvature_correction(curvature, level_set, ls->prm.epsilon_used, ls->prm.dim, ls->mf->spaces[ls->dof_ls].n_owned_entries(),
                                     ls->ctx->stream);
}

int adaflo_b200_ls_reinitialization_vmult(adaflo_b200_ls_t ls, double *dst, const double *src, int diffuse_only)
{
  if (!ls || !dst || !src) return fail(ADAFLO_B200_ERR_ARG, "null argument");
  AB_CHECK(zero(ls, dst));
  LSArgs a = base_args(ls);
  a.src[0] = src;
  a.dst[0] = dst;
  AB_CHECK(ensure_pointdata(ls, &ls->normal_q));
  a.pointdata = ls->normal_q;
  AB_CHECK(import_ghosts(ls, ls->dof_ls, src));
  AB_CHECK(launch_ls(a, diffuse_only ? LS_REINIT_DIFFUSE : LS_REINIT, ls->ctx->stream));
  AB_CHECK(compress(ls, ls->dof_ls, dst));
  return finish_vmult(ls, dst, src);
}

int adaflo_b200_ls_reinitialization_rhs(adaflo_b200_ls_t ls, double *rhs, const double *solution, const double *const normal[3],
                                        int diffuse_only, int first_reinit_step)
{
  if (!ls || !rhs || !solution) return fail(ADAFLO_B200_ERR_ARG, "null argument");
  if (!diffuse_only && first_reinit_step && !normal) return fail(ADAFLO_B200_ERR_ARG, "reinitialization_rhs: normal field required");
  LSArgs a = base_args(ls);
  a.src[0] = solution;
  if (normal)
    for (int d = 0; d < 3; ++d) a.aux[d] = normal[d];
  a.dst[0] = rhs;
  AB_CHECK(ensure_pointdata(ls, &ls->normal_q));
  a.pointdata = ls->normal_q;
  a.first_step = first_reinit_step;
  AB_CHECK(import_ghosts(ls, ls->dof_ls, solution));
  if (normal) AB_CHECK(import_ghosts(ls, ls->dof_ls, 3, normal));
  AB_CHECK(zero_ghosts(ls, ls->dof_ls, rhs));
  AB_CHECK(launch_ls(a, diffuse_only ? LS_REINIT_RHS_DIFFUSE : LS_REINIT_RHS, ls->ctx->stream));
  return compress(ls, ls->dof_ls, rhs);
}

static LsHanging hanging_of(const adaflo_b200_ls_t ls)
{
  LsHanging h{};
  h.n_real = (uint32_t)ls->mf->spaces[ls->dof_ls].n_local_entries();
  h.n_hang = ls->n_hang;
  h.cell_ext = ls->cell_ext;
  h.hang_idx = ls->hang_idx;
  h.hang_n = ls->hang_n;
  return h;
}

int adaflo_b200_ls_compute_heaviside(adaflo_b200_ls_t ls, double *heaviside, const double *level_set)
{
  if (!ls || !heaviside || !level_set) return fail(ADAFLO_B200_ERR_ARG, "null argument");
  const Space &sp = ls->mf->spaces[ls->dof_ls];
  // level_set_okz.cc:516-520: distance scaled by 2 epsilon / concentration_subdivisions
  const double scale = ls->prm.epsilon * 2. / ls->prm.concentration_subdivisions;
  if (ls->cell_ext && ls->ctx->nranks > 1)
    return fail(ADAFLO_B200_ERR_UNSUPPORTED, "compute_heaviside: hanging nodes on partitioned meshes are not implemented");
  ls->heav_hang_valid = true;
  AB_CHECK(import_ghosts(ls, ls->dof_ls, level_set));
  return launch_heaviside(sp.cell_nodes, ls->last_cell, ls->cell_state, level_set, heaviside, (int)sp.nodes_per_cell, ls->mf->n_cells,
                          sp.n_owned_entries(), scale, hanging_of(ls), ls->c_hang, ls->heav_hang, ls->ctx->stream);
}

int adaflo_b200_ls_compute_force(adaflo_b200_ls_t ls, adaflo_b200_ns_t ns, double *dst_u, const double *heaviside, const double *curvature,
                                 double surface_tension, double gravity, const double *interpolation_host)
{
  if (!ls || !ns || !dst_u || !heaviside || !curvature) return fail(ADAFLO_B200_ERR_ARG, "null argument");
  if (ns->mf != ls->mf) return fail(ADAFLO_B200_ERR_ARG, "compute_force: operator and level set must share the MatrixFree data");
  if (ls->cell_ext && !ls->heav_hang_valid)
    return fail(ADAFLO_B200_ERR_STATE, "compute_force: on meshes with hanging nodes the Heaviside values of the hanging dofs come from "
                                       "compute_heaviside, which has not been called");
  adaflo_b200_mf_t mf = ls->mf;
  const Space &sp = mf->spaces[ls->dof_ls], &su = mf->spaces[ns->dof_u];
  const Quad &q = mf->quads[ns->quad_u];
  const int S = ls->prm.concentration_subdivisions, n1 = S + 1;
  ForceArgs f{};
  f.ns.n_cells = mf->n_cells;
  f.ns.cell_u = su.cell_nodes;
  f.ns.hn_u = su.hanging_masks;
  f.ns.geom_row = q.cell_row;
  f.ns.geom_table = q.table;
  f.ns.inv_jac = q.inv_jac;
  f.ns.jxw = q.jxw;
  f.ns.dst_u = dst_u;
  f.cell_ls = sp.cell_nodes;
  f.n_sub = S;
  f.heaviside = heaviside;
  f.hang = hanging_of(ls);
  f.heaviside_hang = ls->heav_hang;
  f.curvature_hang = ls->curv_hang;
  if (ls->cell_ext) AB_CHECK(launch_hanging_mean(f.hang, curvature, ls->curv_hang, ls->ctx->stream));  // distributed curvature
  f.curvature = curvature;
  f.rho_q = ns->var ? ns->rho : nullptr;  // use_variable_parameters, level_set_okz.cc:333-334
  f.mu_q = ns->var ? ns->mu : nullptr;
  if (interpolation_host) {
    AB_CUDA(cudaMemcpyAsync(ls->interp, interpolation_host, 8 * (size_t)n1 * n1 * n1 * sizeof(double), cudaMemcpyHostToDevice, ls->ctx->stream));
    AB_CUDA(cudaStreamSynchronize(ls->ctx->stream));
    f.interp = ls->interp;
  }
  const double gp[3] = {0.11270166537925831, 0.5, 0.8872983346207417};
  for (int p = 0; p < 3; ++p) {
    for (int i = 0; i < 5; ++i) f.hat_v[p][i] = f.hat_d[p][i] = 0;
    int s = (int)(gp[p] * S);
    if (s > S - 1) s = S - 1;  // piecewise polynomials are evaluated on the interval [s/S, (s+1)/S)
    const double xi = gp[p] * S - s;
    f.hat_v[p][s] = 1 - xi;
    f.hat_v[p][s + 1] = xi;
    f.hat_d[p][s] = -(double)S;
    f.hat_d[p][s + 1] = (double)S;
  }
  f.surface_tension = surface_tension;
  f.gravity = gravity;
  f.gravity_dir = ls->prm.dim - 1;
  if (ls->cell_ext && ls->ctx->nranks > 1)
    return fail(ADAFLO_B200_ERR_UNSUPPORTED, "compute_force: hanging nodes on partitioned meshes are not implemented");
  {
    const double *in[2] = {heaviside, curvature};
    AB_CHECK(import_ghosts(ls, ls->dof_ls, 2, in));
    AB_CHECK(zero_ghosts(ls, ns->dof_u, dst_u));
  }
  f.density = ns->prm.density;
  f.density_diff = ns->prm.density_diff;
  f.viscosity = ns->prm.viscosity;
  f.viscosity_diff = ns->prm.viscosity_diff;
  AB_CHECK(launch_ls_force(f, q.geometry_type, ls->ctx->stream));
  return compress(ls, ns->dof_u, dst_u);
}

int adaflo_b200_ls_pointdata_ptr(adaflo_b200_ls_t ls, int which, double **dev_ptr, uint64_t *n)
{
  if (!ls || !dev_ptr) return fail(ADAFLO_B200_ERR_ARG, "null argument");
  const int S = ls->prm.concentration_subdivisions;
  AB_CHECK(ensure_pointdata(ls, which == 0 ? &ls->conv : &ls->normal_q));
  *dev_ptr = which == 0 ? ls->conv : ls->normal_q;
  if (n) *n = (uint64_t)ls->mf->n_cells * S * S * S * 24;
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// Solver-level routines of the level-set classes: right-hand side, Krylov solve with the diagonal
// preconditioner, update -- the bodies of compute_normal / compute_curvature / reinitialize / advance_concentration with
// the matrix-free operators in place of the assembled projection matrix + ILU (the matrix-free CG calls are still in
// the reference as comments: compute_normal.cc:258-260, compute_curvature.cc:331).
namespace {
struct Work {  // pool vector, released on scope exit
  adaflo_b200_ctx_t ctx;
  double *p;
  Work(adaflo_b200_ctx_t c, uint64_t n) : ctx(c), p(pool_get(c, n)) {}
  ~Work() { if (p) pool_put(ctx, p); }
  Work(const Work &) = delete;
  Work &operator=(const Work &) = delete;
};
int ls_zero(adaflo_b200_ls_t ls, double *v) { return zero(ls, v); }
int need_diag(adaflo_b200_ls_t ls, const char *who)
{
  if (!ls->diag || !ls->inv_diag) return fail(ADAFLO_B200_ERR_STATE, std::string(who) + ": the preconditioner diagonal has not been set (ls_set_diagonal)");
  return 0;
}
}  // namespace

int adaflo_b200_ls_compute_normal(adaflo_b200_ls_t ls, double *const normal[3], const double *level_set, int fast_computation, int *iterations)
{
  if (!ls || !normal || !level_set) return fail(ADAFLO_B200_ERR_ARG, "null argument");
  AB_CHECK(need_diag(ls, "compute_normal"));
  const Space &sp = ls->mf->spaces[ls->dof_ls];
  const uint64_t n_owned = sp.n_owned_entries(), n_local = sp.n_local_entries();
  Work r0(ls->ctx, n_local), r1(ls->ctx, n_local), r2(ls->ctx, n_local);
  if (!r0.p || !r1.p || !r2.p) return fail(ADAFLO_B200_ERR_CUDA, "compute_normal: out of device memory");
  double *rhs[3] = {r0.p, r1.p, r2.p};
  for (int d = 0; d < 3; ++d) AB_CHECK(ls_zero(ls, rhs[d]));
  AB_CHECK(adaflo_b200_ls_compute_normal_rhs(ls, rhs, level_set));
  // compute_normal.cc:252-261: ReductionControl(4000, 1e-50, fast ? 1e-5 : 1e-7), SolverCG on the block vector
  adaflo_b200_operator_ref A = {ADAFLO_B200_OP_LS_NORMAL, ls, 0, nullptr, 0, 3};
  adaflo_b200_operator_ref P = {ADAFLO_B200_OP_JACOBI, ls->ctx, 0, ls->inv_diag, n_owned, 3};
  adaflo_b200_solver_control c = {4000, 1e-50, fast_computation ? 1e-5 : 1e-7, 0, 0, 0, nullptr, 0, 0};
  const uint64_t no[3] = {n_owned, n_owned, n_owned}, nl[3] = {n_local, n_local, n_local};
  const int rc = adaflo_b200_solve_cg(ls->ctx, 3, no, nl, adaflo_b200_apply_operator, &A, adaflo_b200_apply_operator, &P, normal, rhs, &c);
  if (iterations) *iterations = c.last_step;
  AB_CHECK(rc);
  for (int d = 0; d < 3; ++d) AB_CHECK(import_ghosts(ls, ls->dof_ls, normal[d]));  // update_ghost_values, :285
  return 0;
}

int adaflo_b200_ls_compute_curvature(adaflo_b200_ls_t ls, double *curvature, const double *level_set, const double *const normal[3],
                                     int curvature_correction, int projection_operator, int *iterations)
{
  if (!ls || !curvature || !level_set || !normal) return fail(ADAFLO_B200_ERR_ARG, "null argument");
  AB_CHECK(need_diag(ls, "compute_curvature"));
  const Space &sp = ls->mf->spaces[ls->dof_ls];
  const uint64_t n_owned = sp.n_owned_entries(), n_local = sp.n_local_entries();
  Work r(ls->ctx, n_local);
  if (!r.p) return fail(ADAFLO_B200_ERR_CUDA, "compute_curvature: out of device memory");
  AB_CHECK(ls_zero(ls, r.p));
  AB_CHECK(adaflo_b200_ls_compute_curvature_rhs(ls, r.p, normal));
  // compute_curvature.cc:326-334: ReductionControl(2000, 1e-50, 1e-8).  projection_operator 0: compute_curvature_vmult with
  // diffusion (ComputeCurvatureMatrix); 1: the operator of the assembled projection_matrix the reference passes to the
  // solver (mass + damping_scale_factor delta^2 Laplace = one block of compute_normal_vmult)
  adaflo_b200_operator_ref A = {projection_operator ? ADAFLO_B200_OP_LS_NORMAL_SCALAR : ADAFLO_B200_OP_LS_CURVATURE, ls, 1, nullptr, 0, 1};
  adaflo_b200_operator_ref P = {ADAFLO_B200_OP_JACOBI, ls->ctx, 0, ls->inv_diag, n_owned, 1};
  adaflo_b200_solver_control c = {2000, 1e-50, 1e-8, 0, 0, 0, nullptr, 0, 0};
  double *x[1] = {curvature};
  const double *b[1] = {r.p};
  const int rc = adaflo_b200_solve_cg(ls->ctx, 1, &n_owned, &n_local, adaflo_b200_apply_operator, &A, adaflo_b200_apply_operator, &P, x, b, &c);
  if (iterations) *iterations = c.last_step;
  AB_CHECK(rc);
  if (curvature_correction) AB_CHECK(adaflo_b200_ls_curvature_correction(ls, curvature, level_set));  // :340-356
  return import_ghosts(ls, ls->dof_ls, curvature);
}

int adaflo_b200_ls_reinitialize(adaflo_b200_ls_t ls, double *solution, double *const normal[3], double dt, unsigned stab_steps, unsigned diff_steps,
                                const double *last_concentration_range, int *iterations, int *n_steps_done)
{
  if (!ls || !solution || !normal) return fail(ADAFLO_B200_ERR_ARG, "null argument");
  AB_CHECK(need_diag(ls, "reinitialize"));
  const Space &sp = ls->mf->spaces[ls->dof_ls];
  const uint64_t n_owned = sp.n_owned_entries(), n_local = sp.n_local_entries();
  Work rhs(ls->ctx, n_local), inc(ls->ctx, n_local);
  if (!rhs.p || !inc.p) return fail(ADAFLO_B200_ERR_CUDA, "reinitialize: out of device memory");
  const double step_size_before = ls->step_size;
  ls->step_size = dt;  // time_stepping.set_desired_time_step(dt), reinitialization.cc:263
  unsigned actual_diff_steps = diff_steps;
  if (last_concentration_range && (last_concentration_range[0] < -1.02 || last_concentration_range[1] > 1.02)) actual_diff_steps += 3;  // :281-283
  adaflo_b200_operator_ref P = {ADAFLO_B200_OP_JACOBI, ls->ctx, 0, ls->inv_diag, n_owned, 1};
  int done = 0, rc = 0;
  for (unsigned tau = 0; tau < actual_diff_steps + stab_steps; ++tau) {
    const bool diffuse = tau < actual_diff_steps, first_reinit_step = (tau == actual_diff_steps);
    if (first_reinit_step) {
      rc = adaflo_b200_ls_compute_normal(ls, normal, solution, 1, nullptr);  // compute_normal(true), :292-293
      if (rc) break;
    }
    if ((rc = ls_zero(ls, rhs.p))) break;
    if ((rc = adaflo_b200_ls_reinitialization_rhs(ls, rhs.p, solution, diffuse ? nullptr : normal, diffuse ? 1 : 0, first_reinit_step ? 1 : 0))) break;
    if ((rc = ls_zero(ls, inc.p))) break;
    adaflo_b200_operator_ref A = {ADAFLO_B200_OP_LS_REINIT, ls, diffuse ? 1 : 0, nullptr, 0, 1};
    adaflo_b200_solver_control c = {2000, 1e-50, 1e-6, 0, 0, 0, nullptr, 0, 0};  // :331-336
    double *x[1] = {inc.p};
    const double *b[1] = {rhs.p};
    rc = adaflo_b200_solve_cg(ls->ctx, 1, &n_owned, &n_local, adaflo_b200_apply_operator, &A, adaflo_b200_apply_operator, &P, x, b, &c);
    if (iterations) iterations[done] = c.last_step;
    ++done;
    if (rc) break;
    if ((rc = adaflo_b200_vec_axpy(ls->ctx, solution, 1.0, inc.p, n_owned))) break;  // solution += increment
    double nn = 0;
    if ((rc = adaflo_b200_vec_dot(ls->ctx, inc.p, inc.p, n_owned, &nn))) break;
    if (std::sqrt(nn) < 1e-6) break;  // :352-354
  }
  ls->step_size = step_size_before;
  if (n_steps_done) *n_steps_done = done;
  AB_CHECK(rc);
  return import_ghosts(ls, ls->dof_ls, solution);
}

int adaflo_b200_ls_advance_concentration(adaflo_b200_ls_t ls, double *solution, const double *solution_old, const double *solution_old_old,
                                         const double *vel_solution, double tol_nl_iteration, double *initial_residual, int *iterations)
{
  if (!ls || !solution || !solution_old || !solution_old_old || !vel_solution) return fail(ADAFLO_B200_ERR_ARG, "null argument");
  AB_CHECK(need_diag(ls, "advance_concentration"));
  const Space &sp = ls->mf->spaces[ls->dof_ls];
  const uint64_t n_owned = sp.n_owned_entries(), n_local = sp.n_local_entries();
  Work rhs(ls->ctx, n_local), inc(ls->ctx, n_local);
  if (!rhs.p || !inc.p) return fail(ADAFLO_B200_ERR_CUDA, "advance_concentration: out of device memory");
  AB_CHECK(ls_zero(ls, rhs.p));
  AB_CHECK(adaflo_b200_ls_advance_concentration_rhs(ls, rhs.p, vel_solution, solution, solution_old, solution_old_old));
  AB_CHECK(ls_zero(ls, inc.p));
  adaflo_b200_operator_ref A = {ADAFLO_B200_OP_LS_ADVECT, ls, 0, nullptr, 0, 1};
  adaflo_b200_operator_ref P = {ADAFLO_B200_OP_JACOBI, ls->ctx, 0, ls->inv_diag, n_owned, 1};
  double *x[1] = {inc.p};
  const double *b[1] = {rhs.p};
  // advance_concentration.cc:626-652: BiCGStab(30, 0.05 tol_nl_iteration, 1e-8), GMRES(3000) if it does not converge
  adaflo_b200_solver_control c = {30, 0.05 * tol_nl_iteration, 1e-8, 0, 0, 0, nullptr, 0, 0};
  int rc = adaflo_b200_solve_bicgstab(ls->ctx, 1, &n_owned, &n_local, adaflo_b200_apply_operator, &A, adaflo_b200_apply_operator, &P, x, b, &c);
  int n_it = c.last_step;
  const double r0 = c.initial_value;
  if (rc == ADAFLO_B200_ERR_NO_CONVERGENCE) {
    adaflo_b200_solver_control g = {3000, 0.05 * tol_nl_iteration, 1e-8, 0, 0, 0, nullptr, 0, 0};
    rc = adaflo_b200_solve_gmres(ls->ctx, 1, &n_owned, &n_local, adaflo_b200_apply_operator, &A, adaflo_b200_apply_operator, &P, x, b, 28, &g);
    n_it = 30 + g.last_step;
  }
  if (initial_residual) *initial_residual = r0;
  if (iterations) *iterations = n_it;
  AB_CHECK(rc);
  AB_CHECK(adaflo_b200_vec_axpy(ls->ctx, solution, 1.0, inc.p, n_owned));
  return import_ghosts(ls, ls->dof_ls, solution);
}

}  // extern "C"
