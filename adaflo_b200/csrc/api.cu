// api.cu -- the C ABI of libadaflo_b200.so (include/adaflo_b200.h): handle
// management, uploads, and the operator-level logic of NavierStokesMatrix
// (source/navier_stokes_matrix.cc:85-483) around the cell kernels.
#include <chrono>
#include <cstdlib>
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <new>

#include "common.cuh"

namespace ab200 {

static thread_local std::string g_error;
unsigned long long g_launches = 0;
void set_error(const std::string &msg) { g_error = msg; }
int fail(int code, const std::string &msg)
{
  g_error = msg;
  return code;
}

template <typename T>
static int upload(T **dev, const T *host, size_t n, cudaStream_t s)
{
  *dev = nullptr;
  if (n == 0) return 0;
  AB_CUDA(cudaMalloc((void **)dev, n * sizeof(T)));
  if (host) AB_CUDA(cudaMemcpyAsync(*dev, host, n * sizeof(T), cudaMemcpyHostToDevice, s));
  return 0;
}

static void free_dev(void *p)
{
  if (p) cudaFree(p);
}

static int mode_of(const adaflo_b200_flow_params &p)
{
  if (p.physical_type == ADAFLO_B200_STOKES) return MODE_STOKES;
  switch (p.linearization) {
    case ADAFLO_B200_COUPLED_IMPLICIT_NEWTON: return MODE_NEWTON;
    case ADAFLO_B200_COUPLED_VELOCITY_EXPLICIT: return MODE_NOCONV;
    default: return MODE_PICARD;
  }
}

static void fill_geometry(NSArgs &a, const Quad &q)
{
  a.geom_row = q.cell_row;
  a.geom_table = q.table;
  a.inv_jac = q.inv_jac;
  a.jxw = q.jxw;
}

static void fill_geometry(PArgs &a, const Quad &q)
{
  a.geom_row = q.cell_row;
  a.geom_table = q.table;
  a.inv_jac = q.inv_jac;
  a.jxw = q.jxw;
  a.geom = q.geometry_type;
}

static NSArgs make_args(adaflo_b200_ns_t ns, bool use_prec)
{
  adaflo_b200_mf_t mf = ns->mf;
  NSArgs a{};
  a.n_cells = mf->n_cells;
  a.work_counter = ns->ctx->work_counter;
  a.np_cell = (int)mf->spaces[ns->dof_p].nodes_per_cell;
  a.cell_u = mf->spaces[ns->dof_u].cell_nodes;
  a.cell_p = mf->spaces[ns->dof_p].cell_nodes;
  a.cell_u_blk = mf->spaces[ns->dof_u].cell_nodes_blk;
  a.cell_p_blk = mf->spaces[ns->dof_p].cell_nodes_blk;
  a.hn_u = mf->spaces[ns->dof_u].hanging_masks;
  a.hn_p = mf->spaces[ns->dof_p].hanging_masks;
  {
    const Space &su = mf->spaces[ns->dof_u], &sp = mf->spaces[ns->dof_p];
    a.gc_first_u = su.hn_first;
    a.gc_ptr_u = su.hn_row_ptr;
    a.gc_cols_u = su.hn_cols;
    a.gc_w_u = su.hn_weights;
    a.gc_first_p = sp.hn_first;
    a.gc_ptr_p = sp.hn_row_ptr;
    a.gc_cols_p = sp.hn_cols;
    a.gc_w_p = sp.hn_weights;
  }
  const bool prec = use_prec && ns->lin_prec;
  a.lin = prec ? ns->lin_prec : ns->lin;
  a.rho = (use_prec && ns->rho_prec) ? ns->rho_prec : ns->rho;
  a.mu = (use_prec && ns->mu_prec) ? ns->mu_prec : ns->mu;
  a.damp = (use_prec && ns->damp_prec) ? ns->damp_prec : ns->damp;
  fill_geometry(a, mf->quads[ns->quad_u]);
  const auto &p = ns->prm;
  const auto &t = ns->tw;
  a.gamma = p.physical_type == ADAFLO_B200_INCOMPRESSIBLE ? t.weight : 0.0;  // :621-623
  a.w_old = t.weight_old;
  a.w_oo = t.weight_old_old;
  a.tau1 = t.tau1;
  a.beta = p.beta_convective_term_momentum_balance;
  a.rho0 = p.density;
  a.mu0 = p.viscosity;
  a.damp0 = p.damping;
  a.tau_gd = p.tau_grad_div;
  a.ex_old = t.extrapolate_old;
  a.ex_oo = t.extrapolate_old_old;
  a.linearization = p.linearization;
  a.physical_type = p.physical_type;
  a.skip_p_test = p.linearization == ADAFLO_B200_PROJECTION;
  return a;
}

static int launch_cells(adaflo_b200_ns_t ns, const NSArgs &a, bool with_p)
{
  const Quad &q = ns->mf->quads[ns->quad_u];
  const int mode = mode_of(ns->prm);
  cudaStream_t s = ns->ctx->stream;
  if (ns->generic) {
    if (a.transpose) return fail(ADAFLO_B200_ERR_UNSUPPORTED, "Tvmult: only the 3-D Q2/Q1 kernels implement the transpose");
    return launch_ns_generic(a, ns->gen_dim, ns->gen_degree_u, 0, with_p ? 1 : 0, mode, q.geometry_type, s);
  }
  switch (q.geometry_type) {
    case ADAFLO_B200_GEOM_CARTESIAN: return launch_ns_vmult_cartesian(a, with_p, mode, ns->var, s);
    case ADAFLO_B200_GEOM_AFFINE: return launch_ns_vmult_affine(a, with_p, mode, ns->var, s);
    default: return launch_ns_vmult_general(a, with_p, mode, ns->var, s);
  }
}

static bool projection_active(const adaflo_b200_ns_t ns)
{
  // navier_stokes_matrix.cc:196-197
  return ns->prm.linearization != ADAFLO_B200_PROJECTION && ns->prm.physical_type != ADAFLO_B200_INCOMPRESSIBLE_STATIONARY;
}

}  // namespace ab200

// canonical <-> decade-blocked copies of a coefficient array.  Canonical is the reference's own storage order,
// AlignedVector<VectorizedArray<double>> indexed [cell batch][q] (navier_stokes_matrix.h:293-395; the linearisation point
// is a pair<Tensor<1,dim,VA>, Tensor<2,dim,VA>> per point, i.e. 12 consecutive VectorizedArrays):
//   canonical[((batch * 27 + q) * n_rows + row) * lanes + lane],  q = qx + 3 qy + 9 qz,
// slot batch * lanes + lane holds cell slot_cell[...] (identity if null; 0xffffffff marks an empty lane).
namespace ab200 {
template <bool TO_CANONICAL>
__global__ void k_coef_canonical(double *blocked, double *canonical, const uint32_t *slot_cell, uint32_t n_slots, uint32_t lanes,
                                 uint32_t n_cells, int n_rows)
{
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // one thread per (slot, q)
  if (i >= (size_t)n_slots * 27) return;
  const uint32_t slot = (uint32_t)(i / 27);
  const int q = (int)(i % 27), qx = q % 3, qy = (q / 3) % 3, qz = q / 9;
  const uint32_t cell = slot_cell ? slot_cell[slot] : slot;
  if (cell >= n_cells) return;
  const uint32_t batch = slot / lanes, lane = slot % lanes;
  double *can = canonical + (((size_t)batch * 27 + q) * n_rows) * lanes + lane;
  for (int r = 0; r < n_rows; ++r) {
    double *b = blocked + (n_rows == 12 ? lin_index(cell, qx, qy, qz) + lin_row(r) : coef_index(cell, r, n_rows, qx, qy, qz));
    if (TO_CANONICAL)
      can[(size_t)r * lanes] = *b;
    else
      *b = can[(size_t)r * lanes];
  }
}
}  // namespace ab200

using namespace ab200;

extern "C" {

const char *adaflo_b200_last_error(void) { return g_error.c_str(); }
const char *adaflo_b200_version(void) { return "adaflo_b200 0.1 (sm_100a)"; }

// ------------------------------------------------------------------ context
int adaflo_b200_ctx_create(adaflo_b200_ctx_t *out, int device, void *stream)
{
  if (!out) return fail(ADAFLO_B200_ERR_ARG, "ctx_create: null output");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(ADAFLO_B200_ERR_CUDA, std::string("no CUDA device available (there is no CPU fallback): ") + cudaGetErrorString(e));
  if (device < 0 || device >= count) return fail(ADAFLO_B200_ERR_ARG, "ctx_create: bad device ordinal");
  AB_CUDA(cudaSetDevice(device));
  adaflo_b200_ctx_t c = new (std::nothrow) adaflo_b200_ctx_s();
  if (!c) return fail(ADAFLO_B200_ERR_ARG, "out of host memory");
  c->device = device;
  c->stream = (cudaStream_t)stream;
  c->comm_stream = nullptr;
  c->nccl_comm = nullptr;
  c->rank = 0;
  c->nranks = 1;
  c->ev_watch = nullptr;
  c->comm_timeout_s = 0.0;  // single rank: plain synchronisation; comm_init arms the watchdog (120 s)
  if (const char *t = std::getenv("ADAFLO_B200_COMM_TIMEOUT")) c->comm_timeout_s = std::atof(t);  // seconds; also arms it without a communicator
  AB_CUDA(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device));
  c->red_capacity = 64 + (size_t)8 * c->sm_count * 8 + 64;
  AB_CUDA(cudaMalloc((void **)&c->red_dev, c->red_capacity * sizeof(double)));
  AB_CUDA(cudaMallocHost((void **)&c->red_host, 64 * sizeof(double)));
  c->work_counter = nullptr;
  AB_CUDA(cudaMalloc(&c->work_counter, 2 * sizeof(unsigned)));
  AB_CUDA(cudaMemset(c->work_counter, 0, 2 * sizeof(unsigned)));
  AB_CUDA(cudaEventCreate(&c->ev_a));
  AB_CUDA(cudaEventCreate(&c->ev_b));
  *out = c;
  return 0;
}

int adaflo_b200_ctx_destroy(adaflo_b200_ctx_t ctx)
{
  if (!ctx) return 0;
  cudaSetDevice(ctx->device);
  solver_pool_release(ctx);
  comm_destroy(ctx);
  free_dev(ctx->red_dev);
  if (ctx->red_host) cudaFreeHost(ctx->red_host);
  cudaFree(ctx->work_counter);
  cudaEventDestroy(ctx->ev_a);
  cudaEventDestroy(ctx->ev_b);
  if (ctx->ev_watch) cudaEventDestroy(ctx->ev_watch);
  if (ctx->comm_stream) cudaStreamDestroy(ctx->comm_stream);
  delete ctx;
  return 0;
}

int adaflo_b200_ctx_set_stream(adaflo_b200_ctx_t ctx, void *stream)
{
  if (!ctx) return fail(ADAFLO_B200_ERR_ARG, "null context");
  ctx->stream = (cudaStream_t)stream;
  return 0;
}

int adaflo_b200_ctx_synchronize(adaflo_b200_ctx_t ctx)
{
  if (!ctx) return fail(ADAFLO_B200_ERR_ARG, "null context");
  AB_CUDA(cudaSetDevice(ctx->device));
  return stream_wait(ctx, ctx->stream);
}

int adaflo_b200_comm_unique_id(unsigned char id[128]) { return comm_unique_id(id); }

int adaflo_b200_ctx_init_comm(adaflo_b200_ctx_t ctx, const unsigned char id[128], int rank, int nranks)
{
  if (!ctx) return fail(ADAFLO_B200_ERR_ARG, "null context");
  if (nranks <= 1) {
    ctx->rank = 0;
    ctx->nranks = 1;
    return 0;
  }
  AB_CHECK(comm_init(ctx, id, rank, nranks));
  if (!ctx->comm_stream) AB_CUDA(cudaStreamCreateWithFlags(&ctx->comm_stream, cudaStreamNonBlocking));
  return 0;
}

// ------------------------------------------------------------- matrix-free
int adaflo_b200_mf_create(adaflo_b200_ctx_t ctx, const adaflo_b200_mf_desc *d, adaflo_b200_mf_t *out)
{
  if (!ctx || !d || !out) return fail(ADAFLO_B200_ERR_ARG, "mf_create: null argument");
  if (d->dim != 3 && d->dim != 2) return fail(ADAFLO_B200_ERR_UNSUPPORTED, "mf_create: dim must be 2 or 3");
  AB_CUDA(cudaSetDevice(ctx->device));
  adaflo_b200_mf_t mf = new (std::nothrow) adaflo_b200_mf_s();
  if (!mf) return fail(ADAFLO_B200_ERR_ARG, "out of host memory");
  mf->ctx = ctx;
  mf->dim = d->dim;
  mf->n_cells = d->n_cells;
  cudaStream_t s = ctx->stream;
  for (int i = 0; i < d->n_spaces; ++i) {
    const adaflo_b200_space_desc &sd = d->spaces[i];
    Space sp{};
    sp.fe_kind = sd.fe_kind;
    sp.degree = sd.degree;
    sp.n_comp = sd.n_components;
    sp.n_owned_nodes = sd.n_owned_nodes;
    sp.n_ghost_nodes = sd.n_ghost_nodes;
    sp.nodes_per_cell = sd.nodes_per_cell;
    if (sd.n_components < 1 || sd.n_components > 3) return fail(ADAFLO_B200_ERR_ARG, "mf_create: n_components must be 1..3");
    if ((uint64_t)sd.n_owned_nodes + sd.n_ghost_nodes > ADAFLO_B200_NODE_MASK)
      return fail(ADAFLO_B200_ERR_ARG, "mf_create: more than 2^29 nodes per rank");
    AB_CHECK(upload(&sp.cell_nodes, sd.cell_nodes, (size_t)d->n_cells * sd.nodes_per_cell, s));
    if (sd.hanging_masks) AB_CHECK(upload(&sp.hanging_masks, sd.hanging_masks, (size_t)d->n_cells, s));
    if (d->dim == 3 && (sd.nodes_per_cell == 27 || sd.nodes_per_cell == 8 || (sd.nodes_per_cell == 9 && sd.fe_kind == ADAFLO_B200_FE_Q_DG0))) {
      // decade-blocked copy for the vmult kernel: lane 3 cw + iz reads in-plane node j of its plane
      const int n1 = sd.nodes_per_cell == 27 ? 3 : 2, nj = n1 * n1;
      const size_t ng = n_groups(d->n_cells);
      std::vector<uint32_t> blk(ng * nj * 30, 0u);
      for (uint32_t c = 0; c < d->n_cells; ++c)
        for (int iz = 0; iz < n1; ++iz)
          for (int j = 0; j < nj; ++j)
            blk[((size_t)(c / kGroup) * nj + j) * 30 + 3 * (c % kGroup) + iz] =
              sd.cell_nodes[(size_t)c * sd.nodes_per_cell + kernel_node(j % n1, j / n1, iz, n1)];
      AB_CHECK(upload(&sp.cell_nodes_blk, blk.data(), blk.size(), s));
      AB_CUDA(cudaStreamSynchronize(s));
    }
    sp.n_constrained = sd.n_constrained;
    AB_CHECK(upload(&sp.constrained, sd.constrained_dofs, (size_t)sd.n_constrained, s));
    if (sd.n_hanging_nodes) {  // general constraint rows (generic-degree path)
      if (!sd.hanging_row_ptr || !sd.hanging_cols || !sd.hanging_weights || sd.n_hanging_nodes > sd.n_owned_nodes)
        return fail(ADAFLO_B200_ERR_ARG, "mf_create: n_hanging_nodes needs hanging_row_ptr / hanging_cols / hanging_weights");
      const uint32_t first = sd.n_owned_nodes - sd.n_hanging_nodes, nnz = sd.hanging_row_ptr[sd.n_hanging_nodes];
      for (uint32_t k = 0; k < nnz; ++k)  // rows are resolved: parents are unconstrained nodes (AffineConstraints::close)
        if ((sd.hanging_cols[k] & ADAFLO_B200_NODE_MASK) >= first && (sd.hanging_cols[k] & ADAFLO_B200_NODE_MASK) < sd.n_owned_nodes)
          return fail(ADAFLO_B200_ERR_ARG, "mf_create: a constraint row refers to another constrained node (rows must be closed)");
      sp.hn_first = first;
      AB_CHECK(upload(&sp.hn_row_ptr, sd.hanging_row_ptr, (size_t)sd.n_hanging_nodes + 1, s));
      AB_CHECK(upload(&sp.hn_cols, sd.hanging_cols, (size_t)nnz, s));
      AB_CHECK(upload(&sp.hn_weights, sd.hanging_weights, (size_t)nnz, s));
    }
    if (sd.fe_kind == ADAFLO_B200_FE_Q_DG0) {
      sp.dg0_dofs.resize(d->n_cells);
      for (uint32_t c = 0; c < d->n_cells; ++c)
        sp.dg0_dofs[c] = sd.cell_nodes[(size_t)c * sd.nodes_per_cell + sd.nodes_per_cell - 1] & ADAFLO_B200_NODE_MASK;
    }
    for (int e = 0; e < sd.n_exchange; ++e) {
      const adaflo_b200_exchange &xd = sd.exchange[e];
      Exchange x{};
      x.peer = xd.peer;
      x.n_send = xd.n_send;
      x.recv_first = xd.recv_first;
      x.n_recv = xd.n_recv;
      AB_CHECK(upload(&x.send_dev, xd.send, (size_t)xd.n_send, s));
      if (xd.n_send) {
        AB_CUDA(cudaMalloc((void **)&x.send_buf, sizeof(double) * (size_t)xd.n_send * sd.n_components));
        x.send_slots = 1;
      }
      sp.exchange.push_back(x);
    }
    mf->spaces.push_back(sp);
  }
  for (int i = 0; i < d->n_quads; ++i) {
    const adaflo_b200_quad_desc &qd = d->quads[i];
    Quad q{};
    q.n_q_1d = qd.n_q_1d;
    q.geometry_type = qd.geometry_type;
    q.n_rows = qd.n_table_rows;
    const size_t nq = (size_t)qd.n_q_1d * qd.n_q_1d * (d->dim == 3 ? qd.n_q_1d : 1);
    if (qd.geometry_type == ADAFLO_B200_GEOM_GENERAL) {
      if (!qd.inv_jacobian || !qd.jxw) return fail(ADAFLO_B200_ERR_ARG, "mf_create: general geometry needs inv_jacobian and jxw");
      if (qd.n_q_1d == 3 && d->dim == 3) {  // decade blocked (see coef_index)
        const size_t ng = n_groups(d->n_cells);
        std::vector<double> kb(ng * 9 * 270, 0.0), jb(ng * 270, 1.0);
        for (uint32_t c = 0; c < d->n_cells; ++c)
          for (int qz = 0; qz < 3; ++qz)
            for (int qy = 0; qy < 3; ++qy)
              for (int qx = 0; qx < 3; ++qx) {
                const int qq = qx + 3 * qy + 9 * qz;
                for (int r = 0; r < 9; ++r) kb[coef_index(c, r, 9, qx, qy, qz)] = qd.inv_jacobian[((size_t)c * 9 + r) * 27 + qq];
                jb[coef_index(c, 0, 1, qx, qy, qz)] = qd.jxw[(size_t)c * 27 + qq];
              }
        AB_CHECK(upload(&q.inv_jac, kb.data(), kb.size(), s));
        AB_CHECK(upload(&q.jxw, jb.data(), jb.size(), s));
        AB_CUDA(cudaStreamSynchronize(s));
      } else {
        AB_CHECK(upload(&q.inv_jac, qd.inv_jacobian, (size_t)d->n_cells * 9 * nq, s));
        AB_CHECK(upload(&q.jxw, qd.jxw, (size_t)d->n_cells * nq, s));
      }
    } else {
      const int w = qd.geometry_type == ADAFLO_B200_GEOM_CARTESIAN ? 4 : 10;
      if (!qd.table || !qd.cell_row) return fail(ADAFLO_B200_ERR_ARG, "mf_create: compressed geometry needs table and cell_row");
      AB_CHECK(upload(&q.table, qd.table, (size_t)qd.n_table_rows * w, s));
      AB_CHECK(upload(&q.cell_row, qd.cell_row, (size_t)d->n_cells, s));
    }
    mf->quads.push_back(q);
  }
  {  // decades whose cells touch a ghost node (they must wait for the halo import)
    size_t lo = SIZE_MAX, hi = 0, count = 0;
    for (int i = 0; i < d->n_spaces; ++i) {
      const adaflo_b200_space_desc &sd = d->spaces[i];
      if (!sd.n_ghost_nodes) continue;
      for (uint32_t c = 0; c < d->n_cells; ++c)
        for (uint32_t k = 0; k < sd.nodes_per_cell; ++k)
          if ((sd.cell_nodes[(size_t)c * sd.nodes_per_cell + k] & ADAFLO_B200_NODE_MASK) >= sd.n_owned_nodes) {
            const size_t g = c / kGroup;
            if (g < lo) lo = g;
            if (g + 1 > hi) hi = g + 1;
            ++count;
            break;
          }
    }
    if (count) {
      mf->ghost_decade_begin = lo;
      mf->ghost_decade_end = hi;
      // worthwhile only if the range is a small part of the mesh
      mf->ghost_range_valid = (hi - lo) * 4 < n_groups(d->n_cells);
    } else
      mf->ghost_range_valid = true;  // no ghosts (lowest rank): empty range, the exports still overlap
  }
  {  // host-vector pipeline metadata (adaflo_b200_ns_vmult_host): monotone node prefixes per cell chunk
    const size_t n_dec = n_groups(d->n_cells);
    // 16 chunks: measured optimum at 128^3 cells (4: 11.6 ms, 8: 11.0, 16: 10.6, 32: 11.8, 64: 13.9 ms per vmult_host;
    // a chunk costs about 70 us of stream hand-over).  Graded chunks (small first and last ones, to shorten the
    // pipeline fill and drain) measured 11.0 vs 10.9 ms: the remainder over max(H2D, D2H) = 8.5 ms is the
    // bidirectional PCIe rate, not fill/drain.
    size_t n_chunks = n_dec < 16 ? (n_dec ? 1 : 0) : 16;
    const char *e_chunks = std::getenv("ADAFLO_B200_HOST_CHUNKS");  // tuning knob
    if (e_chunks && *e_chunks && n_dec >= 16) n_chunks = std::min<size_t>(n_dec, std::max<long>(1, std::atol(e_chunks)));
    mf->pipeline_ok = n_chunks > 0;
    for (size_t k = 0; k < n_chunks; ++k) mf->chunk_decade_end.push_back(k + 1 == n_chunks ? n_dec : (k + 1) * (n_dec / n_chunks));
    for (int i = 0; i < d->n_spaces && mf->pipeline_ok; ++i) {
      const adaflo_b200_space_desc &sd = d->spaces[i];
      Space &sp = mf->spaces[i];
      std::vector<uint32_t> cmax(n_chunks, 0), cmin(n_chunks, UINT32_MAX);
      size_t k = 0;
      for (uint32_t c = 0; c < d->n_cells; ++c) {
        while (c / kGroup >= mf->chunk_decade_end[k]) ++k;
        for (uint32_t j = 0; j < sd.nodes_per_cell; ++j) {
          const uint32_t n = sd.cell_nodes[(size_t)c * sd.nodes_per_cell + j] & ADAFLO_B200_NODE_MASK;
          cmax[k] = std::max(cmax[k], n + 1);
          cmin[k] = std::min(cmin[k], n);
        }
      }
      const uint32_t n_nodes = sd.n_owned_nodes + sd.n_ghost_nodes;
      sp.chunk_need_hi.assign(n_chunks, 0);
      sp.chunk_final_lo.assign(n_chunks, n_nodes);
      uint32_t run = 0;
      for (size_t q = 0; q < n_chunks; ++q) sp.chunk_need_hi[q] = run = std::max(run, cmax[q]);
      run = n_nodes;
      for (size_t q = n_chunks; q-- > 0;) {
        sp.chunk_final_lo[q] = run;  // nodes below the minimum of all later chunks are final after chunk q
        run = std::min(run, cmin[q]);
      }
      // nodes exported to neighbours (multi-GPU): uploaded first so that the ghost import can start at once
      sp.send_lo = n_nodes;
      sp.send_hi = 0;
      for (int e = 0; e < sd.n_exchange; ++e)
        for (uint32_t q = 0; q < sd.exchange[e].n_send; ++q) {
          sp.send_lo = std::min(sp.send_lo, sd.exchange[e].send[q]);
          sp.send_hi = std::max(sp.send_hi, sd.exchange[e].send[q] + 1);
        }
      if (sp.send_hi <= sp.send_lo) sp.send_lo = sp.send_hi = n_nodes;
      // constrained entries must be sorted to be applied prefix by prefix
      sp.chunk_constrained_end.assign(n_chunks, sd.n_constrained);
      for (uint32_t q = 1; q < sd.n_constrained; ++q)
        if (sd.constrained_dofs[q] < sd.constrained_dofs[q - 1]) mf->pipeline_ok = false;
      if (mf->pipeline_ok)
        for (size_t q = 0; q < n_chunks; ++q)
          sp.chunk_constrained_end[q] = (uint32_t)(std::lower_bound(sd.constrained_dofs, sd.constrained_dofs + sd.n_constrained,
                                                                    (uint32_t)std::min<uint64_t>((uint64_t)sp.chunk_final_lo[q] * sd.n_components, UINT32_MAX)) -
                                                   sd.constrained_dofs);
    }
  }
  AB_CUDA(cudaStreamSynchronize(s));  // host arrays may be freed by the caller now
  *out = mf;
  return 0;
}

int adaflo_b200_mf_destroy(adaflo_b200_mf_t mf)
{
  if (!mf) return 0;
  cudaSetDevice(mf->ctx->device);
  for (Space &sp : mf->spaces) {
    free_dev(sp.cell_nodes);
    free_dev(sp.cell_nodes_blk);
    free_dev(sp.hanging_masks);
    free_dev(sp.constrained);
    free_dev(sp.hn_row_ptr);
    free_dev(sp.hn_cols);
    free_dev(sp.hn_weights);
    free_dev(sp.numbering);
    free_dev(sp.numbering_stage);
    for (Exchange &x : sp.exchange) {
      free_dev(x.send_dev);
      free_dev(x.send_buf);
    }
  }
  for (Quad &q : mf->quads) {
    free_dev(q.table);
    free_dev(q.cell_row);
    free_dev(q.inv_jac);
    free_dev(q.jxw);
  }
  delete mf;
  return 0;
}

// ---- dof numbering of the host application (node-major by default)
extern "C++" {
namespace {
__global__ void k_from_host_order(double *__restrict__ dst, const double *__restrict__ src, const uint32_t *__restrict__ pos, uint64_t n)
{
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) dst[i] = src[pos[i]];
}
__global__ void k_to_host_order(double *__restrict__ dst, const double *__restrict__ src, const uint32_t *__restrict__ pos, uint64_t n)
{
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) dst[pos[i]] = src[i];
}
}  // namespace

namespace ab200 {
// host vector (owned entries, host numbering) -> device vector in the library's node-major layout, and back; asynchronous
int upload_owned(adaflo_b200_mf_t mf, int space, double *dev, const double *host, cudaStream_t s)
{
  const Space &sp = mf->spaces[space];
  const uint64_t n = sp.n_owned_entries();
  if (!sp.numbering) {
    AB_CUDA(cudaMemcpyAsync(dev, host, n * sizeof(double), cudaMemcpyHostToDevice, s));
    return 0;
  }
  AB_CUDA(cudaMemcpyAsync(sp.numbering_stage, host, n * sizeof(double), cudaMemcpyHostToDevice, s));
  if (n) k_from_host_order<<<(unsigned)std::min<uint64_t>((n + 255) / 256, 4096), 256, 0, s>>>(dev, sp.numbering_stage, sp.numbering, n);
  AB_COUNT_LAUNCH();
  AB_CUDA(cudaGetLastError());
  return 0;
}
int download_owned(adaflo_b200_mf_t mf, int space, double *host, const double *dev, cudaStream_t s)
{
  const Space &sp = mf->spaces[space];
  const uint64_t n = sp.n_owned_entries();
  if (!sp.numbering) {
    AB_CUDA(cudaMemcpyAsync(host, dev, n * sizeof(double), cudaMemcpyDeviceToHost, s));
    return 0;
  }
  if (n) k_to_host_order<<<(unsigned)std::min<uint64_t>((n + 255) / 256, 4096), 256, 0, s>>>(sp.numbering_stage, dev, sp.numbering, n);
  AB_COUNT_LAUNCH();
  AB_CUDA(cudaGetLastError());
  AB_CUDA(cudaMemcpyAsync(host, sp.numbering_stage, n * sizeof(double), cudaMemcpyDeviceToHost, s));
  return 0;
}
}  // namespace ab200
}  // extern "C++"

int adaflo_b200_mf_set_dof_numbering(adaflo_b200_mf_t mf, int space, const uint32_t *host_position)
{
  if (!mf || space < 0 || space >= (int)mf->spaces.size()) return fail(ADAFLO_B200_ERR_ARG, "mf_set_dof_numbering: bad argument");
  Space &sp = mf->spaces[space];
  AB_CUDA(cudaSetDevice(mf->ctx->device));
  free_dev(sp.numbering);
  free_dev(sp.numbering_stage);
  sp.numbering = nullptr;
  sp.numbering_stage = nullptr;
  if (!host_position) return 0;  // back to the node-major default
  const uint64_t n = sp.n_owned_entries();
  std::vector<unsigned char> seen(n, 0);
  for (uint64_t i = 0; i < n; ++i) {
    if (host_position[i] >= n || seen[host_position[i]]) return fail(ADAFLO_B200_ERR_ARG, "mf_set_dof_numbering: not a permutation of the owned entries");
    seen[host_position[i]] = 1;
  }
  AB_CUDA(cudaMalloc((void **)&sp.numbering, std::max<uint64_t>(1, n) * sizeof(uint32_t)));
  AB_CUDA(cudaMalloc((void **)&sp.numbering_stage, std::max<uint64_t>(1, n) * sizeof(double)));
  AB_CUDA(cudaMemcpy(sp.numbering, host_position, n * sizeof(uint32_t), cudaMemcpyHostToDevice));
  return 0;
}

int adaflo_b200_mf_upload_vector(adaflo_b200_mf_t mf, int space, double *vec_dev, const double *vec_host)
{
  if (!mf || space < 0 || space >= (int)mf->spaces.size() || !vec_dev || !vec_host) return fail(ADAFLO_B200_ERR_ARG, "mf_upload_vector: bad argument");
  AB_CHECK(upload_owned(mf, space, vec_dev, vec_host, mf->ctx->stream));
  AB_CHECK(stream_wait(mf->ctx, mf->ctx->stream));
  return 0;
}

int adaflo_b200_mf_download_vector(adaflo_b200_mf_t mf, int space, double *vec_host, const double *vec_dev)
{
  if (!mf || space < 0 || space >= (int)mf->spaces.size() || !vec_dev || !vec_host) return fail(ADAFLO_B200_ERR_ARG, "mf_download_vector: bad argument");
  AB_CHECK(download_owned(mf, space, vec_host, vec_dev, mf->ctx->stream));
  AB_CHECK(stream_wait(mf->ctx, mf->ctx->stream));
  return 0;
}

int adaflo_b200_mf_vector_size(adaflo_b200_mf_t mf, int space, uint64_t *n_owned, uint64_t *n_local)
{
  if (!mf || space < 0 || space >= (int)mf->spaces.size()) return fail(ADAFLO_B200_ERR_ARG, "mf_vector_size: bad argument");
  if (n_owned) *n_owned = mf->spaces[space].n_owned_entries();
  if (n_local) *n_local = mf->spaces[space].n_local_entries();
  return 0;
}

int adaflo_b200_mf_n_cells(adaflo_b200_mf_t mf, uint32_t *n_cells)
{
  if (!mf || !n_cells) return fail(ADAFLO_B200_ERR_ARG, "mf_n_cells: null argument");
  *n_cells = mf->n_cells;
  return 0;
}

int adaflo_b200_mf_space_info(adaflo_b200_mf_t mf, int space, int *fe_kind, int *degree, int *n_components)
{
  if (!mf || space < 0 || space >= (int)mf->spaces.size()) return fail(ADAFLO_B200_ERR_ARG, "mf_space_info: bad argument");
  if (fe_kind) *fe_kind = mf->spaces[space].fe_kind;
  if (degree) *degree = mf->spaces[space].degree;
  if (n_components) *n_components = mf->spaces[space].n_comp;
  return 0;
}

int adaflo_b200_mf_update_ghost_values(adaflo_b200_mf_t mf, int space, double *vec)
{
  if (!mf || space < 0 || space >= (int)mf->spaces.size()) return fail(ADAFLO_B200_ERR_ARG, "update_ghost_values: bad argument");
  return comm_update_ghosts(mf, space, vec);
}

int adaflo_b200_mf_compress_add(adaflo_b200_mf_t mf, int space, double *vec)
{
  if (!mf || space < 0 || space >= (int)mf->spaces.size()) return fail(ADAFLO_B200_ERR_ARG, "compress_add: bad argument");
  return comm_compress_add(mf, space, vec);
}

namespace ab200 {
__global__ void k_distribute_constraints(double *vec, const uint32_t first, const uint32_t n_rows, const int n_comp, const uint32_t *ptr,
                                         const uint32_t *cols, const double *w)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows * n_comp) return;
  const uint32_t r = i / n_comp, c = i % n_comp;
  double v = 0.0;
  for (uint32_t k = ptr[r], end = ptr[r + 1]; k < end; ++k) v = fma(w[k], vec[(size_t)n_comp * (cols[k] & ADAFLO_B200_NODE_MASK) + c], v);
  vec[(size_t)n_comp * (first + r) + c] = v;  // rows are closed: no parent is a constrained node
}
}  // namespace ab200

int adaflo_b200_mf_distribute_constraints(adaflo_b200_mf_t mf, int space, double *vec)
{
  if (!mf || !vec || space < 0 || space >= (int)mf->spaces.size()) return fail(ADAFLO_B200_ERR_ARG, "distribute_constraints: bad argument");
  const Space &sp = mf->spaces[space];
  if (!sp.hn_row_ptr) return 0;
  AB_CUDA(cudaSetDevice(mf->ctx->device));
  const uint32_t n_rows = sp.n_owned_nodes - sp.hn_first, n = n_rows * sp.n_comp;
  k_distribute_constraints<<<(n + 127) / 128, 128, 0, mf->ctx->stream>>>(vec, sp.hn_first, n_rows, sp.n_comp, sp.hn_row_ptr, sp.hn_cols, sp.hn_weights);
  AB_COUNT_LAUNCH();
  AB_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------- NavierStokesMatrix
int adaflo_b200_ns_create(adaflo_b200_ctx_t ctx, adaflo_b200_mf_t mf, const adaflo_b200_flow_params *prm, int dof_u, int dof_p,
                          int quad_u, int quad_p, int pressure_average_fix, adaflo_b200_ns_t *out)
{
  if (!ctx || !mf || !prm || !out) return fail(ADAFLO_B200_ERR_ARG, "ns_create: null argument");
  const int ns_ = (int)mf->spaces.size(), nq_ = (int)mf->quads.size();
  if (dof_u < 0 || dof_u >= ns_ || dof_p < 0 || dof_p >= ns_ || quad_u < 0 || quad_u >= nq_ || quad_p < 0 || quad_p >= nq_)
    return fail(ADAFLO_B200_ERR_ARG, "ns_create: dof/quad index out of range");
  const Space &su = mf->spaces[dof_u], &sp = mf->spaces[dof_p];
  // EXPAND_OPERATIONS, navier_stokes_matrix.cc:64-82: the reference dispatches pressure degree 1..5;
  // this library implements degree_p == 1 (Q2/Q1 and Q2/(Q1+DG0))
  // 3-D Q2/Q1(+DG0) runs on the tuned kernels; every other Taylor-Hood pair (velocity degree 2..5, dim 2 or 3) on the generic
  // sum-factorisation kernel of ns_generic.cu.  ADAFLO_B200_NS_GENERIC=1 forces the generic path (tests).
  const char *force_env = std::getenv("ADAFLO_B200_NS_GENERIC");  // read at every create: tests switch it at run time
  const bool force_generic = force_env && force_env[0] == '1';
  const bool tuned = mf->dim == 3 && su.degree == 2 && su.n_comp == 3 && su.nodes_per_cell == 27 && sp.degree == 1 && sp.n_comp == 1 &&
                     (sp.nodes_per_cell == 8 || sp.nodes_per_cell == 9) && mf->quads[quad_u].n_q_1d == 3 && mf->quads[quad_p].n_q_1d == 2;
  const bool generic = !tuned || (force_generic && sp.nodes_per_cell == 8 && !su.hanging_masks && prm->density_diff == 0 && prm->viscosity_diff == 0 &&
                                  (prm->linearization == ADAFLO_B200_COUPLED_IMPLICIT_NEWTON || prm->linearization == ADAFLO_B200_COUPLED_IMPLICIT_PICARD));
  if (generic) {
    const int K = su.degree, dim = mf->dim;
    const uint32_t nn = dim == 2 ? (K + 1) * (K + 1) : (K + 1) * (K + 1) * (K + 1), npn = dim == 2 ? K * K : K * K * K;
    if (K < 2 || K > 5 || sp.degree != K - 1 || su.fe_kind != ADAFLO_B200_FE_Q || sp.fe_kind != ADAFLO_B200_FE_Q || su.n_comp != dim ||
        sp.n_comp != 1 || su.nodes_per_cell != nn || sp.nodes_per_cell != npn || mf->quads[quad_u].n_q_1d != K + 1)
      return fail(ADAFLO_B200_ERR_UNSUPPORTED, "ns_create: Taylor-Hood Q_k/Q_(k-1) with k = 2..5, FE_Q spaces, QGauss(k+1), dim 2 or 3 "
                                               "(navier_stokes_matrix.cc:64-82); the element-constant pressure needs 3-D Q2/Q1");
    if (su.hanging_masks || sp.hanging_masks || prm->density_diff != 0 || prm->viscosity_diff != 0 || prm->augmented_taylor_hood)
      return fail(ADAFLO_B200_ERR_UNSUPPORTED, "ns_create: hanging nodes, variable coefficients and the augmented pressure are implemented "
                                               "for 3-D Q2/Q1 only");
    if (prm->linearization != ADAFLO_B200_COUPLED_IMPLICIT_NEWTON && prm->linearization != ADAFLO_B200_COUPLED_IMPLICIT_PICARD)
      return fail(ADAFLO_B200_ERR_UNSUPPORTED, "ns_create: the generic-degree path implements the coupled implicit Newton and Picard linearisations");
  }
  if ((su.hanging_masks != nullptr) != (sp.hanging_masks != nullptr))
    return fail(ADAFLO_B200_ERR_ARG, "ns_create: hanging-node masks must be given for both spaces or for none");
  if (!generic && (su.hn_row_ptr || sp.hn_row_ptr))
    return fail(ADAFLO_B200_ERR_UNSUPPORTED, "ns_create: the tuned 3-D Q2/Q1 kernels take hanging_masks, not general constraint rows");
  AB_CUDA(cudaSetDevice(ctx->device));
  adaflo_b200_ns_t ns = new (std::nothrow) adaflo_b200_ns_s();
  if (!ns) return fail(ADAFLO_B200_ERR_ARG, "out of host memory");
  std::memset((void *)ns, 0, sizeof(*ns));
  ns->ctx = ctx;
  ns->mf = mf;
  ns->prm = *prm;
  ns->tw = adaflo_b200_time_weights{1.0, -1.0, 0.0, 1.0, 1.0, 0.0};
  ns->dof_u = dof_u;
  ns->dof_p = dof_p;
  ns->quad_u = quad_u;
  ns->quad_p = quad_p;
  ns->n_q = 27;
  ns->generic = generic;
  ns->gen_dim = mf->dim;
  ns->gen_degree_u = su.degree;
  ns->var = prm->density_diff != 0 || prm->viscosity_diff != 0;  // :97-98
  cudaStream_t s = ctx->stream;
  size_t size = n_groups(mf->n_cells) * 270;  // decade-blocked per-point arrays
  if (generic) {  // plain [cell][row][q] with dim + dim^2 rows; size * 12 below covers it
    const int n1 = su.degree + 1;
    ns->n_q = mf->dim == 2 ? n1 * n1 : n1 * n1 * n1;
    size = ((size_t)mf->n_cells * ns->n_q * (mf->dim + mf->dim * mf->dim) + 11) / 12;
  }
  if (ns->var) {  // :100-106
    AB_CUDA(cudaMalloc((void **)&ns->rho, size * sizeof(double)));
    AB_CUDA(cudaMalloc((void **)&ns->mu, size * sizeof(double)));
    AB_CUDA(cudaMalloc((void **)&ns->damp, size * sizeof(double)));
    AB_CHECK(vec_set(s, ns->rho, prm->density, size));
    AB_CHECK(vec_set(s, ns->mu, prm->viscosity, size));
    AB_CHECK(vec_set(s, ns->damp, 0.0, size));
  }
  if (prm->linearization != ADAFLO_B200_COUPLED_VELOCITY_EXPLICIT) {  // :107-108
    AB_CUDA(cudaMalloc((void **)&ns->lin, size * 12 * sizeof(double)));
    AB_CUDA(cudaMemsetAsync(ns->lin, 0, size * 12 * sizeof(double), s));
  }
  const uint64_t n_p_local = sp.n_local_entries(), n_p_owned = sp.n_owned_entries();
  AB_CUDA(cudaMalloc((void **)&ns->tmp_p, (n_p_local + 1) * sizeof(double)));

  // pressure null space, :110-168
  if (pressure_average_fix || prm->augmented_taylor_hood) {
    double *pres_mass = ns->tmp_p;
    AB_CUDA(cudaMemsetAsync(pres_mass, 0, n_p_local * sizeof(double), s));
    PArgs pa{};
    pa.n_cells = mf->n_cells;
    pa.np_cell = (int)sp.nodes_per_cell;
    pa.cell_p = sp.cell_nodes;
    pa.hn_p = sp.hanging_masks;
    pa.dst = pres_mass;
    pa.c0 = 1.0;
    fill_geometry(pa, mf->quads[quad_p]);
    if (generic) {  // integrals of the pressure shape functions with the velocity quadrature (exact for them)
      NSArgs ga = make_args(ns, false);
      ga.dst_p = pres_mass;
      if (mf->quads[quad_p].n_q_1d == su.degree) {  // the pressure quadrature QGauss(degree_p + 1), like the reference
        fill_geometry(ga, mf->quads[quad_p]);
        AB_CHECK(launch_ns_generic(ga, mf->dim, su.degree - 1, 3, 1, 0, mf->quads[quad_p].geometry_type, s));
      } else
        AB_CHECK(launch_ns_generic(ga, mf->dim, su.degree, 2, 1, 0, mf->quads[quad_u].geometry_type, s));
    } else
      AB_CHECK(launch_pressure_op(pa, 2, 2, s));
    AB_CHECK(comm_compress_add(mf, dof_p, pres_mass));
    std::vector<double> pm(n_p_local), m(n_p_local), w(n_p_local);
    AB_CUDA(cudaMemcpyAsync(pm.data(), pres_mass, n_p_local * sizeof(double), cudaMemcpyDeviceToHost, s));
    std::vector<uint32_t> con(sp.n_constrained);
    if (sp.n_constrained)
      AB_CUDA(cudaMemcpyAsync(con.data(), sp.constrained, sp.n_constrained * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    AB_CUDA(cudaStreamSynchronize(s));
    std::vector<char> is_dg0(n_p_local, 0);
    for (uint32_t dof : sp.dg0_dofs) is_dg0[dof] = 1;
    for (int mode = 0; mode < 2; ++mode) {
      if (mode == 0 && !pressure_average_fix) continue;
      if (mode == 1 && !prm->augmented_taylor_hood) continue;
      if (mode == 1 && sp.fe_kind != ADAFLO_B200_FE_Q_DG0)
        return fail(ADAFLO_B200_ERR_ARG, "ns_create: augmented_taylor_hood requires an FE_Q_DG0 pressure space");
      std::fill(m.begin(), m.end(), 0.0);
      std::fill(w.begin(), w.end(), 0.0);
      for (uint64_t i = 0; i < n_p_owned; ++i)
        if ((mode == 1) == (is_dg0[i] != 0)) {
          m[i] = 1.0;
          w[i] = pm[i];
        }
      for (uint32_t c : con) m[c] = 0.0;  // :159-163
      AB_CHECK(upload(&ns->modes[mode], m.data(), n_p_local, s));
      AB_CHECK(upload(&ns->mode_weights[mode], w.data(), n_p_local, s));
      const double *ys[1] = {ns->mode_weights[mode]};
      double dot = 0;
      AB_CHECK(vec_multi_dot(ctx, ns->modes[mode], ys, 1, n_p_owned, &dot, true));
      ns->inv_weights[mode] = 1.0 / dot;  // :165-166
      ns->has_mode[mode] = true;
    }
  }
  AB_CUDA(cudaStreamSynchronize(s));
  *out = ns;
  return 0;
}

int adaflo_b200_ns_destroy(adaflo_b200_ns_t ns)
{
  if (!ns) return 0;
  cudaSetDevice(ns->ctx->device);
  cudaStreamSynchronize(ns->ctx->stream);
  void *ptrs[] = {ns->lin, ns->rho, ns->mu, ns->damp, ns->lin_prec, ns->rho_prec, ns->mu_prec, ns->damp_prec, ns->modes[0],
                  ns->modes[1], ns->mode_weights[0], ns->mode_weights[1], ns->stage_u[0], ns->stage_u[1], ns->stage_p[0],
                  ns->stage_p[1], ns->tmp_p};
  for (void *p : ptrs) free_dev(p);
  if (ns->pipe_events) {
    for (cudaEvent_t e : *ns->pipe_events) cudaEventDestroy(e);
    delete ns->pipe_events;
  }
  if (ns->up_stream) cudaStreamDestroy(ns->up_stream);
  if (ns->down_stream) cudaStreamDestroy(ns->down_stream);
  if (ns->kernel_events) {
    for (auto &e : *ns->kernel_events) {
      cudaEventDestroy(e.first);
      cudaEventDestroy(e.second);
    }
    delete ns->kernel_events;
  }
  delete ns;
  return 0;
}

int adaflo_b200_ns_set_flow_params(adaflo_b200_ns_t ns, const adaflo_b200_flow_params *prm)
{
  if (!ns || !prm) return fail(ADAFLO_B200_ERR_ARG, "null argument");
  const bool var = prm->density_diff != 0 || prm->viscosity_diff != 0;
  if (var != ns->var) return fail(ADAFLO_B200_ERR_STATE, "set_flow_params: cannot change the variable-coefficient mode after create");
  if (prm->linearization != ADAFLO_B200_COUPLED_VELOCITY_EXPLICIT && !ns->lin)
    return fail(ADAFLO_B200_ERR_STATE, "set_flow_params: linearization storage was not allocated at create");
  ns->prm = *prm;
  return 0;
}

int adaflo_b200_ns_set_time_weights(adaflo_b200_ns_t ns, const adaflo_b200_time_weights *w)
{
  if (!ns || !w) return fail(ADAFLO_B200_ERR_ARG, "null argument");
  ns->tw = *w;
  return 0;
}

int adaflo_b200_ns_set_old_solutions(adaflo_b200_ns_t ns, const double *old_u, const double *old_old_u)
{
  if (!ns) return fail(ADAFLO_B200_ERR_ARG, "null argument");
  ns->old_u = old_u;
  ns->old_old_u = old_old_u;
  return 0;
}

static int stats_begin(adaflo_b200_ns_t ns, std::chrono::steady_clock::time_point &t0)
{
  t0 = std::chrono::steady_clock::now();
  (void)ns;
  return 0;
}

static void stats_end(adaflo_b200_ns_t ns, const std::chrono::steady_clock::time_point &t0)
{
  // matvec_timer semantics (:227,260-261): host wall time of the call.  Launches are
  // asynchronous; the time of the last call is completed by the next synchronising call.
  ns->matvec_count++;
  ns->matvec_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

static int vmult_impl(adaflo_b200_ns_t ns, double *dst_u, double *dst_p, const double *src_u, const double *src_p, bool transpose)
{
  if (!ns) return fail(ADAFLO_B200_ERR_STATE, "vmult: operator not initialised");
  if (!dst_u || !dst_p || !src_u || !src_p) return fail(ADAFLO_B200_ERR_ARG, "vmult: null vector");
  adaflo_b200_mf_t mf = ns->mf;
  adaflo_b200_ctx_t ctx = ns->ctx;
  cudaStream_t s = ctx->stream;
  const Space &su = mf->spaces[ns->dof_u], &sp = mf->spaces[ns->dof_p];
  const int mode = mode_of(ns->prm);
  if ((mode == MODE_NEWTON || mode == MODE_PICARD) && !ns->lin) return fail(ADAFLO_B200_ERR_STATE, "vmult: no linearization storage");
  std::chrono::steady_clock::time_point t0;
  stats_begin(ns, t0);
  const double *in_p = src_p;
  if (transpose && projection_active(ns) && (ns->has_mode[0] || ns->has_mode[1])) {
    // (P1 P0)^T = P0^T P1^T applied to the pressure source
    AB_CUDA(cudaMemcpyAsync(ns->tmp_p, src_p, sp.n_local_entries() * sizeof(double), cudaMemcpyDeviceToDevice, s));
    for (int m = 1; m >= 0; --m)
      if (ns->has_mode[m])
        AB_CHECK(vec_project(ctx, ns->tmp_p, ns->modes[m], ns->mode_weights[m], ns->inv_weights[m], sp.n_owned_entries(),
                             sp.n_local_entries()));
    in_p = ns->tmp_p;
  }
  // cell_loop: ghost import of src, cells, compress(add) of dst (:236-242).  Like deal.II, the import is
  // overlapped with the cells that touch no ghost dof: it runs on the communication stream while the
  // interior decades run on the compute stream.
  // Measured on 2 x B200, 128^3 cells per GPU: 2.21 ms overlapped vs 2.12 ms serial -- the 1.6 MB import costs
  // less than the tail of a second kernel launch, so the overlap is opt-in (ADAFLO_B200_HALO_OVERLAP=1).
  static const bool overlap_enabled = [] {
    const char *k = std::getenv("ADAFLO_B200_VMULT_KERNEL"), *o = std::getenv("ADAFLO_B200_HALO_OVERLAP");
    return !(k && !std::strcmp(k, "v1")) && (o && !std::strcmp(o, "1"));  // the legacy v1 kernel has no decade range
  }();
  const bool overlap = overlap_enabled && !ns->generic && ctx->nranks > 1 && mf->ghost_range_valid && ctx->comm_stream;
  // Default on several GPUs: the import runs on the communication stream next to the zero-fill of dst (0.42 GB per GPU at
  // 128^3 cells: 65-100 us, longer than the import), and the cell kernel -- ONE launch -- waits for both.  This hides the
  // import without the second launch that made the split overlap above slower (ADAFLO_B200_FILL_OVERLAP=0 switches it off).
  static const bool fill_overlap_enabled = [] {
    const char *o = std::getenv("ADAFLO_B200_FILL_OVERLAP");
    return !(o && !std::strcmp(o, "0"));
  }();
  const bool fill_overlap = !overlap && fill_overlap_enabled && ctx->nranks > 1 && ctx->comm_stream;
  if (ctx->nranks > 1) {
    const int spaces[2] = {ns->dof_u, ns->dof_p};
    double *vecs[2] = {const_cast<double *>(src_u), const_cast<double *>(in_p)};
    if (overlap || fill_overlap) {
      AB_CUDA(cudaEventRecord(ctx->ev_a, s));
      AB_CUDA(cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_a, 0));
      cudaStream_t keep = ctx->stream;
      ctx->stream = ctx->comm_stream;
      const int rc = comm_update_ghosts_n(mf, 2, spaces, vecs);
      ctx->stream = keep;
      AB_CHECK(rc);
      AB_CUDA(cudaEventRecord(ctx->ev_b, ctx->comm_stream));
    } else
      AB_CHECK(comm_update_ghosts_n(mf, 2, spaces, vecs));  // one NCCL group for both blocks
  }
  AB_CUDA(cudaMemsetAsync(dst_u, 0, su.n_local_entries() * sizeof(double), s));  // dst = 0, :229
  AB_CUDA(cudaMemsetAsync(dst_p, 0, sp.n_local_entries() * sizeof(double), s));
  if (fill_overlap) AB_CUDA(cudaStreamWaitEvent(s, ctx->ev_b, 0));
  NSArgs a = make_args(ns, false);
  a.src_u = src_u;
  a.src_p = in_p;
  a.dst_u = dst_u;
  a.dst_p = dst_p;
  a.transpose = transpose ? 1 : 0;
  {  // paced source prefetch (ns_r1.cuh): stride in doubles, even, so that 16-byte aligned vectors give aligned requests
    const size_t n_dec = n_groups(mf->n_cells);
    const uint64_t stride = n_dec ? ((su.n_local_entries() + n_dec - 1) / n_dec + 1) & ~(uint64_t)1 : 0;
    const bool aligned = (reinterpret_cast<uintptr_t>(src_u) & 15u) == 0;
    a.pf_stride = aligned ? stride : 0;
    a.pf_u_entries = su.n_local_entries();
  }
  if (ns->time_kernel) {
    if (ns->kernel_events_used == ns->kernel_events->size()) {
      cudaEvent_t e0, e1;
      AB_CUDA(cudaEventCreate(&e0));
      AB_CUDA(cudaEventCreate(&e1));
      ns->kernel_events->push_back({e0, e1});
    }
    AB_CUDA(cudaEventRecord((*ns->kernel_events)[ns->kernel_events_used].first, s));
  }
  if (overlap) {  // interior decades first, then wait for the import and run the decades touching ghosts
    const size_t n_dec = n_groups(mf->n_cells);
    NSArgs b = a;
    if (mf->ghost_decade_begin > 0) {
      b.decade_begin = 0;
      b.decade_end = mf->ghost_decade_begin;
      AB_CHECK(launch_cells(ns, b, true));
    }
    if (mf->ghost_decade_end < n_dec) {
      b.decade_begin = mf->ghost_decade_end;
      b.decade_end = n_dec;
      AB_CHECK(launch_cells(ns, b, true));
    }
    AB_CUDA(cudaStreamWaitEvent(s, ctx->ev_b, 0));
    a.decade_begin = mf->ghost_decade_begin;
    a.decade_end = mf->ghost_decade_end;
  }
  const bool cells_left = !overlap || a.decade_end > a.decade_begin;
  if (cells_left) AB_CHECK(launch_cells(ns, a, true));
  if (ns->time_kernel) AB_CUDA(cudaEventRecord((*ns->kernel_events)[ns->kernel_events_used++].second, s));
  if (ctx->nranks > 1) {
    const int spaces[2] = {ns->dof_u, ns->dof_p};
    double *vecs[2] = {dst_u, dst_p};
    AB_CHECK(comm_compress_add_n(mf, 2, spaces, vecs));
  }
  // constrained rows, :247-256
  AB_CHECK(vec_copy_indexed(s, dst_u, src_u, su.constrained, su.n_constrained, 1.0));
  AB_CHECK(vec_copy_indexed(s, dst_p, in_p, sp.constrained, sp.n_constrained, -1.0));
  if (!transpose && projection_active(ns))  // :258
    for (int m = 0; m < 2; ++m)
      if (ns->has_mode[m])
        AB_CHECK(vec_project(ctx, dst_p, ns->mode_weights[m], ns->modes[m], ns->inv_weights[m], sp.n_owned_entries(),
                             sp.n_local_entries()));
  stats_end(ns, t0);
  return 0;
}

int adaflo_b200_ns_vmult(adaflo_b200_ns_t ns, double *dst_u, double *dst_p, const double *src_u, const double *src_p)
{
  return vmult_impl(ns, dst_u, dst_p, src_u, src_p, false);
}

int adaflo_b200_ns_Tvmult(adaflo_b200_ns_t ns, double *dst_u, double *dst_p, const double *src_u, const double *src_p)
{
  return vmult_impl(ns, dst_u, dst_p, src_u, src_p, true);
}

int adaflo_b200_ns_residual(adaflo_b200_ns_t ns, double *rhs_u, double *rhs_p, const double *src_u, const double *src_p,
                            const double *user_u, const double *user_p)
{
  if (!ns) return fail(ADAFLO_B200_ERR_STATE, "residual: operator not initialised");
  if (!rhs_u || !rhs_p || !src_u || !src_p || !user_u || !user_p) return fail(ADAFLO_B200_ERR_ARG, "residual: null vector");
  adaflo_b200_mf_t mf = ns->mf;
  adaflo_b200_ctx_t ctx = ns->ctx;
  cudaStream_t s = ctx->stream;
  const Space &su = mf->spaces[ns->dof_u], &sp = mf->spaces[ns->dof_p];
  const bool needs_old = ns->prm.physical_type == ADAFLO_B200_INCOMPRESSIBLE;
  if (needs_old && (!ns->old_u || !ns->old_old_u)) return fail(ADAFLO_B200_ERR_STATE, "residual: old solutions not set");
  if (ns->prm.physical_type != ADAFLO_B200_STOKES && ns->prm.linearization != ADAFLO_B200_COUPLED_VELOCITY_EXPLICIT && !ns->lin)
    return fail(ADAFLO_B200_ERR_STATE, "residual: no linearization storage");
  if (ctx->nranks > 1) {
    AB_CHECK(comm_update_ghosts(mf, ns->dof_u, const_cast<double *>(src_u)));
    AB_CHECK(comm_update_ghosts(mf, ns->dof_p, const_cast<double *>(src_p)));
    if (needs_old) {
      AB_CHECK(comm_update_ghosts(mf, ns->dof_u, const_cast<double *>(ns->old_u)));
      AB_CHECK(comm_update_ghosts(mf, ns->dof_u, const_cast<double *>(ns->old_old_u)));
    }
    // ghost part of the accumulation target must start from zero
    if (su.n_ghost_nodes) AB_CUDA(cudaMemsetAsync(rhs_u + su.n_owned_entries(), 0, sizeof(double) * su.n_ghost_nodes * su.n_comp, s));
    if (sp.n_ghost_nodes) AB_CUDA(cudaMemsetAsync(rhs_p + sp.n_owned_entries(), 0, sizeof(double) * sp.n_ghost_nodes * sp.n_comp, s));
  }
  NSArgs a = make_args(ns, false);
  a.src_u = src_u;
  a.src_p = src_p;
  a.old_u = ns->old_u;
  a.old_old_u = ns->old_old_u;
  a.dst_u = rhs_u;
  a.dst_p = rhs_p;
  if (ns->generic)
    AB_CHECK(launch_ns_generic(a, ns->gen_dim, ns->gen_degree_u, 1, 1, mode_of(ns->prm), mf->quads[ns->quad_u].geometry_type, s));
  else
    AB_CHECK(launch_ns_residual(a, mf->quads[ns->quad_u].geometry_type, ns->var, s));
  if (ctx->nranks > 1) {
    AB_CHECK(comm_compress_add(mf, ns->dof_u, rhs_u));
    AB_CHECK(comm_compress_add(mf, ns->dof_p, rhs_p));
  }
  // system_rhs.sadd(-1., 1., user_rhs), :292
  AB_CHECK(vec_sadd(s, rhs_u, -1.0, 1.0, user_u, su.n_owned_entries()));
  AB_CHECK(vec_sadd(s, rhs_p, -1.0, 1.0, user_p, sp.n_owned_entries()));
  return 0;
}

int adaflo_b200_ns_velocity_vmult(adaflo_b200_ns_t ns, double *dst_u, const double *src_u)
{
  if (!ns) return fail(ADAFLO_B200_ERR_STATE, "velocity_vmult: operator not initialised");
  if (!dst_u || !src_u) return fail(ADAFLO_B200_ERR_ARG, "velocity_vmult: null vector");
  adaflo_b200_mf_t mf = ns->mf;
  cudaStream_t s = ns->ctx->stream;
  const Space &su = mf->spaces[ns->dof_u];
  if (ns->ctx->nranks > 1) AB_CHECK(comm_update_ghosts(mf, ns->dof_u, const_cast<double *>(src_u)));
  AB_CUDA(cudaMemsetAsync(dst_u, 0, su.n_local_entries() * sizeof(double), s));
  NSArgs a = make_args(ns, true);  // preconditioner snapshot if present, :349-355
  a.src_u = src_u;
  a.dst_u = dst_u;
  AB_CHECK(launch_cells(ns, a, false));
  if (ns->ctx->nranks > 1) AB_CHECK(comm_compress_add(mf, ns->dof_u, dst_u));
  AB_CHECK(vec_copy_indexed(s, dst_u, src_u, su.constrained, su.n_constrained, 1.0));  // :377-381
  return 0;
}

int adaflo_b200_ns_divergence_vmult_add(adaflo_b200_ns_t ns, double *dst_p, const double *src_u, int weight_by_viscosity)
{
  if (!ns) return fail(ADAFLO_B200_ERR_STATE, "divergence_vmult_add: operator not initialised");
  if (!dst_p || !src_u) return fail(ADAFLO_B200_ERR_ARG, "divergence_vmult_add: null vector");
  if (ns->generic) {  // degree / dimension dispatch: constant coefficients, single rank
    NSArgs a = make_args(ns, false);
    a.src_u = src_u;
    a.dst_p = dst_p;
    a.div_weight = weight_by_viscosity ? -ns->prm.viscosity : -1.0;  // :948-953
    return launch_ns_generic(a, ns->gen_dim, ns->gen_degree_u, 4, 0, mode_of(ns->prm), ns->mf->quads[ns->quad_u].geometry_type, ns->ctx->stream);
  }
  adaflo_b200_mf_t mf = ns->mf;
  cudaStream_t s = ns->ctx->stream;
  const Space &sp = mf->spaces[ns->dof_p];
  if (ns->ctx->nranks > 1) {
    AB_CHECK(comm_update_ghosts(mf, ns->dof_u, const_cast<double *>(src_u)));
    if (sp.n_ghost_nodes) AB_CUDA(cudaMemsetAsync(dst_p + sp.n_owned_entries(), 0, sizeof(double) * sp.n_ghost_nodes, s));
  }
  PArgs pa{};
  pa.n_cells = mf->n_cells;
  pa.np_cell = (int)sp.nodes_per_cell;
  pa.cell_u = mf->spaces[ns->dof_u].cell_nodes;
  pa.cell_p = sp.cell_nodes;
  pa.hn_u = mf->spaces[ns->dof_u].hanging_masks;
  pa.hn_p = sp.hanging_masks;
  pa.src = src_u;
  pa.dst = dst_p;
  pa.resolve = ns->prm.linearization != ADAFLO_B200_PROJECTION;  // :940-943
  pa.c0 = ns->prm.viscosity;
  pa.coeff_q = (weight_by_viscosity && ns->var) ? ns->mu : nullptr;  // :948-953
  fill_geometry(pa, mf->quads[ns->quad_u]);
  AB_CHECK(launch_divergence(pa, weight_by_viscosity != 0, s));
  if (ns->ctx->nranks > 1) AB_CHECK(comm_compress_add(mf, ns->dof_p, dst_p));
  return 0;
}

static int pressure_common(adaflo_b200_ns_t ns, double *dst_p, const double *src_p, PArgs &pa, int kind, int quad)
{
  adaflo_b200_mf_t mf = ns->mf;
  cudaStream_t s = ns->ctx->stream;
  const Space &sp = mf->spaces[ns->dof_p];
  if (ns->ctx->nranks > 1) AB_CHECK(comm_update_ghosts(mf, ns->dof_p, const_cast<double *>(src_p)));
  AB_CUDA(cudaMemsetAsync(dst_p, 0, sp.n_local_entries() * sizeof(double), s));
  pa.n_cells = mf->n_cells;
  pa.np_cell = (int)sp.nodes_per_cell;
  pa.cell_p = sp.cell_nodes;
  pa.hn_p = sp.hanging_masks;
  pa.src = src_p;
  pa.dst = dst_p;
  fill_geometry(pa, mf->quads[quad]);
  AB_CHECK(launch_pressure_op(pa, kind, mf->quads[quad].n_q_1d, s));
  if (ns->ctx->nranks > 1) AB_CHECK(comm_compress_add(mf, ns->dof_p, dst_p));
  AB_CHECK(vec_copy_indexed(s, dst_p, src_p, sp.constrained, sp.n_constrained, 1.0));
  return 0;
}

// pressure mass / Poisson of the degree / dimension dispatch: constant coefficient c0, the pressure quadrature QGauss(degree_p + 1)
static int generic_pressure_op(adaflo_b200_ns_t ns, double *dst_p, const double *src_p, int kind, double c0)
{
  adaflo_b200_mf_t mf = ns->mf;
  cudaStream_t s = ns->ctx->stream;
  const Space &sp = mf->spaces[ns->dof_p];
  if (mf->quads[ns->quad_p].n_q_1d != sp.degree + 1)
    return fail(ADAFLO_B200_ERR_UNSUPPORTED, "generic pressure operators: the pressure quadrature must be QGauss(degree_p + 1) (navier_stokes_matrix.cc:1003-1005)");
  AB_CUDA(cudaMemsetAsync(dst_p, 0, sp.n_local_entries() * sizeof(double), s));
  NSArgs a = make_args(ns, false);
  fill_geometry(a, mf->quads[ns->quad_p]);
  a.src_p = src_p;
  a.dst_p = dst_p;
  a.scalar_c0 = c0;
  AB_CHECK(launch_ns_generic_scalar(a, ns->gen_dim, sp.degree, kind, mf->quads[ns->quad_p].geometry_type, s));
  return vec_copy_indexed(s, dst_p, src_p, sp.constrained, sp.n_constrained, 1.0);  // :411-415, :443-446
}

int adaflo_b200_ns_pressure_mass_vmult(adaflo_b200_ns_t ns, double *dst_p, const double *src_p)
{
  if (!ns) return fail(ADAFLO_B200_ERR_STATE, "pressure_mass_vmult: operator not initialised");
  if (!dst_p || !src_p) return fail(ADAFLO_B200_ERR_ARG, "pressure_mass_vmult: null vector");
  PArgs pa{};
  const auto &p = ns->prm;
  if (ns->generic)  // :1059-1063 with constant coefficients
    return generic_pressure_op(ns, dst_p, src_p, 0, p.physical_type == ADAFLO_B200_INCOMPRESSIBLE_STATIONARY ? 1.0 : 1.0 / (p.viscosity + p.tau_grad_div));
  if (p.linearization == ADAFLO_B200_PROJECTION || p.physical_type == ADAFLO_B200_INCOMPRESSIBLE_STATIONARY) {
    pa.coeff_mode = 0;
    pa.c0 = 1.0;  // :1059-1062
  } else if (ns->var) {
    pa.coeff_mode = 1;
    pa.coeff_q = ns->mu_prec ? ns->mu_prec : ns->mu;  // :430-431
    pa.c1 = p.tau_grad_div;
  } else {
    pa.coeff_mode = 0;
    pa.c0 = 1.0 / (p.viscosity + p.tau_grad_div);  // :1063
  }
  AB_CHECK(pressure_common(ns, dst_p, src_p, pa, 0, ns->quad_p));
  if (ns->has_mode[1] && p.linearization != ADAFLO_B200_PROJECTION) {  // :449-454
    const Space &sp = ns->mf->spaces[ns->dof_p];
    AB_CHECK(vec_project(ns->ctx, dst_p, ns->mode_weights[1], ns->modes[1], ns->inv_weights[1], sp.n_owned_entries(),
                         sp.n_local_entries()));
  }
  return 0;
}

int adaflo_b200_ns_pressure_poisson_vmult(adaflo_b200_ns_t ns, double *dst_p, const double *src_p)
{
  if (!ns) return fail(ADAFLO_B200_ERR_STATE, "pressure_poisson_vmult: operator not initialised");
  if (!dst_p || !src_p) return fail(ADAFLO_B200_ERR_ARG, "pressure_poisson_vmult: null vector");
  PArgs pa{};
  const auto &p = ns->prm;
  if (ns->generic)  // :1003-1028 with constant coefficients
    return generic_pressure_op(ns, dst_p, src_p, 1, p.physical_type == ADAFLO_B200_INCOMPRESSIBLE_STATIONARY ? 1.0 : 1.0 / (ns->tw.weight * p.density));
  const bool var = ns->var && p.linearization != ADAFLO_B200_PROJECTION;  // :975-977
  const double *rho = ns->rho_prec ? ns->rho_prec : ns->rho;             // :398-399
  const bool stationary = p.physical_type == ADAFLO_B200_INCOMPRESSIBLE_STATIONARY;
  int quad = ns->quad_p;
  if (var && !stationary) {  // :979-1001
    pa.coeff_mode = 2;
    pa.coeff_q = rho;
    pa.c0 = ns->tw.weight;
    quad = ns->quad_u;
  } else if (stationary) {
    pa.coeff_mode = 0;
    pa.c0 = 1.0;  // :1020-1022
  } else if (var) {
    pa.coeff_mode = 3;
    pa.coeff_q = rho;
    pa.c0 = ns->tw.weight;
  } else {
    pa.coeff_mode = 0;
    const double rho_value = p.density < p.density + p.density_diff ? p.density : p.density + p.density_diff;  // :1017-1019
    pa.c0 = 1.0 / (ns->tw.weight * rho_value);
  }
  return pressure_common(ns, dst_p, src_p, pa, 1, quad);
}

int adaflo_b200_ns_pressure_convdiff_vmult(adaflo_b200_ns_t ns, double *dst_p, const double *src_p)
{
  if (ns && ns->generic) return fail(ADAFLO_B200_ERR_UNSUPPORTED, "this entry point is implemented for the 3-D Q2/Q1 operator only (generic-degree path: vmult, residual, velocity_vmult, projection)");
  if (!ns) return fail(ADAFLO_B200_ERR_STATE, "pressure_convdiff_vmult: operator not initialised");
  if (!dst_p || !src_p) return fail(ADAFLO_B200_ERR_ARG, "pressure_convdiff_vmult: null vector");
  if (!ns->lin) return fail(ADAFLO_B200_ERR_UNSUPPORTED, "pressure_convdiff_vmult needs linearized velocities (:1111)");
  PArgs pa{};
  if (ns->var) {
    pa.coeff_mode = 4;
    pa.coeff_q = ns->mu;
  } else {
    pa.coeff_mode = 0;
    pa.c0 = ns->prm.viscosity;
  }
  return pressure_common(ns, dst_p, src_p, pa, 1, ns->quad_u);
}

int adaflo_b200_ns_compute_diagonal(adaflo_b200_ns_t ns, double *diag_u, double *diag_p)
{
  // Exact diagonal of vmult.  Velocity block: diag(C^T A_cell C) summed over cells (one unit-source pass per
  // cell-local slot inside the cell kernel), 1 on constrained rows.  Pressure block: the operator has no
  // pressure-pressure coupling and the mean projection removes a multiple of (A^T w)_p = 0, so the diagonal is
  // 0, and -1 on constrained rows (navier_stokes_matrix.cc:247-256).
  if (!ns) return fail(ADAFLO_B200_ERR_STATE, "compute_diagonal: operator not initialised");
  if (!diag_u || !diag_p) return fail(ADAFLO_B200_ERR_ARG, "compute_diagonal: null vector");
  adaflo_b200_mf_t mf = ns->mf;
  cudaStream_t s = ns->ctx->stream;
  const Space &su = mf->spaces[ns->dof_u], &sp = mf->spaces[ns->dof_p];
  AB_CUDA(cudaMemsetAsync(diag_u, 0, su.n_local_entries() * sizeof(double), s));
  AB_CUDA(cudaMemsetAsync(diag_p, 0, sp.n_local_entries() * sizeof(double), s));
  NSArgs a = make_args(ns, false);
  a.src_u = diag_u;  // not read in diagonal mode
  a.dst_u = diag_u;
  a.diag = 1;
  if (ns->generic) {  // one pass per cell-local slot (the meshes of this path are small); on hanging nodes see ns_generic.cu
    const int n1 = ns->gen_degree_u + 1, nn = ns->gen_dim == 2 ? n1 * n1 : n1 * n1 * n1;
    for (int slot = 0; slot < nn * ns->gen_dim; ++slot) {
      a.diag = slot + 1;
      AB_CHECK(launch_cells(ns, a, false));
    }
  } else
    AB_CHECK(launch_cells(ns, a, false));
  if (ns->ctx->nranks > 1) AB_CHECK(comm_compress_add(mf, ns->dof_u, diag_u));
  AB_CHECK(vec_set_indexed(s, diag_u, su.constrained, su.n_constrained, 1.0));
  AB_CHECK(vec_set_indexed(s, diag_p, sp.constrained, sp.n_constrained, -1.0));
  return 0;
}

int adaflo_b200_ns_fix_linearization_point(adaflo_b200_ns_t ns)
{
  if (!ns) return fail(ADAFLO_B200_ERR_STATE, "operator not initialised");
  cudaStream_t s = ns->ctx->stream;
  size_t size = n_groups(ns->mf->n_cells) * 270;
  if (ns->generic) size = ((size_t)ns->mf->n_cells * ns->n_q * (ns->gen_dim + ns->gen_dim * ns->gen_dim) + 11) / 12;
  struct { double **dst; double *src; size_t n; } items[] = {{&ns->lin_prec, ns->lin, size * 12},
                                                             {&ns->rho_prec, ns->rho, size},
                                                             {&ns->mu_prec, ns->mu, size},
                                                             {&ns->damp_prec, ns->damp, size}};
  for (auto &it : items) {
    if (!it.src) continue;
    if (!*it.dst) AB_CUDA(cudaMalloc((void **)it.dst, it.n * sizeof(double)));
    AB_CUDA(cudaMemcpyAsync(*it.dst, it.src, it.n * sizeof(double), cudaMemcpyDeviceToDevice, s));
  }
  return 0;
}

int adaflo_b200_ns_apply_pressure_average_projection(adaflo_b200_ns_t ns, double *p)
{
  if (!ns || !p) return fail(ADAFLO_B200_ERR_ARG, "null argument");
  if (!projection_active(ns)) return 0;
  const Space &sp = ns->mf->spaces[ns->dof_p];
  for (int m = 0; m < 2; ++m)
    if (ns->has_mode[m])
      AB_CHECK(vec_project(ns->ctx, p, ns->mode_weights[m], ns->modes[m], ns->inv_weights[m], sp.n_owned_entries(), sp.n_local_entries()));
  return 0;
}

int adaflo_b200_ns_apply_pressure_shift(adaflo_b200_ns_t ns, double shift, double *p)
{
  if (!ns || !p) return fail(ADAFLO_B200_ERR_ARG, "null argument");
  if (!ns->has_mode[0]) return fail(ADAFLO_B200_ERR_STATE, "apply_pressure_shift: no constant pressure mode (:215)");
  return vec_axpy(ns->ctx->stream, p, shift, ns->modes[0], ns->mf->spaces[ns->dof_p].n_local_entries());
}

int adaflo_b200_ns_coefficient_ptr(adaflo_b200_ns_t ns, int which, double **dev_ptr, uint64_t *n)
{
  if (ns && ns->generic) return fail(ADAFLO_B200_ERR_UNSUPPORTED, "this entry point is implemented for the 3-D Q2/Q1 operator only (generic-degree path: vmult, residual, velocity_vmult, projection)");
  if (!ns || !dev_ptr) return fail(ADAFLO_B200_ERR_ARG, "null argument");
  const uint64_t size = (uint64_t)n_groups(ns->mf->n_cells) * 270;
  double *p = nullptr;
  uint64_t cnt = size;
  switch (which) {
    case ADAFLO_B200_COEF_DENSITY: p = ns->rho; break;
    case ADAFLO_B200_COEF_VISCOSITY: p = ns->mu; break;
    case ADAFLO_B200_COEF_DAMPING: p = ns->damp; break;
    case ADAFLO_B200_COEF_LINEARIZED_VELOCITY: p = ns->lin; cnt = size * 12; break;
    default: return fail(ADAFLO_B200_ERR_ARG, "coefficient_ptr: unknown coefficient");
  }
  *dev_ptr = p;
  if (n) *n = p ? cnt : 0;
  return 0;
}


static int coef_canonical(adaflo_b200_ns_t ns, int which, double *canonical, int on_host, uint32_t n_batches, uint32_t lanes,
                          const uint32_t *slot_cell_host, bool to_canonical)
{
  if (!ns || !canonical || lanes == 0) return fail(ADAFLO_B200_ERR_ARG, "coefficient_canonical: bad argument");
  double *blocked = nullptr;
  uint64_t cnt = 0;
  AB_CHECK(adaflo_b200_ns_coefficient_ptr(ns, which, &blocked, &cnt));
  if (!blocked) return fail(ADAFLO_B200_ERR_STATE, "coefficient_canonical: array not allocated (constant coefficients)");
  const int n_rows = which == ADAFLO_B200_COEF_LINEARIZED_VELOCITY ? 12 : 1;
  const uint64_t n_slots64 = (uint64_t)n_batches * lanes;
  if (n_slots64 == 0) return 0;
  if (n_slots64 > 0xffffffffull) return fail(ADAFLO_B200_ERR_ARG, "coefficient_canonical: too many slots");
  const uint32_t n_slots = (uint32_t)n_slots64;
  if (!slot_cell_host && n_slots < ns->mf->n_cells) return fail(ADAFLO_B200_ERR_ARG, "coefficient_canonical: fewer slots than cells");
  cudaStream_t s = ns->ctx->stream;
  const size_t n = (size_t)n_slots * 27 * n_rows;
  double *dev = canonical;
  uint32_t *map = nullptr;
  if (on_host) {
    AB_CUDA(cudaMalloc((void **)&dev, n * sizeof(double)));
    if (!to_canonical) AB_CUDA(cudaMemcpyAsync(dev, canonical, n * sizeof(double), cudaMemcpyHostToDevice, s));
    else AB_CUDA(cudaMemsetAsync(dev, 0, n * sizeof(double), s));
  }
  if (slot_cell_host) {
    AB_CUDA(cudaMalloc((void **)&map, (size_t)n_slots * sizeof(uint32_t)));
    AB_CUDA(cudaMemcpyAsync(map, slot_cell_host, (size_t)n_slots * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
  }
  const unsigned grid = (unsigned)(((size_t)n_slots * 27 + 255) / 256);
  if (to_canonical)
    ab200::k_coef_canonical<true><<<grid, 256, 0, s>>>(blocked, dev, map, n_slots, lanes, ns->mf->n_cells, n_rows);
  else
    ab200::k_coef_canonical<false><<<grid, 256, 0, s>>>(blocked, dev, map, n_slots, lanes, ns->mf->n_cells, n_rows);
  AB_COUNT_LAUNCH();
  AB_CUDA(cudaGetLastError());
  if (on_host && to_canonical) AB_CUDA(cudaMemcpyAsync(canonical, dev, n * sizeof(double), cudaMemcpyDeviceToHost, s));
  if (on_host || map) {
    AB_CUDA(cudaStreamSynchronize(s));
    if (on_host) cudaFree(dev);
    if (map) cudaFree(map);
  }
  return 0;
}

int adaflo_b200_ns_set_coefficient_canonical(adaflo_b200_ns_t ns, int which, const double *canonical, int on_host, uint32_t n_batches,
                                             uint32_t lanes, const uint32_t *slot_cell_host)
{
  if (ns && ns->generic) return fail(ADAFLO_B200_ERR_UNSUPPORTED, "this entry point is implemented for the 3-D Q2/Q1 operator only (generic-degree path: vmult, residual, velocity_vmult, projection)");
  return coef_canonical(ns, which, const_cast<double *>(canonical), on_host, n_batches, lanes, slot_cell_host, false);
}

int adaflo_b200_ns_get_coefficient_canonical(adaflo_b200_ns_t ns, int which, double *canonical, int on_host, uint32_t n_batches,
                                             uint32_t lanes, const uint32_t *slot_cell_host)
{
  if (ns && ns->generic) return fail(ADAFLO_B200_ERR_UNSUPPORTED, "this entry point is implemented for the 3-D Q2/Q1 operator only (generic-degree path: vmult, residual, velocity_vmult, projection)");
  return coef_canonical(ns, which, canonical, on_host, n_batches, lanes, slot_cell_host, true);
}

int adaflo_b200_ns_nullspace(adaflo_b200_ns_t ns, int mode, const double **modes, const double **weights, double *inv)
{
  if (!ns || mode < 0 || mode > 1) return fail(ADAFLO_B200_ERR_ARG, "nullspace: bad argument");
  if (modes) *modes = ns->modes[mode];
  if (weights) *weights = ns->mode_weights[mode];
  if (inv) *inv = ns->inv_weights[mode];
  return 0;
}

int adaflo_b200_ns_get_matvec_statistics(adaflo_b200_ns_t ns, uint64_t *count, double *seconds)
{
  if (!ns) return fail(ADAFLO_B200_ERR_ARG, "null argument");
  AB_CUDA(cudaStreamSynchronize(ns->ctx->stream));
  if (count) *count = ns->matvec_count;
  if (seconds) *seconds = ns->matvec_seconds;
  ns->matvec_count = 0;  // :1203-1204
  ns->matvec_seconds = 0;
  return 0;
}

int adaflo_b200_ns_set_kernel_timing(adaflo_b200_ns_t ns, int enable)
{
  if (!ns) return fail(ADAFLO_B200_ERR_ARG, "null argument");
  if (!ns->kernel_events) ns->kernel_events = new std::vector<std::pair<cudaEvent_t, cudaEvent_t>>();
  ns->time_kernel = enable != 0;
  return 0;
}

int adaflo_b200_ns_get_kernel_timing(adaflo_b200_ns_t ns, uint64_t *launches, double *seconds)
{
  if (!ns) return fail(ADAFLO_B200_ERR_ARG, "null argument");
  AB_CUDA(cudaStreamSynchronize(ns->ctx->stream));
  if (ns->kernel_events) {
    for (size_t i = 0; i < ns->kernel_events_used; ++i) {
      float ms = 0;
      AB_CUDA(cudaEventElapsedTime(&ms, (*ns->kernel_events)[i].first, (*ns->kernel_events)[i].second));
      ns->kernel_seconds += 1e-3 * ms;
      ns->kernel_launches++;
    }
    ns->kernel_events_used = 0;
  }
  if (launches) *launches = ns->kernel_launches;
  if (seconds) *seconds = ns->kernel_seconds;
  ns->kernel_launches = 0;
  ns->kernel_seconds = 0;
  return 0;
}

uint64_t adaflo_b200_kernel_launch_count(void) { return (uint64_t)g_launches; }

// Host-pointer entries of the sub-block operators: what the reference's Trilinos wrappers do around them
// (VelocityMatrix::Multiply navier_stokes_preconditioner.cc:332-343, PressurePoissonMatrix::Multiply :421-433: memcpy the
// Epetra vector into a distributed vector, vmult, memcpy back).  Owned entries only; staging vectors are reused.
namespace {
enum { HOST_VELOCITY = 0, HOST_POISSON = 1, HOST_MASS = 2, HOST_CONVDIFF = 3 };
int ensure_stage(adaflo_b200_ns_t ns)
{
  const Space &su = ns->mf->spaces[ns->dof_u], &sp = ns->mf->spaces[ns->dof_p];
  for (int i = 0; i < 2; ++i) {
    if (!ns->stage_u[i]) AB_CUDA(cudaMalloc((void **)&ns->stage_u[i], su.n_local_entries() * sizeof(double)));
    if (!ns->stage_p[i]) AB_CUDA(cudaMalloc((void **)&ns->stage_p[i], sp.n_local_entries() * sizeof(double)));
  }
  return 0;
}
int sub_operator_host(adaflo_b200_ns_t ns, int which, double *dst_host, const double *src_host)
{
  if (!ns) return fail(ADAFLO_B200_ERR_STATE, "host sub-operator: operator not initialised");
  if (!dst_host || !src_host) return fail(ADAFLO_B200_ERR_ARG, "host sub-operator: null vector");
  AB_CHECK(ensure_stage(ns));
  cudaStream_t s = ns->ctx->stream;
  double *src = which == HOST_VELOCITY ? ns->stage_u[0] : ns->stage_p[0], *dst = which == HOST_VELOCITY ? ns->stage_u[1] : ns->stage_p[1];
  const int space = which == HOST_VELOCITY ? ns->dof_u : ns->dof_p;
  AB_CHECK(upload_owned(ns->mf, space, src, src_host, s));
  switch (which) {
    case HOST_VELOCITY: AB_CHECK(adaflo_b200_ns_velocity_vmult(ns, dst, src)); break;
    case HOST_POISSON: AB_CHECK(adaflo_b200_ns_pressure_poisson_vmult(ns, dst, src)); break;
    case HOST_MASS: AB_CHECK(adaflo_b200_ns_pressure_mass_vmult(ns, dst, src)); break;
    default: AB_CHECK(adaflo_b200_ns_pressure_convdiff_vmult(ns, dst, src)); break;
  }
  AB_CHECK(download_owned(ns->mf, space, dst_host, dst, s));
  AB_CUDA(cudaStreamSynchronize(s));
  return 0;
}
}  // namespace

int adaflo_b200_ns_velocity_vmult_host(adaflo_b200_ns_t ns, double *dst_u_host, const double *src_u_host)
{
  return sub_operator_host(ns, HOST_VELOCITY, dst_u_host, src_u_host);
}
int adaflo_b200_ns_pressure_poisson_vmult_host(adaflo_b200_ns_t ns, double *dst_p_host, const double *src_p_host)
{
  return sub_operator_host(ns, HOST_POISSON, dst_p_host, src_p_host);
}
int adaflo_b200_ns_pressure_mass_vmult_host(adaflo_b200_ns_t ns, double *dst_p_host, const double *src_p_host)
{
  return sub_operator_host(ns, HOST_MASS, dst_p_host, src_p_host);
}
int adaflo_b200_ns_pressure_convdiff_vmult_host(adaflo_b200_ns_t ns, double *dst_p_host, const double *src_p_host)
{
  return sub_operator_host(ns, HOST_CONVDIFF, dst_p_host, src_p_host);
}
int adaflo_b200_ns_divergence_vmult_add_host(adaflo_b200_ns_t ns, double *dst_p_host, const double *src_u_host, int weight_by_viscosity)
{
  if (!ns) return fail(ADAFLO_B200_ERR_STATE, "divergence_vmult_add_host: operator not initialised");
  if (!dst_p_host || !src_u_host) return fail(ADAFLO_B200_ERR_ARG, "divergence_vmult_add_host: null vector");
  AB_CHECK(ensure_stage(ns));
  cudaStream_t s = ns->ctx->stream;
  const Space &sp = ns->mf->spaces[ns->dof_p];
  AB_CHECK(upload_owned(ns->mf, ns->dof_u, ns->stage_u[0], src_u_host, s));
  AB_CUDA(cudaMemsetAsync(ns->stage_p[1], 0, sp.n_local_entries() * sizeof(double), s));
  AB_CHECK(upload_owned(ns->mf, ns->dof_p, ns->stage_p[1], dst_p_host, s));  // "add": dst is an input
  AB_CHECK(adaflo_b200_ns_divergence_vmult_add(ns, ns->stage_p[1], ns->stage_u[0], weight_by_viscosity));
  AB_CHECK(download_owned(ns->mf, ns->dof_p, dst_p_host, ns->stage_p[1], s));
  AB_CUDA(cudaStreamSynchronize(s));
  return 0;
}

int adaflo_b200_ns_vmult_host(adaflo_b200_ns_t ns, double *dst_u_host, double *dst_p_host, const double *src_u_host,
                              const double *src_p_host)
{
  if (!ns) return fail(ADAFLO_B200_ERR_STATE, "vmult_host: operator not initialised");
  if (!dst_u_host || !dst_p_host || !src_u_host || !src_p_host) return fail(ADAFLO_B200_ERR_ARG, "vmult_host: null vector");
  cudaStream_t s = ns->ctx->stream;
  const Space &su = ns->mf->spaces[ns->dof_u], &sp = ns->mf->spaces[ns->dof_p];
  const size_t nu = su.n_local_entries(), np = sp.n_local_entries();
  for (int i = 0; i < 2; ++i) {
    if (!ns->stage_u[i]) AB_CUDA(cudaMalloc((void **)&ns->stage_u[i], nu * sizeof(double)));
    if (!ns->stage_p[i]) AB_CUDA(cudaMalloc((void **)&ns->stage_p[i], np * sizeof(double)));
  }
  adaflo_b200_mf_t mf = ns->mf;
  static const bool pipe_enabled = [] { const char *e = std::getenv("ADAFLO_B200_HOST_PIPELINE"); return !(e && !std::strcmp(e, "0")); }();
  static const bool legacy_kernel = [] { const char *k = std::getenv("ADAFLO_B200_VMULT_KERNEL"); return k && !std::strcmp(k, "v1"); }();
  const int mode = mode_of(ns->prm);
  const bool multi = ns->ctx->nranks > 1;
  const bool renumbered = su.numbering || sp.numbering;  // host numbering is not node-major: permute on the device, no chunk pipeline
  if (!pipe_enabled || legacy_kernel || renumbered || ns->generic || !mf->pipeline_ok || ((mode == MODE_NEWTON || mode == MODE_PICARD) && !ns->lin)) {
    // plain path: upload, vmult, download
    AB_CHECK(upload_owned(mf, ns->dof_u, ns->stage_u[0], src_u_host, s));
    AB_CHECK(upload_owned(mf, ns->dof_p, ns->stage_p[0], src_p_host, s));
    AB_CHECK(vmult_impl(ns, ns->stage_u[1], ns->stage_p[1], ns->stage_u[0], ns->stage_p[0], false));
    AB_CHECK(download_owned(mf, ns->dof_u, dst_u_host, ns->stage_u[1], s));
    AB_CHECK(download_owned(mf, ns->dof_p, dst_p_host, ns->stage_p[1], s));
    AB_CUDA(cudaStreamSynchronize(s));
    return 0;
  }
  // Pipelined path: PCIe is full duplex, so the upload of the source prefix needed by cell chunk k+1, the cell
  // kernel on chunk k and the download of the result prefix that chunk k-1 completed run concurrently on three
  // streams.  The prefixes are monotone by construction (mf_create), so any cell numbering is handled; a
  // numbering without locality merely degenerates to upload - compute - download.
  // Several GPUs: the entries this rank exports are uploaded FIRST and the ghost import runs at once; the
  // compress(add) of the result follows the chunk that finishes the cells touching ghosts (the first one for
  // contiguous partitions), and the exported range is held back from the download until it has been added.
  const size_t n_chunks = mf->chunk_decade_end.size();
  if (!ns->up_stream) AB_CUDA(cudaStreamCreateWithFlags(&ns->up_stream, cudaStreamNonBlocking));
  if (!ns->down_stream) AB_CUDA(cudaStreamCreateWithFlags(&ns->down_stream, cudaStreamNonBlocking));
  if (!ns->pipe_events) ns->pipe_events = new std::vector<cudaEvent_t>();
  while (ns->pipe_events->size() < 2 * n_chunks + 1) {
    cudaEvent_t e;
    AB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    ns->pipe_events->push_back(e);
  }
  std::chrono::steady_clock::time_point t0;
  stats_begin(ns, t0);
  double *du = ns->stage_u[1], *dp = ns->stage_p[1];
  const double *xu = ns->stage_u[0], *xp = ns->stage_p[0];
  cudaEvent_t ev_start = (*ns->pipe_events)[2 * n_chunks];
  AB_CUDA(cudaEventRecord(ev_start, s));  // earlier work on the compute stream may still use the staging vectors
  AB_CUDA(cudaStreamWaitEvent(ns->up_stream, ev_start, 0));
  AB_CUDA(cudaStreamWaitEvent(ns->down_stream, ev_start, 0));
  AB_CUDA(cudaMemsetAsync(du, 0, su.n_local_entries() * sizeof(double), s));
  AB_CUDA(cudaMemsetAsync(dp, 0, sp.n_local_entries() * sizeof(double), s));
  NSArgs a = make_args(ns, false);
  a.src_u = xu;
  a.src_p = xp;
  a.dst_u = du;
  a.dst_p = dp;
  size_t up_u = 0, up_p = 0, down_u = 0, n_con = 0;
  bool compressed = !multi;
  // in-order upload of [from, to) that leaves out the export range [skip0, skip1) uploaded up front (multi-GPU)
  cudaStream_t up = ns->up_stream;
  auto upload = [up](double *dev, const double *host, size_t from, size_t to, size_t skip0, size_t skip1) -> int {
    const size_t a1 = std::min(to, std::max(from, skip0)), b0 = std::max(from, std::min(to, skip1));
    if (a1 > from) AB_CUDA(cudaMemcpyAsync(dev + from, host + from, (a1 - from) * sizeof(double), cudaMemcpyHostToDevice, up));
    if (to > b0) AB_CUDA(cudaMemcpyAsync(dev + b0, host + b0, (to - b0) * sizeof(double), cudaMemcpyHostToDevice, up));
    return 0;
  };
  const size_t skip_u0 = multi ? (size_t)su.send_lo * su.n_comp : 0, skip_u1 = multi ? (size_t)su.send_hi * su.n_comp : 0;
  const size_t skip_p0 = multi ? (size_t)sp.send_lo * sp.n_comp : 0, skip_p1 = multi ? (size_t)sp.send_hi * sp.n_comp : 0;
  if (multi) {
    cudaEvent_t ev_s = (*ns->pipe_events)[0];  // (re-recorded as ev_up of chunk 0 below)
    if (su.send_hi > su.send_lo)
      AB_CUDA(cudaMemcpyAsync(ns->stage_u[0] + (size_t)su.send_lo * su.n_comp, src_u_host + (size_t)su.send_lo * su.n_comp,
                              (size_t)(su.send_hi - su.send_lo) * su.n_comp * sizeof(double), cudaMemcpyHostToDevice, ns->up_stream));
    if (sp.send_hi > sp.send_lo)
      AB_CUDA(cudaMemcpyAsync(ns->stage_p[0] + (size_t)sp.send_lo * sp.n_comp, src_p_host + (size_t)sp.send_lo * sp.n_comp,
                              (size_t)(sp.send_hi - sp.send_lo) * sp.n_comp * sizeof(double), cudaMemcpyHostToDevice, ns->up_stream));
    AB_CUDA(cudaEventRecord(ev_s, ns->up_stream));
    AB_CUDA(cudaStreamWaitEvent(s, ev_s, 0));
    const int spaces[2] = {ns->dof_u, ns->dof_p};
    double *vecs[2] = {ns->stage_u[0], ns->stage_p[0]};
    AB_CHECK(comm_update_ghosts_n(mf, 2, spaces, vecs));
  }
  for (size_t k = 0; k < n_chunks; ++k) {
    cudaEvent_t ev_up = (*ns->pipe_events)[2 * k], ev_done = (*ns->pipe_events)[2 * k + 1];
    const size_t need_u = std::min<size_t>((size_t)su.chunk_need_hi[k] * su.n_comp, su.n_owned_entries());
    const size_t need_p = std::min<size_t>((size_t)sp.chunk_need_hi[k] * sp.n_comp, sp.n_owned_entries());
    if (need_u > up_u) AB_CHECK(upload(ns->stage_u[0], src_u_host, up_u, need_u, skip_u0, skip_u1));
    if (need_p > up_p) AB_CHECK(upload(ns->stage_p[0], src_p_host, up_p, need_p, skip_p0, skip_p1));
    up_u = std::max(up_u, need_u);
    up_p = std::max(up_p, need_p);
    if (k + 1 == n_chunks) {  // whatever no cell references (none on conforming meshes)
      if (su.n_owned_entries() > up_u) AB_CHECK(upload(ns->stage_u[0], src_u_host, up_u, su.n_owned_entries(), skip_u0, skip_u1));
      if (sp.n_owned_entries() > up_p) AB_CHECK(upload(ns->stage_p[0], src_p_host, up_p, sp.n_owned_entries(), skip_p0, skip_p1));
    }
    AB_CUDA(cudaEventRecord(ev_up, ns->up_stream));
    AB_CUDA(cudaStreamWaitEvent(s, ev_up, 0));
    a.decade_begin = k ? mf->chunk_decade_end[k - 1] : 0;
    a.decade_end = mf->chunk_decade_end[k];
    AB_CHECK(launch_cells(ns, a, true));
    if (!compressed && (mf->chunk_decade_end[k] >= mf->ghost_decade_end || k + 1 == n_chunks)) {
      const int spaces[2] = {ns->dof_u, ns->dof_p};  // all cells touching ghosts are done: export their contributions
      double *vecs[2] = {du, dp};
      AB_CHECK(comm_compress_add_n(mf, 2, spaces, vecs));
      compressed = true;
    }
    // constrained rows of the velocity prefix that is final now (:247-256), then its download
    size_t fin_u = k + 1 == n_chunks ? su.n_owned_entries() : std::min<size_t>((size_t)su.chunk_final_lo[k] * su.n_comp, up_u);
    if (!compressed) fin_u = std::min<size_t>(fin_u, (size_t)su.send_lo * su.n_comp);  // entries neighbours still add to
    size_t con_end = k + 1 == n_chunks ? su.n_constrained : su.chunk_constrained_end[k];
    if (k + 1 < n_chunks && (size_t)su.chunk_final_lo[k] * su.n_comp > up_u) con_end = n_con;  // (orphan nodes: wait for the last chunk)
    if (con_end > n_con) AB_CHECK(vec_copy_indexed(s, du, xu, su.constrained + n_con, (uint32_t)(con_end - n_con), 1.0));
    n_con = std::max(n_con, con_end);
    AB_CUDA(cudaEventRecord(ev_done, s));
    if (fin_u > down_u) {
      AB_CUDA(cudaStreamWaitEvent(ns->down_stream, ev_done, 0));
      AB_CUDA(cudaMemcpyAsync(dst_u_host + down_u, du + down_u, (fin_u - down_u) * sizeof(double), cudaMemcpyDeviceToHost, ns->down_stream));
      down_u = fin_u;
    }
  }
  // pressure block: constrained rows and the global null-space projection need the whole vector
  AB_CHECK(vec_copy_indexed(s, dp, xp, sp.constrained, sp.n_constrained, -1.0));
  if (projection_active(ns))
    for (int m = 0; m < 2; ++m)
      if (ns->has_mode[m])
        AB_CHECK(vec_project(ns->ctx, dp, ns->mode_weights[m], ns->modes[m], ns->inv_weights[m], sp.n_owned_entries(), sp.n_local_entries()));
  AB_CUDA(cudaMemcpyAsync(dst_p_host, dp, sp.n_owned_entries() * sizeof(double), cudaMemcpyDeviceToHost, s));
  stats_end(ns, t0);
  AB_CHECK(stream_wait(ns->ctx, ns->down_stream));
  AB_CHECK(stream_wait(ns->ctx, s));
  return 0;
}

// ------------------------------------------------- Krylov-step vector kernels
int adaflo_b200_vec_alloc(adaflo_b200_ctx_t ctx, uint64_t n, double **dev_out)
{
  if (!ctx || !dev_out) return fail(ADAFLO_B200_ERR_ARG, "vec_alloc: null argument");
  AB_CUDA(cudaSetDevice(ctx->device));
  *dev_out = nullptr;
  if (n == 0) return 0;
  AB_CUDA(cudaMalloc((void **)dev_out, n * sizeof(double)));
  AB_CUDA(cudaMemsetAsync(*dev_out, 0, n * sizeof(double), ctx->stream));
  return 0;
}
int adaflo_b200_vec_free(adaflo_b200_ctx_t ctx, double *dev)
{
  if (!ctx) return fail(ADAFLO_B200_ERR_ARG, "vec_free: null context");
  if (dev) {
    AB_CUDA(cudaStreamSynchronize(ctx->stream));
    AB_CUDA(cudaFree(dev));
  }
  return 0;
}
int adaflo_b200_vec_upload(adaflo_b200_ctx_t ctx, double *dev, const double *host, uint64_t n)
{
  if (!ctx || (n && (!dev || !host))) return fail(ADAFLO_B200_ERR_ARG, "vec_upload: null argument");
  AB_CUDA(cudaMemcpyAsync(dev, host, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  AB_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}
int adaflo_b200_vec_download(adaflo_b200_ctx_t ctx, double *host, const double *dev, uint64_t n)
{
  if (!ctx || (n && (!dev || !host))) return fail(ADAFLO_B200_ERR_ARG, "vec_download: null argument");
  AB_CUDA(cudaMemcpyAsync(host, dev, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  AB_CHECK(stream_wait(ctx, ctx->stream));
  return 0;
}
int adaflo_b200_vec_set(adaflo_b200_ctx_t ctx, double *x, double v, uint64_t n) { return vec_set(ctx->stream, x, v, n); }
int adaflo_b200_vec_axpy(adaflo_b200_ctx_t ctx, double *y, double a, const double *x, uint64_t n) { return vec_axpy(ctx->stream, y, a, x, n); }
int adaflo_b200_vec_sadd(adaflo_b200_ctx_t ctx, double *y, double sc, double a, const double *x, uint64_t n)
{
  return vec_sadd(ctx->stream, y, sc, a, x, n);
}
int adaflo_b200_vec_scale(adaflo_b200_ctx_t ctx, double *y, const double *sc, uint64_t n) { return vec_scale(ctx->stream, y, sc, n); }
int adaflo_b200_vec_dot(adaflo_b200_ctx_t ctx, const double *x, const double *y, uint64_t n, double *result_host)
{
  const double *ys[1] = {y};
  return vec_multi_dot(ctx, x, ys, 1, n, result_host, true);
}
int adaflo_b200_vec_multi_dot(adaflo_b200_ctx_t ctx, const double *x, const double *const *ys, int k, uint64_t n, double *results_host)
{
  return vec_multi_dot(ctx, x, ys, k, n, results_host, true);
}
int adaflo_b200_vec_multi_axpy(adaflo_b200_ctx_t ctx, double *y, const double *a_host, const double *const *xs, int k, uint64_t n)
{
  if (!ctx || (k > 0 && n && (!y || !a_host || !xs))) return fail(ADAFLO_B200_ERR_ARG, "vec_multi_axpy: null argument");
  return vec_multi_axpy(ctx, y, a_host, xs, k, n);
}
int adaflo_b200_vec_add_and_dot(adaflo_b200_ctx_t ctx, double *y, double a, const double *x, const double *z, uint64_t n, double *result_host)
{
  return vec_add_and_dot(ctx, y, a, x, z, n, result_host);
}
int adaflo_b200_jacobi_reinit(adaflo_b200_ctx_t ctx, double *inv_diag, const double *diag, uint64_t n) { return jacobi_reinit(ctx, inv_diag, diag, n); }
int adaflo_b200_jacobi_vmult(adaflo_b200_ctx_t ctx, double *dst, const double *src, const double *inv_diag, uint64_t n, int n_blocks)
{
  return jacobi_vmult(ctx->stream, dst, src, inv_diag, n, n_blocks);
}

}  // extern "C"
