// ns_generic.cu -- degree / dimension dispatch of the Navier-Stokes operator (EXPAND_OPERATIONS, source/navier_stokes_matrix.cc:
// 64-82: the reference instantiates local_operation for pressure degrees 1..5 and NavierStokesMatrix for dim 1, 2, 3, :1210-1212).
//
// The tuned kernels (ns_r1.cuh, ns_residual_r1.cu) are specialised for 3-D Taylor-Hood Q2/Q1.  Every other pair handled here --
// velocity degree K = 2..5 with pressure degree K - 1, dim = 2 or 3, e.g. tests/beltrami_2d.prm (Q4/Q3, dim 2) -- goes through one
// templated sum-factorisation kernel: one thread block per cell, one thread per quadrature point, 1-D shape matrices as
// kernel arguments, all tensor contractions as sweeps over shared-memory arrays:
//   gather -> values at the (K+1)^dim Gauss points (dim sweeps per component) -> reference gradients by the collocation
//   derivative in the Gauss-point Lagrange basis (1 sweep per direction: K + 1 Gauss values determine the degree-K polynomial)
//   -> per-point physics (source/navier_stokes_matrix.cc:702-893, the formulas of ns_r1.cuh for any dim) -> transposed sweeps
//   -> red.global.add scatter.
// FE_Q(K) of deal.II is the Lagrange basis on the Gauss-Lobatto points (equidistant for K <= 2); cell-local nodes are
// lexicographic, x fastest.  Conforming meshes or general constraint rows (hanging nodes: gc_read / gc_add), constant coefficients,
// Cartesian / affine / general geometry.  This is the
// correctness path for the configurations outside the benchmark (<= 10^4 cells in the reference's tests), not a tuned one.
#include <cmath>
#include <vector>

#include "common.cuh"
#include "ns_q2q1.cuh"

namespace ab200 {

constexpr int kGenMaxN1 = 6;  // velocity degree <= 5

struct GenTables {
  double N[kGenMaxN1 * kGenMaxN1];    // [q][i] velocity shape value at Gauss point q
  double Dq[kGenMaxN1 * kGenMaxN1];   // [q][j] derivative of the Gauss-point Lagrange polynomial j at Gauss point q
  double NP[kGenMaxN1 * kGenMaxN1];   // [q][i] pressure shape value (degree K - 1: K nodes) at Gauss point q
  double W[kGenMaxN1];                // Gauss weights on [0, 1]
};

template <int DIM>
__device__ __forceinline__ int ipow(int b)
{
  return DIM == 2 ? b * b : b * b * b;
}

// out = (M applied along direction dir) in.  M is [n_out][n_in] (TRANSPOSE: M^T, i.e. M is [n_in][n_out]); the array has extent
// ext[d] in the other directions and n_in (input) / n_out (output) in direction dir.  ACC: out += ...
template <int DIM, bool TRANSPOSE, bool ACC>
__device__ void sweep(const double *M, const int m_stride, const int n_out, const int n_in, const int dir, const int (&ext)[3], const double *in,
                      double *out)
{
  int eo[3] = {ext[0], ext[1], DIM == 3 ? ext[2] : 1}, ei[3] = {ext[0], ext[1], DIM == 3 ? ext[2] : 1};
  eo[dir] = n_out;
  ei[dir] = n_in;
  const int total = eo[0] * eo[1] * eo[2];
  const int si[3] = {1, ei[0], ei[0] * ei[1]};
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    int c[3] = {idx % eo[0], (idx / eo[0]) % eo[1], idx / (eo[0] * eo[1])};
    const int o = c[dir];
    c[dir] = 0;
    const double *p = in + c[0] * si[0] + c[1] * si[1] + c[2] * si[2];
    double s = ACC ? out[idx] : 0.0;
    for (int i = 0; i < n_in; ++i) s = fma(TRANSPOSE ? M[i * m_stride + o] : M[o * m_stride + i], p[i * si[dir]], s);
    out[idx] = s;
  }
  __syncthreads();
}

// General constraint rows (hanging nodes of any degree; adaflo_b200_space_desc::hanging_*): what FEEvaluation::read_dof_values and
// distribute_local_to_global do with an AffineConstraints row.  A flagged (homogeneous Dirichlet) component of a parent is read as
// zero and never written.
template <int NC>
__device__ __forceinline__ double gc_read(const double *vec, const uint32_t row, const int c, const uint32_t *ptr, const uint32_t *cols, const double *w)
{
  double v = 0.0;
  for (uint32_t k = __ldg(ptr + row), end = __ldg(ptr + row + 1); k < end; ++k) {
    const uint32_t pe = __ldg(cols + k);
    if (!((pe >> (ADAFLO_B200_FLAG_SHIFT + c)) & 1u)) v = fma(__ldg(w + k), vec[(size_t)NC * (pe & ADAFLO_B200_NODE_MASK) + c], v);
  }
  return v;
}
template <int NC>
__device__ __forceinline__ void gc_add(double *vec, const uint32_t row, const int c, const uint32_t *ptr, const uint32_t *cols, const double *w, const double val)
{
  for (uint32_t k = __ldg(ptr + row), end = __ldg(ptr + row + 1); k < end; ++k) {
    const uint32_t pe = __ldg(cols + k);
    if (!((pe >> (ADAFLO_B200_FLAG_SHIFT + c)) & 1u)) red_add(vec + (size_t)NC * (pe & ADAFLO_B200_NODE_MASK) + c, __ldg(w + k) * val);
  }
}

enum { GEN_VMULT = 0, GEN_RESIDUAL = 1, GEN_PRESSURE_MASS = 2, GEN_NODAL_MASS = 3, GEN_DIVERGENCE = 4 };
// GEN_DIVERGENCE: local_divergence (navier_stokes_matrix.cc:920-962): dst_p += int q (weight div u), velocity quadrature.
// GEN_NODAL_MASS: integrals of the degree-K shape functions with QGauss(K + 1), scattered through cell_p into dst_p.  Instantiated
// with K = pressure degree it gives the integrals of the pressure shape functions with the PRESSURE quadrature QGauss(degree_p + 1),
// which is what the reference uses for the null-space weights (navier_stokes_matrix.cc:110-168).

// Cells per block: the block is (points of a cell) x (cells), about 128 threads, so that the small cells of the 2-D / low-degree
// configurations fill warps (2-D Q2: 14 cells x 9 points instead of one 9-thread block per cell).  All barriers of the kernel sit
// in block-uniform control flow (conditions on kernel arguments only), so the cells of a block simply run in lock step.
template <int NQ>
struct GenCellsPerBlock {
  static constexpr int value = NQ >= 64 ? 1 : 128 / NQ;  // measured: two 64-point cells per block are slower than one (3-D Q3: 9.6 vs 11.0 GDoF/s)
};

template <int DIM, int K, int KIND>
__global__ void ns_generic_kernel(const NSArgs A, const GenTables T, const int with_p, const int mode, const int geom)
{
  constexpr int n1 = K + 1, NN = DIM == 2 ? n1 * n1 : n1 * n1 * n1, NQ = NN, NPN = DIM == 2 ? K * K : K * K * K;
  constexpr int NROWS = DIM + DIM * DIM;
  constexpr int CPB = GenCellsPerBlock<NQ>::value;
  __shared__ double tabN[n1 * n1], tabD[n1 * n1], tabP[n1 * K];
  __shared__ double work_[CPB][2][NN];          // ping-pong arrays of the sweeps
  __shared__ double val_[CPB][DIM][NQ];         // values at the points, then the value test integrands
  __shared__ double grad_[CPB][DIM][DIM][NQ];   // reference gradients [component][direction], then the gradient test integrands
  __shared__ double hist_[CPB][DIM][NQ];        // residual: w_old u_old + w_oo u_old_old at the points
  __shared__ double pq_[CPB][NQ];               // pressure at the points, then the pressure test integrand
  const int slot = threadIdx.y;
  double(*work)[NN] = work_[slot];
  double(*val)[NQ] = val_[slot];
  double(*grad)[DIM][NQ] = grad_[slot];
  double(*hist)[NQ] = hist_[slot];
  double *pq = pq_[slot];
  // a block's last slots may lie beyond the mesh: they follow the barriers on the last cell's data and write nothing
  const bool active = (size_t)blockIdx.x * CPB + slot < A.n_cells;
  const uint32_t cell = active ? blockIdx.x * CPB + slot : A.n_cells - 1;
  const int tid = threadIdx.x;
  constexpr bool RESOLVE = KIND == GEN_VMULT || KIND == GEN_DIVERGENCE;  // read_dof_values; the residual reads plainly
  // compute_diagonal: unit source in cell-local slot (node dslot % NN, component dslot / NN), only that slot's row is scattered
  const int dslot = KIND == GEN_VMULT ? A.diag - 1 : -1;
  for (int i = tid; i < n1 * n1; i += blockDim.x) {
    tabN[i] = T.N[(i / n1) * kGenMaxN1 + i % n1];
    tabD[i] = T.Dq[(i / n1) * kGenMaxN1 + i % n1];
  }
  for (int i = tid; i < n1 * K; i += blockDim.x) tabP[i] = T.NP[(i / K) * kGenMaxN1 + i % K];
  const uint32_t *cu = A.cell_u + (size_t)cell * NN;
  const uint32_t *cp = A.cell_p + (size_t)cell * NPN;
  const int ext_q[3] = {n1, n1, n1};
  __syncthreads();

  if (KIND != GEN_PRESSURE_MASS && KIND != GEN_NODAL_MASS) {
    // ---- velocity: gather, interpolate to the points, collocation gradients (also the old solutions' values for the residual)
    const int n_fields = KIND == GEN_RESIDUAL ? 3 : 1;
    for (int fld = 0; fld < n_fields; ++fld) {
      const double *vec = fld == 0 ? A.src_u : (fld == 1 ? A.old_u : A.old_old_u);
      const double wgt = fld == 1 ? A.w_old : A.w_oo;
      if (fld > 0 && (!vec || A.physical_type != ADAFLO_B200_INCOMPRESSIBLE)) continue;  // block-uniform
      for (int c = 0; c < DIM; ++c) {
        for (int i = tid; i < NN; i += blockDim.x) {
          const uint32_t e = __ldg(cu + i), node = e & ADAFLO_B200_NODE_MASK;
          // vmult resolves the constraints (read_dof_values), the residual reads plainly (read_dof_values_plain, :663-664:
          // also the entry of a hanging node, which AffineConstraints::distribute has filled, navier_stokes.cc:1252)
          if (dslot >= 0)
            work[0][i] = (i == dslot % NN && c == dslot / NN) ? 1.0 : 0.0;
          else if (RESOLVE && node >= A.gc_first_u)
            work[0][i] = gc_read<DIM>(vec, node - A.gc_first_u, c, A.gc_ptr_u, A.gc_cols_u, A.gc_w_u);
          else
            work[0][i] = (RESOLVE && ((e >> (ADAFLO_B200_FLAG_SHIFT + c)) & 1u)) ? 0.0 : vec[(size_t)DIM * node + c];
        }
        __syncthreads();
        int cur = 0;
        for (int d = 0; d < DIM; ++d) {
          sweep<DIM, false, false>(tabN, n1, n1, n1, d, ext_q, work[cur], work[cur ^ 1]);
          cur ^= 1;
        }
        if (fld == 0) {
          for (int i = tid; i < NQ; i += blockDim.x) val[c][i] = work[cur][i];
          __syncthreads();
          for (int d = 0; d < DIM; ++d) sweep<DIM, false, false>(tabD, n1, n1, n1, d, ext_q, val[c], grad[c][d]);
        } else {
          for (int i = tid; i < NQ; i += blockDim.x) hist[c][i] = (fld == 1 ? 0.0 : hist[c][i]) + wgt * work[cur][i];
          __syncthreads();
        }
      }
    }
    // ---- pressure at the points
    if (with_p) {
      for (int i = tid; i < NPN; i += blockDim.x) {
        const uint32_t e = __ldg(cp + i), node = e & ADAFLO_B200_NODE_MASK;
        if (RESOLVE && node >= A.gc_first_p)
          work[0][i] = gc_read<1>(A.src_p, node - A.gc_first_p, 0, A.gc_ptr_p, A.gc_cols_p, A.gc_w_p);
        else
          work[0][i] = (RESOLVE && (e >> ADAFLO_B200_FLAG_SHIFT)) ? 0.0 : A.src_p[node];
      }
      __syncthreads();
      int cur = 0, ext[3] = {K, K, K};
      for (int d = 0; d < DIM; ++d) {
        sweep<DIM, false, false>(tabP, K, n1, K, d, ext, work[cur], work[cur ^ 1]);
        ext[d] = n1;
        cur ^= 1;
      }
      for (int i = tid; i < NQ; i += blockDim.x) pq[i] = work[cur][i];
      __syncthreads();
    }
  }

  // ---- per-point physics: thread q
  if (tid < NQ) {
    const int q = tid;
    int qi[3] = {q % n1, (q / n1) % n1, DIM == 3 ? q / (n1 * n1) : 0};
    double Kj[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}, jxw;
    const double wq = T.W[qi[0]] * T.W[qi[1]] * (DIM == 3 ? T.W[qi[2]] : 1.0);
    if (geom == ADAFLO_B200_GEOM_CARTESIAN) {
      const double *row = A.geom_table + 4 * (size_t)__ldg(A.geom_row + cell);
      for (int f = 0; f < DIM; ++f) Kj[f][f] = row[f];
      jxw = row[3] * wq;
    } else if (geom == ADAFLO_B200_GEOM_AFFINE) {
      const double *row = A.geom_table + 10 * (size_t)__ldg(A.geom_row + cell);
      for (int f = 0; f < DIM; ++f)
        for (int e = 0; e < DIM; ++e) Kj[f][e] = row[3 * f + e];
      jxw = row[9] * wq;
    } else if (geom == 3) {  // per-point data of QGauss(3) in 3-D: stored decade blocked for the tuned kernels (coef_index)
      for (int f = 0; f < DIM; ++f)
        for (int e = 0; e < DIM; ++e) Kj[f][e] = __ldg(A.inv_jac + coef_index(cell, 3 * f + e, 9, qi[0], qi[1], qi[2]));
      jxw = __ldg(A.jxw + coef_index(cell, 0, 1, qi[0], qi[1], qi[2]));
    } else {
      for (int f = 0; f < DIM; ++f)
        for (int e = 0; e < DIM; ++e) Kj[f][e] = __ldg(A.inv_jac + ((size_t)cell * 9 + 3 * f + e) * NQ + q);
      jxw = __ldg(A.jxw + (size_t)cell * NQ + q);
    }
    if (KIND == GEN_PRESSURE_MASS || KIND == GEN_NODAL_MASS) {
      pq[q] = jxw;  // integral of every pressure shape function
    } else {
      double u[DIM], g[DIM][DIM];
      for (int d = 0; d < DIM; ++d) {
        u[d] = val[d][q];
        for (int e = 0; e < DIM; ++e) {
          double s = 0;
          for (int f = 0; f < DIM; ++f) s = fma(grad[d][f][q], Kj[f][e], s);
          g[d][e] = s;
        }
      }
      double div = 0;
      for (int d = 0; d < DIM; ++d) div += g[d][d];
      if (KIND == GEN_DIVERGENCE) {
        pq[q] = A.div_weight * div * jxw;
      } else {
      double conv[DIM];
      for (int d = 0; d < DIM; ++d) conv[d] = 0;
      double *lin = A.lin ? A.lin + (size_t)cell * NROWS * NQ + q : nullptr;  // row r at lin[r * NQ]
      const bool stokes = A.physical_type == ADAFLO_B200_STOKES;
      if (KIND == GEN_VMULT) {
        if (mode != MODE_STOKES) {
          for (int d = 0; d < DIM; ++d) conv[d] = A.gamma * u[d];
          if (mode == MODE_NEWTON) {
            double lu[DIM], lg[DIM][DIM], ltr = 0;
            for (int d = 0; d < DIM; ++d) {
              lu[d] = lin[d * NQ];
              for (int e = 0; e < DIM; ++e) lg[d][e] = lin[(DIM + DIM * d + e) * NQ];
              ltr += lg[d][d];
            }
            const double f1 = A.beta * div, f2 = A.beta * ltr;
            for (int d = 0; d < DIM; ++d) {
              double res = f1 * lu[d] + f2 * u[d];
              for (int e = 0; e < DIM; ++e) res += lu[e] * g[d][e] + u[e] * lg[d][e];
              conv[d] += A.tau1 * res;
            }
          } else if (mode == MODE_PICARD) {
            double lu[DIM];
            for (int d = 0; d < DIM; ++d) lu[d] = lin[d * NQ];
            const double ldiv = lin[DIM * NQ];
            for (int d = 0; d < DIM; ++d) {
              double res = A.beta * ldiv * u[d];
              for (int e = 0; e < DIM; ++e) res += lu[e] * g[d][e];
              conv[d] += A.tau1 * res;
            }
          }
          for (int d = 0; d < DIM; ++d) conv[d] = (conv[d] * A.rho0 - A.damp0 * u[d]) * jxw;
        }
      } else if (!stokes) {  // residual, navier_stokes_matrix.cc:725-800
        for (int d = 0; d < DIM; ++d) {
          double res = A.beta * div * u[d];
          for (int e = 0; e < DIM; ++e) res += u[e] * g[d][e];
          conv[d] = A.gamma * u[d] + (A.physical_type == ADAFLO_B200_INCOMPRESSIBLE ? hist[d][q] : 0.0) + A.tau1 * res;
        }
        if (lin && active) {
          for (int d = 0; d < DIM; ++d) lin[d * NQ] = u[d];
          if (A.linearization == ADAFLO_B200_COUPLED_IMPLICIT_NEWTON) {
            for (int d = 0; d < DIM; ++d)
              for (int e = 0; e < DIM; ++e) lin[(DIM + DIM * d + e) * NQ] = g[d][e];
          } else
            lin[DIM * NQ] = div;
        }
        for (int d = 0; d < DIM; ++d) conv[d] = (conv[d] * A.rho0 - A.damp0 * u[d]) * jxw;
      }
      const double tmu = A.mu0 * A.tau1;
      const double pres = with_p ? pq[q] : 0.0;
      for (int d = 0; d < DIM; ++d) {
        double S[DIM];
        for (int e = 0; e < DIM; ++e) S[e] = tmu * (g[d][e] + g[e][d]);
        S[d] += A.tau_gd * div - pres;
        val[d][q] = conv[d];
        for (int f = 0; f < DIM; ++f) {
          double s = 0;
          for (int e = 0; e < DIM; ++e) s = fma(Kj[f][e], S[e], s);
          grad[d][f][q] = s * jxw;
        }
      }
      pq[q] = -div * jxw;
      }  // KIND != GEN_DIVERGENCE
    }
  }
  __syncthreads();

  // ---- integrate and scatter
  if (KIND != GEN_PRESSURE_MASS && KIND != GEN_NODAL_MASS && KIND != GEN_DIVERGENCE) {
    for (int c = 0; c < DIM; ++c) {
      for (int d = 0; d < DIM; ++d) sweep<DIM, true, true>(tabD, n1, n1, n1, d, ext_q, grad[c][d], val[c]);  // val += Dq^T grad_d
      int cur = 0;
      for (int i = tid; i < NQ; i += blockDim.x) work[0][i] = val[c][i];
      __syncthreads();
      for (int d = 0; d < DIM; ++d) {
        sweep<DIM, true, false>(tabN, n1, n1, n1, d, ext_q, work[cur], work[cur ^ 1]);
        cur ^= 1;
      }
      for (int i = tid; i < NN && active; i += blockDim.x) {
        const uint32_t e = __ldg(cu + i), node = e & ADAFLO_B200_NODE_MASK;
        if (dslot >= 0) {
          if (i != dslot % NN || c != dslot / NN) continue;
          if (node >= A.gc_first_u) {  // a constrained slot adds w_k^2 a_ii to parent k (couplings between two slots of one parent are left out)
            for (uint32_t k = __ldg(A.gc_ptr_u + node - A.gc_first_u), end = __ldg(A.gc_ptr_u + node - A.gc_first_u + 1); k < end; ++k) {
              const uint32_t pe = __ldg(A.gc_cols_u + k);
              const double w = __ldg(A.gc_w_u + k);
              if (!((pe >> (ADAFLO_B200_FLAG_SHIFT + c)) & 1u)) red_add(A.dst_u + (size_t)DIM * (pe & ADAFLO_B200_NODE_MASK) + c, w * w * work[cur][i]);
            }
          } else if (!((e >> (ADAFLO_B200_FLAG_SHIFT + c)) & 1u))
            red_add(A.dst_u + (size_t)DIM * node + c, work[cur][i]);
          continue;
        }
        if (node >= A.gc_first_u)
          gc_add<DIM>(A.dst_u, node - A.gc_first_u, c, A.gc_ptr_u, A.gc_cols_u, A.gc_w_u, work[cur][i]);
        else if (!((e >> (ADAFLO_B200_FLAG_SHIFT + c)) & 1u))
          red_add(A.dst_u + (size_t)DIM * node + c, work[cur][i]);
      }
      __syncthreads();
    }
  }
  if (KIND == GEN_NODAL_MASS) {
    int cur = 0;
    for (int i = tid; i < NQ; i += blockDim.x) work[0][i] = pq[i];
    __syncthreads();
    for (int d = 0; d < DIM; ++d) {
      sweep<DIM, true, false>(tabN, n1, n1, n1, d, ext_q, work[cur], work[cur ^ 1]);
      cur ^= 1;
    }
    for (int i = tid; i < NN && active; i += blockDim.x) {
      const uint32_t node = __ldg(A.cell_p + (size_t)cell * NN + i) & ADAFLO_B200_NODE_MASK;
      if (node >= A.gc_first_p)
        gc_add<1>(A.dst_p, node - A.gc_first_p, 0, A.gc_ptr_p, A.gc_cols_p, A.gc_w_p, work[cur][i]);
      else
        red_add(A.dst_p + node, work[cur][i]);
    }
  } else if (KIND == GEN_PRESSURE_MASS || KIND == GEN_DIVERGENCE || (with_p && !A.skip_p_test && dslot < 0)) {
    int cur = 0, ext[3] = {n1, n1, n1};
    for (int i = tid; i < NQ; i += blockDim.x) work[0][i] = pq[i];
    __syncthreads();
    for (int d = 0; d < DIM; ++d) {
      sweep<DIM, true, false>(tabP, K, K, n1, d, ext, work[cur], work[cur ^ 1]);
      ext[d] = K;
      cur ^= 1;
    }
    for (int i = tid; i < NPN && active; i += blockDim.x) {
      const uint32_t e = __ldg(cp + i), node = e & ADAFLO_B200_NODE_MASK;
      if (node >= A.gc_first_p)
        gc_add<1>(A.dst_p, node - A.gc_first_p, 0, A.gc_ptr_p, A.gc_cols_p, A.gc_w_p, work[cur][i]);
      else if (KIND == GEN_PRESSURE_MASS || !(e >> ADAFLO_B200_FLAG_SHIFT))
        red_add(A.dst_p + node, work[cur][i]);
    }
  }
}

// ---------------------------------------------------------------- scalar operators on the pressure space
// local_pressure_mass / local_pressure_poisson with constant coefficients (navier_stokes_matrix.cc:966-1073): a scalar field of
// degree K with QGauss(K + 1) -- for the pressure of Q_(K+1)/Q_K this is the reference's pressure quadrature quad_index_p.
//   SC_MASS:    dst += int phi (c0 p);   SC_POISSON: dst += int grad phi . (c0 grad p)
enum { SC_MASS = 0, SC_POISSON = 1 };
template <int DIM, int K, int KIND>
__global__ void gen_scalar_kernel(const NSArgs A, const GenTables T, const int geom)
{
  constexpr int n1 = K + 1, NN = DIM == 2 ? n1 * n1 : n1 * n1 * n1, NQ = NN;
  __shared__ double tabN[n1 * n1], tabD[n1 * n1];
  __shared__ double work[2][NN], val[NQ], grad[DIM][NQ];
  const uint32_t cell = blockIdx.x;
  const int tid = threadIdx.x;
  for (int i = tid; i < n1 * n1; i += blockDim.x) {
    tabN[i] = T.N[(i / n1) * kGenMaxN1 + i % n1];
    tabD[i] = T.Dq[(i / n1) * kGenMaxN1 + i % n1];
  }
  const uint32_t *cp = A.cell_p + (size_t)cell * NN;
  const int ext_q[3] = {n1, n1, n1};
  for (int i = tid; i < NN; i += blockDim.x) {
    const uint32_t e = __ldg(cp + i), node = e & ADAFLO_B200_NODE_MASK;
    if (node >= A.gc_first_p)
      work[0][i] = gc_read<1>(A.src_p, node - A.gc_first_p, 0, A.gc_ptr_p, A.gc_cols_p, A.gc_w_p);
    else
      work[0][i] = (e >> ADAFLO_B200_FLAG_SHIFT) ? 0.0 : A.src_p[node];
  }
  __syncthreads();
  int cur = 0;
  for (int d = 0; d < DIM; ++d) {
    sweep<DIM, false, false>(tabN, n1, n1, n1, d, ext_q, work[cur], work[cur ^ 1]);
    cur ^= 1;
  }
  for (int i = tid; i < NQ; i += blockDim.x) val[i] = work[cur][i];
  __syncthreads();
  if (KIND == SC_POISSON)
    for (int d = 0; d < DIM; ++d) sweep<DIM, false, false>(tabD, n1, n1, n1, d, ext_q, val, grad[d]);
  if (tid < NQ) {
    const int q = tid;
    const int qi[3] = {q % n1, (q / n1) % n1, DIM == 3 ? q / (n1 * n1) : 0};
    double Kj[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}, jxw;
    const double wq = T.W[qi[0]] * T.W[qi[1]] * (DIM == 3 ? T.W[qi[2]] : 1.0);
    if (geom == ADAFLO_B200_GEOM_CARTESIAN) {
      const double *row = A.geom_table + 4 * (size_t)__ldg(A.geom_row + cell);
      for (int f = 0; f < DIM; ++f) Kj[f][f] = row[f];
      jxw = row[3] * wq;
    } else if (geom == ADAFLO_B200_GEOM_AFFINE) {
      const double *row = A.geom_table + 10 * (size_t)__ldg(A.geom_row + cell);
      for (int f = 0; f < DIM; ++f)
        for (int e = 0; e < DIM; ++e) Kj[f][e] = row[3 * f + e];
      jxw = row[9] * wq;
    } else {
      for (int f = 0; f < DIM; ++f)
        for (int e = 0; e < DIM; ++e) Kj[f][e] = __ldg(A.inv_jac + ((size_t)cell * 9 + 3 * f + e) * NQ + q);
      jxw = __ldg(A.jxw + (size_t)cell * NQ + q);
    }
    if (KIND == SC_MASS)
      val[q] *= A.scalar_c0 * jxw;
    else {
      double g[DIM];
      for (int e = 0; e < DIM; ++e) {
        double s = 0;
        for (int f = 0; f < DIM; ++f) s = fma(grad[f][q], Kj[f][e], s);
        g[e] = s * A.scalar_c0 * jxw;
      }
      for (int f = 0; f < DIM; ++f) {
        double s = 0;
        for (int e = 0; e < DIM; ++e) s = fma(Kj[f][e], g[e], s);
        grad[f][q] = s;
      }
      val[q] = 0.0;
    }
  }
  __syncthreads();
  if (KIND == SC_POISSON)
    for (int d = 0; d < DIM; ++d) sweep<DIM, true, true>(tabD, n1, n1, n1, d, ext_q, grad[d], val);
  cur = 0;
  for (int i = tid; i < NQ; i += blockDim.x) work[0][i] = val[i];
  __syncthreads();
  for (int d = 0; d < DIM; ++d) {
    sweep<DIM, true, false>(tabN, n1, n1, n1, d, ext_q, work[cur], work[cur ^ 1]);
    cur ^= 1;
  }
  for (int i = tid; i < NN; i += blockDim.x) {
    const uint32_t e = __ldg(cp + i), node = e & ADAFLO_B200_NODE_MASK;
    if (node >= A.gc_first_p)
      gc_add<1>(A.dst_p, node - A.gc_first_p, 0, A.gc_ptr_p, A.gc_cols_p, A.gc_w_p, work[cur][i]);
    else if (!(e >> ADAFLO_B200_FLAG_SHIFT))
      red_add(A.dst_p + node, work[cur][i]);
  }
}

// ---------------------------------------------------------------- host: 1-D shape data in long double
namespace {
// Legendre polynomial P_n and derivative at x in [-1, 1]
void legendre(int n, long double x, long double &p, long double &dp)
{
  long double p0 = 1, p1 = x;
  if (n == 0) { p = 1; dp = 0; return; }
  for (int k = 2; k <= n; ++k) {
    const long double pk = ((2 * k - 1) * x * p1 - (k - 1) * p0) / k;
    p0 = p1;
    p1 = pk;
  }
  p = p1;
  dp = n * (x * p1 - p0) / (x * x - 1);
}
std::vector<long double> gauss_points(int n, std::vector<long double> &w)  // on [0, 1]
{
  std::vector<long double> x(n);
  w.resize(n);
  const long double pi = acosl(-1.0L);
  for (int i = 0; i < n; ++i) {
    long double z = -cosl(pi * (i + 0.75L) / (n + 0.5L)), p, dp;
    for (int it = 0; it < 100; ++it) {
      legendre(n, z, p, dp);
      const long double dz = p / dp;
      z -= dz;
      if (fabsl(dz) < 1e-19L) break;
    }
    legendre(n, z, p, dp);
    x[i] = 0.5L * (z + 1);
    w[i] = 1.0L / ((1 - z * z) * dp * dp);  // (2 / ((1 - z^2) P'^2)) / 2
  }
  return x;
}
std::vector<long double> gauss_lobatto_points(int n)  // n points on [0, 1]: the support points of FE_Q(n - 1)
{
  std::vector<long double> x(n);
  x[0] = 0;
  x[n - 1] = 1;
  const int m = n - 1;  // interior points: roots of P'_m
  const long double pi = acosl(-1.0L);
  for (int i = 1; i < n - 1; ++i) {
    long double z = -cosl(pi * i / m);
    for (int it = 0; it < 200; ++it) {
      // Newton on f = P'_m: f' from the Legendre ODE (1 - z^2) P'' = 2 z P' - m (m + 1) P
      long double p, dp;
      legendre(m, z, p, dp);
      const long double ddp = (2 * z * dp - m * (m + 1) * p) / (1 - z * z);
      const long double dz = dp / ddp;
      z -= dz;
      if (fabsl(dz) < 1e-19L) break;
    }
    x[i] = 0.5L * (z + 1);
  }
  return x;
}
long double lagrange(const std::vector<long double> &nodes, int j, long double x, bool derivative)
{
  const int n = (int)nodes.size();
  if (!derivative) {
    long double v = 1;
    for (int m = 0; m < n; ++m)
      if (m != j) v *= (x - nodes[m]) / (nodes[j] - nodes[m]);
    return v;
  }
  long double s = 0;
  for (int i = 0; i < n; ++i) {
    if (i == j) continue;
    long double v = 1 / (nodes[j] - nodes[i]);
    for (int m = 0; m < n; ++m)
      if (m != j && m != i) v *= (x - nodes[m]) / (nodes[j] - nodes[m]);
    s += v;
  }
  return s;
}
GenTables make_tables(int K)
{
  GenTables T{};
  const int n1 = K + 1;
  std::vector<long double> w;
  const std::vector<long double> xq = gauss_points(n1, w), xu = gauss_lobatto_points(n1), xp = K >= 2 ? gauss_lobatto_points(K) : std::vector<long double>{0.5L};
  for (int q = 0; q < n1; ++q) {
    T.W[q] = (double)w[q];
    for (int i = 0; i < n1; ++i) {
      T.N[q * kGenMaxN1 + i] = (double)lagrange(xu, i, xq[q], false);
      T.Dq[q * kGenMaxN1 + i] = (double)lagrange(xq, i, xq[q], true);
    }
    for (int i = 0; i < K; ++i) T.NP[q * kGenMaxN1 + i] = (double)lagrange(xp, i, xq[q], false);
  }
  return T;
}

template <int DIM, int K>
int launch_dk(const NSArgs &a, int kind, int with_p, int mode, int geom, cudaStream_t s)
{
  static const GenTables T = make_tables(K);
  constexpr int n1 = K + 1, NQ = DIM == 2 ? n1 * n1 : n1 * n1 * n1, CPB = GenCellsPerBlock<NQ>::value;
  if (!a.n_cells) return 0;
  const dim3 threads(NQ, CPB);
  const unsigned blocks = (a.n_cells + CPB - 1) / CPB;
  switch (kind) {
    case GEN_VMULT: ns_generic_kernel<DIM, K, GEN_VMULT><<<blocks, threads, 0, s>>>(a, T, with_p, mode, geom); break;
    case GEN_RESIDUAL: ns_generic_kernel<DIM, K, GEN_RESIDUAL><<<blocks, threads, 0, s>>>(a, T, with_p, mode, geom); break;
    case GEN_NODAL_MASS: ns_generic_kernel<DIM, K, GEN_NODAL_MASS><<<blocks, threads, 0, s>>>(a, T, with_p, mode, geom); break;
    case GEN_DIVERGENCE: ns_generic_kernel<DIM, K, GEN_DIVERGENCE><<<blocks, threads, 0, s>>>(a, T, 0, mode, geom); break;
    default: ns_generic_kernel<DIM, K, GEN_PRESSURE_MASS><<<blocks, threads, 0, s>>>(a, T, with_p, mode, geom); break;
  }
  AB_COUNT_LAUNCH();
  AB_CUDA(cudaGetLastError());
  return 0;
}
template <int DIM, int K>
int launch_scalar_dk(const NSArgs &a, int kind, int geom, cudaStream_t s)
{
  static const GenTables T = make_tables(K);
  constexpr int n1 = K + 1, NQ = DIM == 2 ? n1 * n1 : n1 * n1 * n1;
  const int threads = (NQ + 31) / 32 * 32;
  if (!a.n_cells) return 0;
  if (kind == SC_MASS)
    gen_scalar_kernel<DIM, K, SC_MASS><<<a.n_cells, threads, 0, s>>>(a, T, geom);
  else
    gen_scalar_kernel<DIM, K, SC_POISSON><<<a.n_cells, threads, 0, s>>>(a, T, geom);
  AB_COUNT_LAUNCH();
  AB_CUDA(cudaGetLastError());
  return 0;
}
}  // namespace

// scalar operators on the pressure space of degree degree_p with QGauss(degree_p + 1): kind 0 mass, 1 Laplace; coefficient a.scalar_c0
int launch_ns_generic_scalar(const NSArgs &a, int dim, int degree_p, int kind, int geom, cudaStream_t s)
{
  if (geom == ADAFLO_B200_GEOM_GENERAL && dim == 3 && degree_p == 2)
    return fail(ADAFLO_B200_ERR_UNSUPPORTED, "generic pressure operators: per-point geometry of QGauss(3) in 3-D is stored for the tuned kernels");
#define AB_G(D, K) \
  if (dim == D && degree_p == K) return launch_scalar_dk<D, K>(a, kind, geom, s);
  AB_G(2, 1) AB_G(2, 2) AB_G(2, 3) AB_G(2, 4) AB_G(3, 1) AB_G(3, 2) AB_G(3, 3) AB_G(3, 4)
#undef AB_G
  return fail(ADAFLO_B200_ERR_UNSUPPORTED, "generic pressure operators: pressure degree 1..4 in 2 or 3 dimensions");
}

// kind: 0 vmult (with_p = 0: velocity block), 1 residual, 2 integrals of the pressure shape functions into dst_p, 4 divergence
int launch_ns_generic(const NSArgs &a, int dim, int degree_u, int kind, int with_p, int mode, int geom, cudaStream_t s)
{
  if (geom == ADAFLO_B200_GEOM_GENERAL && dim == 3 && degree_u == 2) geom = 3;  // QGauss(3) in 3-D: decade-blocked arrays (api.cu mf_create)
#define AB_G(D, K) \
  if (dim == D && degree_u == K) return launch_dk<D, K>(a, kind, with_p, mode, geom, s);
  if (kind == GEN_NODAL_MASS) {
    AB_G(2, 1) AB_G(3, 1)
  }
  AB_G(2, 2) AB_G(2, 3) AB_G(2, 4) AB_G(2, 5) AB_G(3, 2) AB_G(3, 3) AB_G(3, 4) AB_G(3, 5)
#undef AB_G
  return fail(ADAFLO_B200_ERR_UNSUPPORTED, "ns operator: velocity degree 2..5 with pressure degree - 1 in 2 or 3 dimensions (navier_stokes_matrix.cc:64-82)");
}

}  // namespace ab200
