// vector_ops.cu -- bandwidth-bound vector kernels of the Krylov step
// (SURVEY 8(a) rows a9, a17, a19): axpy / sadd / scale, fused multi-dot,
// add_and_dot, Jacobi apply (source/diagonal_preconditioner.cc:38-124),
// constrained-row copy and pressure mean projection
// (source/navier_stokes_matrix.cc:191-205,247-256), ghost pack/unpack.
//
// All kernels are grid-stride with 16-byte vector accesses where alignment
// allows; reductions are two-stage and deterministic (fixed block order).
#include <cstdlib>
#include "common.cuh"

namespace ab200 {

constexpr int kVecThreads = 256;

static inline unsigned grid_for(uint64_t n, int sm_count, int per_thread = 4)
{
  uint64_t blocks = (n + (uint64_t)kVecThreads * per_thread - 1) / ((uint64_t)kVecThreads * per_thread);
  const uint64_t cap = (uint64_t)sm_count * 8;
  if (blocks > cap) blocks = cap;
  if (blocks == 0) blocks = 1;
  return (unsigned)blocks;
}

static int sm_count_cached()  // per device: a process may drive several GPUs
{
  static int sm[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return 148;
  if (!sm[dev]) {
    cudaDeviceGetAttribute(&sm[dev], cudaDevAttrMultiProcessorCount, dev);
    if (sm[dev] <= 0) sm[dev] = 148;
  }
  return sm[dev];
}

__global__ void k_set(double *x, double v, uint64_t n)
{
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) x[i] = v;
}
__global__ void k_axpy(double *__restrict__ y, double a, const double *__restrict__ x, uint64_t n)
{
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    y[i] = fma(a, x[i], y[i]);
}
__global__ void k_sadd(double *__restrict__ y, double sc, double a, const double *__restrict__ x, uint64_t n)
{
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    y[i] = sc * y[i] + a * x[i];
}
__global__ void k_scale(double *__restrict__ y, const double *__restrict__ sc, uint64_t n)
{
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) y[i] *= sc[i];
}
__global__ void k_copy_indexed(double *__restrict__ dst, const double *__restrict__ src, const uint32_t *__restrict__ idx,
                               uint32_t n, double sign)
{
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t j = idx[i];
    dst[j] = sign * src[j];
  }
}
__global__ void k_pack(double *__restrict__ buf, const double *__restrict__ vec, const uint32_t *__restrict__ idx, uint32_t n_nodes,
                       int n_comp)
{
  const uint64_t n = (uint64_t)n_nodes * n_comp;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t node = idx[i / n_comp];
    buf[i] = vec[(uint64_t)node * n_comp + (i % n_comp)];
  }
}
__global__ void k_unpack_add(double *__restrict__ vec, const double *__restrict__ buf, const uint32_t *__restrict__ idx,
                             uint32_t n_nodes, int n_comp)
{
  const uint64_t n = (uint64_t)n_nodes * n_comp;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t node = idx[i / n_comp];
    vec[(uint64_t)node * n_comp + (i % n_comp)] += buf[i];  // send lists hold each node once
  }
}

int vec_set(cudaStream_t s, double *x, double v, uint64_t n)
{
  if (n == 0) return 0;
  if (v == 0.0) {
    AB_CUDA(cudaMemsetAsync(x, 0, n * sizeof(double), s));
    return 0;
  }
  k_set<<<grid_for(n, sm_count_cached()), kVecThreads, 0, s>>>(x, v, n);
  AB_COUNT_LAUNCH();
  AB_CUDA(cudaGetLastError());
  return 0;
}
int vec_axpy(cudaStream_t s, double *y, double a, const double *x, uint64_t n)
{
  if (n == 0) return 0;
  k_axpy<<<grid_for(n, sm_count_cached()), kVecThreads, 0, s>>>(y, a, x, n);
  AB_COUNT_LAUNCH();
  AB_CUDA(cudaGetLastError());
  return 0;
}
int vec_sadd(cudaStream_t s, double *y, double sc, double a, const double *x, uint64_t n)
{
  if (n == 0) return 0;
  k_sadd<<<grid_for(n, sm_count_cached()), kVecThreads, 0, s>>>(y, sc, a, x, n);
  AB_COUNT_LAUNCH();
  AB_CUDA(cudaGetLastError());
  return 0;
}
int vec_scale(cudaStream_t s, double *y, const double *sc, uint64_t n)
{
  if (n == 0) return 0;
  k_scale<<<grid_for(n, sm_count_cached()), kVecThreads, 0, s>>>(y, sc, n);
  AB_COUNT_LAUNCH();
  AB_CUDA(cudaGetLastError());
  return 0;
}
int vec_copy_indexed(cudaStream_t s, double *dst, const double *src, const uint32_t *idx, uint32_t n, double sign)
{
  if (n == 0) return 0;
  k_copy_indexed<<<grid_for(n, sm_count_cached(), 1), kVecThreads, 0, s>>>(dst, src, idx, n, sign);
  AB_COUNT_LAUNCH();
  AB_CUDA(cudaGetLastError());
  return 0;
}
__global__ void k_set_indexed(double *__restrict__ dst, const uint32_t *__restrict__ idx, uint32_t n, double value)
{
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[idx[i]] = value;
}
int vec_set_indexed(cudaStream_t s, double *dst, const uint32_t *idx, uint32_t n, double value)
{
  if (n == 0) return 0;
  k_set_indexed<<<grid_for(n, sm_count_cached(), 1), kVecThreads, 0, s>>>(dst, idx, n, value);
  AB_COUNT_LAUNCH();
  AB_CUDA(cudaGetLastError());
  return 0;
}
int vec_pack(cudaStream_t s, double *buf, const double *vec, const uint32_t *idx, uint32_t n_nodes, int n_comp)
{
  if (n_nodes == 0) return 0;
  k_pack<<<grid_for((uint64_t)n_nodes * n_comp, sm_count_cached(), 1), kVecThreads, 0, s>>>(buf, vec, idx, n_nodes, n_comp);
  AB_COUNT_LAUNCH();
  AB_CUDA(cudaGetLastError());
  return 0;
}
int vec_unpack_add(cudaStream_t s, double *vec, const double *buf, const uint32_t *idx, uint32_t n_nodes, int n_comp)
{
  if (n_nodes == 0) return 0;
  k_unpack_add<<<grid_for((uint64_t)n_nodes * n_comp, sm_count_cached(), 1), kVecThreads, 0, s>>>(vec, buf, idx, n_nodes, n_comp);
  AB_COUNT_LAUNCH();
  AB_CUDA(cudaGetLastError());
  return 0;
}

// -------------------------------------------------------------- reductions
constexpr int kMaxDots = 8;  // dots fused per pass
struct DotPtrs {
  const double *y[kMaxDots];
};

__device__ __forceinline__ double block_sum(double v, double *sh)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double r = 0;
  if (warp == 0) {
    r = lane < (blockDim.x >> 5) ? sh[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  }
  return r;  // valid in warp 0
}

// partial[j * gridDim.x + block] = sum over the block's slice of x . y_j ; optional fused update y0 += a x before the dot with z
template <int K>
__global__ void __launch_bounds__(kVecThreads) k_multi_dot(const double *__restrict__ x, DotPtrs ys, uint64_t n, double *__restrict__ partial)
{
  __shared__ double sh[32];
  double acc[K];
#pragma unroll
  for (int j = 0; j < K; ++j) acc[j] = 0;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const double xv = x[i];
#pragma unroll
    for (int j = 0; j < K; ++j) acc[j] = fma(xv, ys.y[j][i], acc[j]);
  }
#pragma unroll
  for (int j = 0; j < K; ++j) {
    const double r = block_sum(acc[j], sh);
    if (threadIdx.x == 0) partial[(size_t)j * gridDim.x + blockIdx.x] = r;
  }
}

// same with 16-byte loads, two per stream in flight (4 entries per thread and trip); all pointers 16-byte aligned.
// Measured on B200 at 50.9 M entries, K = 8: the 8-byte version reaches 5.3 TB/s (profiles/r2_i_launches.csv: 0.69 ms).
template <int K>
__global__ void __launch_bounds__(kVecThreads) k_multi_dot_v2(const double *__restrict__ x, DotPtrs ys, uint64_t n, double *__restrict__ partial)
{
  __shared__ double sh[32];
  double acc[K];
#pragma unroll
  for (int j = 0; j < K; ++j) acc[j] = 0;
  const uint64_t n4 = n / 4, stride = (uint64_t)gridDim.x * blockDim.x;
  const double2 *x2 = reinterpret_cast<const double2 *>(x);
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
    const double2 xa = __ldg(x2 + 2 * i), xb = __ldg(x2 + 2 * i + 1);
    double2 ya[K], yb[K];
#pragma unroll
    for (int j = 0; j < K; ++j) {
      ya[j] = __ldg(reinterpret_cast<const double2 *>(ys.y[j]) + 2 * i);
      yb[j] = __ldg(reinterpret_cast<const double2 *>(ys.y[j]) + 2 * i + 1);
    }
#pragma unroll
    for (int j = 0; j < K; ++j) acc[j] = fma(xb.y, yb[j].y, fma(xb.x, yb[j].x, fma(xa.y, ya[j].y, fma(xa.x, ya[j].x, acc[j]))));
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {  // tail
    const uint64_t i = 4 * n4 + threadIdx.x;
#pragma unroll
    for (int j = 0; j < K; ++j) acc[j] = fma(x[i], ys.y[j][i], acc[j]);
  }
#pragma unroll
  for (int j = 0; j < K; ++j) {
    const double r = block_sum(acc[j], sh);
    if (threadIdx.x == 0) partial[(size_t)j * gridDim.x + blockIdx.x] = r;
  }
}

// y += sum_j a[j] x_j: one read-modify-write of y for up to kMaxDots vectors
struct AxpyArgs {
  const double *x[kMaxDots];
  double a[kMaxDots];
};
template <int K>
__global__ void __launch_bounds__(kVecThreads) k_multi_axpy(double *__restrict__ y, AxpyArgs p, uint64_t n)
{
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    double v = y[i];
#pragma unroll
    for (int j = 0; j < K; ++j) v = fma(p.a[j], p.x[j][i], v);
    y[i] = v;
  }
}

__global__ void __launch_bounds__(kVecThreads) k_add_and_dot(double *__restrict__ y, double a, const double *__restrict__ x,
                                                            const double *__restrict__ z, uint64_t n, double *__restrict__ partial)
{
  __shared__ double sh[32];
  double acc = 0;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const double v = fma(a, x[i], y[i]);
    y[i] = v;
    acc = fma(v, z[i], acc);
  }
  const double r = block_sum(acc, sh);
  if (threadIdx.x == 0) partial[blockIdx.x] = r;
}

__global__ void __launch_bounds__(kVecThreads) k_absmax(const double *__restrict__ x, uint64_t n, double *__restrict__ partial)
{
  __shared__ double sh[32];
  double m = 0;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) m = fmax(m, fabs(x[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) sh[warp] = m;
  __syncthreads();
  if (warp == 0) {
    m = lane < (blockDim.x >> 5) ? sh[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) partial[blockIdx.x] = m;
  }
}

// second stage: one block per result, fixed summation order
__global__ void __launch_bounds__(kVecThreads) k_finish(const double *__restrict__ partial, int n_partial, double *__restrict__ out, int op_max)
{
  __shared__ double sh[32];
  const double *p = partial + (size_t)blockIdx.x * n_partial;
  double v = 0;
  for (int i = threadIdx.x; i < n_partial; i += blockDim.x) v = op_max ? fmax(v, p[i]) : v + p[i];
  if (op_max) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    if (warp == 0) {
      v = lane < (blockDim.x >> 5) ? sh[lane] : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
      if (lane == 0) out[blockIdx.x] = v;
    }
  } else {
    const double r = block_sum(v, sh);
    if (threadIdx.x == 0) out[blockIdx.x] = r;
  }
}

// layout of ctx->red_dev: [0, 64) results, [64, ...) partials
constexpr int kResultSlots = 64;

static int finish_to_host(adaflo_b200_ctx_t ctx, int k, int n_partial, int op_max, double *results_host, bool to_host,
                          int result_offset = 0)
{
  k_finish<<<k, kVecThreads, 0, ctx->stream>>>(ctx->red_dev + kResultSlots, n_partial, ctx->red_dev + result_offset, op_max);
  AB_COUNT_LAUNCH();
  AB_CUDA(cudaGetLastError());
  if (ctx->nranks > 1) AB_CHECK(comm_allreduce(ctx, ctx->red_dev + result_offset, k, op_max));
  if (to_host) {
    AB_CUDA(cudaMemcpyAsync(ctx->red_host, ctx->red_dev + result_offset, sizeof(double) * k, cudaMemcpyDeviceToHost, ctx->stream));
    AB_CHECK(stream_wait(ctx, ctx->stream));
    for (int j = 0; j < k; ++j) results_host[j] = ctx->red_host[j];
  }
  return 0;
}

int vec_multi_dot(adaflo_b200_ctx_t ctx, const double *x, const double *const *ys, int k, uint64_t n, double *results_host, bool to_host)
{
  if (k <= 0) return 0;
  if (k > kResultSlots) return fail(ADAFLO_B200_ERR_ARG, "multi_dot: at most 64 dots per call");
  const unsigned grid = grid_for(n, ctx->sm_count, 8);
  if ((size_t)kResultSlots + (size_t)kMaxDots * grid > ctx->red_capacity) return fail(ADAFLO_B200_ERR_STATE, "reduction scratch too small");
  for (int j0 = 0; j0 < k; j0 += kMaxDots) {
    const int kk = (k - j0 < kMaxDots) ? k - j0 : kMaxDots;
    DotPtrs p;
    for (int j = 0; j < kMaxDots; ++j) p.y[j] = ys[j0 + (j < kk ? j : 0)];
    double *partial = ctx->red_dev + kResultSlots;
    bool aligned = (reinterpret_cast<uintptr_t>(x) & 15u) == 0 && n >= 4096;
    for (int j = 0; j < kk; ++j) aligned = aligned && (reinterpret_cast<uintptr_t>(p.y[j]) & 15u) == 0;
    static const bool v2_enabled = [] { const char *e = std::getenv("ADAFLO_B200_DOT_V2"); return !(e && e[0] == '0'); }();
    if (aligned && v2_enabled) {
      switch (kk) {
        case 1: k_multi_dot_v2<1><<<grid, kVecThreads, 0, ctx->stream>>>(x, p, n, partial); break;
        case 2: k_multi_dot_v2<2><<<grid, kVecThreads, 0, ctx->stream>>>(x, p, n, partial); break;
        case 3: k_multi_dot_v2<3><<<grid, kVecThreads, 0, ctx->stream>>>(x, p, n, partial); break;
        case 4: k_multi_dot_v2<4><<<grid, kVecThreads, 0, ctx->stream>>>(x, p, n, partial); break;
        case 5: k_multi_dot_v2<5><<<grid, kVecThreads, 0, ctx->stream>>>(x, p, n, partial); break;
        case 6: k_multi_dot_v2<6><<<grid, kVecThreads, 0, ctx->stream>>>(x, p, n, partial); break;
        case 7: k_multi_dot_v2<7><<<grid, kVecThreads, 0, ctx->stream>>>(x, p, n, partial); break;
        default: k_multi_dot_v2<8><<<grid, kVecThreads, 0, ctx->stream>>>(x, p, n, partial); break;
      }
    } else
    switch (kk) {
      case 1: k_multi_dot<1><<<grid, kVecThreads, 0, ctx->stream>>>(x, p, n, partial); break;
      case 2: k_multi_dot<2><<<grid, kVecThreads, 0, ctx->stream>>>(x, p, n, partial); break;
      case 3: k_multi_dot<3><<<grid, kVecThreads, 0, ctx->stream>>>(x, p, n, partial); break;
      case 4: k_multi_dot<4><<<grid, kVecThreads, 0, ctx->stream>>>(x, p, n, partial); break;
      case 5: k_multi_dot<5><<<grid, kVecThreads, 0, ctx->stream>>>(x, p, n, partial); break;
      case 6: k_multi_dot<6><<<grid, kVecThreads, 0, ctx->stream>>>(x, p, n, partial); break;
      case 7: k_multi_dot<7><<<grid, kVecThreads, 0, ctx->stream>>>(x, p, n, partial); break;
      default: k_multi_dot<8><<<grid, kVecThreads, 0, ctx->stream>>>(x, p, n, partial); break;
    }
    AB_COUNT_LAUNCH();
  AB_CUDA(cudaGetLastError());
    k_finish<<<kk, kVecThreads, 0, ctx->stream>>>(partial, (int)grid, ctx->red_dev + j0, 0);
    AB_COUNT_LAUNCH();
  AB_CUDA(cudaGetLastError());
  }
  if (ctx->nranks > 1) AB_CHECK(comm_allreduce(ctx, ctx->red_dev, k, 0));
  if (to_host) {
    AB_CUDA(cudaMemcpyAsync(ctx->red_host, ctx->red_dev, sizeof(double) * k, cudaMemcpyDeviceToHost, ctx->stream));
    AB_CHECK(stream_wait(ctx, ctx->stream));
    for (int j = 0; j < k; ++j) results_host[j] = ctx->red_host[j];
  }
  return 0;
}

int vec_multi_axpy(adaflo_b200_ctx_t ctx, double *y, const double *a_host, const double *const *xs, int k, uint64_t n)
{
  if (n == 0) return 0;
  const unsigned grid = grid_for(n, ctx->sm_count, 8);
  for (int j0 = 0; j0 < k; j0 += kMaxDots) {
    const int kk = k - j0 < kMaxDots ? k - j0 : kMaxDots;
    AxpyArgs p;
    for (int j = 0; j < kMaxDots; ++j) {
      p.x[j] = xs[j0 + (j < kk ? j : 0)];
      p.a[j] = j < kk ? a_host[j0 + j] : 0.0;
    }
    switch (kk) {
      case 1: k_multi_axpy<1><<<grid, kVecThreads, 0, ctx->stream>>>(y, p, n); break;
      case 2: k_multi_axpy<2><<<grid, kVecThreads, 0, ctx->stream>>>(y, p, n); break;
      case 3: k_multi_axpy<3><<<grid, kVecThreads, 0, ctx->stream>>>(y, p, n); break;
      case 4: k_multi_axpy<4><<<grid, kVecThreads, 0, ctx->stream>>>(y, p, n); break;
      case 5: k_multi_axpy<5><<<grid, kVecThreads, 0, ctx->stream>>>(y, p, n); break;
      case 6: k_multi_axpy<6><<<grid, kVecThreads, 0, ctx->stream>>>(y, p, n); break;
      case 7: k_multi_axpy<7><<<grid, kVecThreads, 0, ctx->stream>>>(y, p, n); break;
      default: k_multi_axpy<8><<<grid, kVecThreads, 0, ctx->stream>>>(y, p, n); break;
    }
    AB_COUNT_LAUNCH();
    AB_CUDA(cudaGetLastError());
  }
  return 0;
}

int vec_add_and_dot(adaflo_b200_ctx_t ctx, double *y, double a, const double *x, const double *z, uint64_t n, double *result_host)
{
  const unsigned grid = grid_for(n, ctx->sm_count, 8);
  k_add_and_dot<<<grid, kVecThreads, 0, ctx->stream>>>(y, a, x, z, n, ctx->red_dev + kResultSlots);
  AB_COUNT_LAUNCH();
  AB_CUDA(cudaGetLastError());
  return finish_to_host(ctx, 1, (int)grid, 0, result_host, true);
}

// p -= (w . p) * inv_weight * mode, the scalar never leaves the device
__global__ void k_project(double *__restrict__ p, const double *__restrict__ mode, const double *__restrict__ product, double inv_weight, uint64_t n)
{
  const double f = -(*product) * inv_weight;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) p[i] = fma(f, mode[i], p[i]);
}

int vec_project(adaflo_b200_ctx_t ctx, double *p, const double *w, const double *mode, double inv_weight, uint64_t n_owned, uint64_t n_local)
{
  const double *ys[1] = {p};
  AB_CHECK(vec_multi_dot(ctx, w, ys, 1, n_owned, nullptr, false));
  k_project<<<grid_for(n_local, ctx->sm_count), kVecThreads, 0, ctx->stream>>>(p, mode, ctx->red_dev, inv_weight, n_local);
  AB_COUNT_LAUNCH();
  AB_CUDA(cudaGetLastError());
  return 0;
}

// DiagonalPreconditioner::reinit, diagonal_preconditioner.cc:38-60
__global__ void k_jacobi_reinit(double *__restrict__ inv, const double *__restrict__ diag, const double *__restrict__ linfty, uint64_t n)
{
  const double thr = 1e-10 * (*linfty);
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const double d = diag[i];
    inv[i] = fabs(d) > thr ? 1.0 / d : 1.0;
  }
}

int jacobi_reinit(adaflo_b200_ctx_t ctx, double *inv, const double *diag, uint64_t n)
{
  const unsigned grid = grid_for(n, ctx->sm_count, 8);
  k_absmax<<<grid, kVecThreads, 0, ctx->stream>>>(diag, n, ctx->red_dev + kResultSlots);
  AB_COUNT_LAUNCH();
  AB_CUDA(cudaGetLastError());
  AB_CHECK(finish_to_host(ctx, 1, (int)grid, 1, nullptr, false));
  k_jacobi_reinit<<<grid_for(n, ctx->sm_count), kVecThreads, 0, ctx->stream>>>(inv, diag, ctx->red_dev, n);
  AB_COUNT_LAUNCH();
  AB_CUDA(cudaGetLastError());
  return 0;
}

// DiagonalPreconditioner::vmult, diagonal_preconditioner.cc:99-124: one scalar-space
// inverse applied to every block (n_blocks > 1) or an equally shaped inverse
__global__ void k_jacobi(double *__restrict__ dst, const double *__restrict__ src, const double *__restrict__ inv, uint64_t n, int n_blocks)
{
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const double d = inv[i];
    for (int b = 0; b < n_blocks; ++b) dst[(uint64_t)b * n + i] = src[(uint64_t)b * n + i] * d;
  }
}

int jacobi_vmult(cudaStream_t s, double *dst, const double *src, const double *inv, uint64_t n, int n_blocks)
{
  if (n == 0) return 0;
  k_jacobi<<<grid_for(n, sm_count_cached()), kVecThreads, 0, s>>>(dst, src, inv, n, n_blocks);
  AB_COUNT_LAUNCH();
  AB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace ab200
