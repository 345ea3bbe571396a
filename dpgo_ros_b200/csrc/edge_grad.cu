// Riemannian gradient of LARGE agents straight from the edge list (the HBM-bound regime, BASELINE
// config 5: 12 500 poses and ~130 000 edges per agent).
//
//   egrad = X Q + G,  rgrad = Proj_X(egrad),  f = 1/2 <X Q, X> + <G, X>          (QuadraticProblem::f / EucGrad)
//
// matrix-free per edge e = (i -> j, R, t, kappa, tau, w), exactly the formulation of SURVEY 8(d):
//   E = X_j - X_i T,   T = [R t; 0 1],   Om = w diag(kappa, kappa, kappa, tau)
//   egrad_j += E Om,   egrad_i -= (E Om) T^T
// so the graph is read as ONE 128-byte record per edge  [R(9) | t(3) | w kappa | w tau | pad(2)]  -- the
// 128 B / edge of the roofline accounting -- instead of the two 4x4 blocks per edge (+ diagonal) of the block-CSR /
// ELL copy the small-agent phases use (2x the bytes).  No atomics, bitwise reproducible: the accumulation is BY POSE
// over a per-pose incidence list (edge id + which end this pose is); an edge's record is read once from HBM by
// whichever of its two poses comes first and hits L2 for the other (90 % of the loop closures of config 5 join poses
// less than 2000 apart).
//
// One WARP per pose: the four 8-lane groups take every fourth incident edge (lane a of a group holds row a of the
// r x 4 pose blocks, as everywhere in this library), issue the loads of up to kBatch edges before the first multiply
// -- incidence items (edge id + the pose at the other end), then the records and those poses together: two dependent
// trips per batch -- and are summed in a fixed order at the end.  The kernel is launched on its own (not inside the persistent kernel, whose one CTA of
// 8 warps per SM cannot keep enough bytes in flight): 16 warps per SM, each with 20 edges in flight.
#include "kernels.h"

namespace dpgo {

void count_launch();

namespace {

constexpr int kBatch = 5;   // edges per 8-lane group in flight (128 registers: two CTAs of 8 warps per SM)

__device__ __forceinline__ double gsum8e(double v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  return v;
}

// a = x T for one row x of an r x 4 pose block: [x_Y R | x_Y t + x_p]
__device__ __forceinline__ void row_times_T(const double (&x)[4], const double *rec, double (&a)[4]) {
#pragma unroll
  for (int c = 0; c < 3; ++c) a[c] = x[0] * rec[c * 3] + x[1] * rec[c * 3 + 1] + x[2] * rec[c * 3 + 2];
  a[3] = x[0] * rec[9] + x[1] * rec[10] + x[2] * rec[11] + x[3];
}
// acc += s * (z Om) T^T for one row z: [(z Om)_Y R^T + (z Om)_p t^T | (z Om)_p]
__device__ __forceinline__ void add_row_times_OmTt(const double (&z)[4], const double *rec, double s, double (&acc)[4]) {
  const double k = rec[12], tau = rec[13];
  const double z0 = z[0] * k, z1 = z[1] * k, z2 = z[2] * k, z3 = z[3] * tau;
#pragma unroll
  for (int c = 0; c < 3; ++c) acc[c] += s * (z0 * rec[c] + z1 * rec[3 + c] + z2 * rec[6 + c] + z3 * rec[9 + c]);
  acc[3] += s * z3;
}

}  // namespace

// pack the 128-byte records from the measurement arrays (after every weight change)
__global__ void k_pack_edge_records(MeasDev M, double *rec) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= M.count) return;
  double *o = rec + (size_t)e * 16;
  const double w = M.skip[e] ? 0.0 : M.w[e];
#pragma unroll
  for (int q = 0; q < 9; ++q) o[q] = M.R[(size_t)e * 9 + q];
#pragma unroll
  for (int q = 0; q < 3; ++q) o[9 + q] = M.t[(size_t)e * 3 + q];
  o[12] = w * M.kappa[e];
  o[13] = w * M.tau[e];
  o[14] = o[15] = 0.0;
}

template <int R>
__global__ void __launch_bounds__(256, 2) k_edge_grad(const __grid_constant__ EdgeGradArgs a) {
  __shared__ __align__(16) double stage_all[8][4][kBatch][16];   // warp, group, batch slot, record
  __shared__ double sm_part[8][2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 3, row = lane & 7;
  const int r = R;
  const bool act_row = row < r;
  if (a.tmarks && threadIdx.x == 0) atomicMin(a.tmarks, globaltimer_ns());
  double pf = 0, pg2 = 0;
  for (int j = blockIdx.x * 8 + warp; j < a.n; j += gridDim.x * 8) {
    const size_t off = (size_t)j * 4 * r;
    double x[4];
    ld4(a.Xin + off, r, row, act_row, x);
    double accq[4] = {0, 0, 0, 0}, accg[4] = {0, 0, 0, 0};
    const int lo = a.inc_ptr[j], hi = a.inc_ptr[j + 1];
    double(*stage)[16] = stage_all[warp][g];
    for (int base = lo; base < hi; base += 4 * kBatch) {
      // trip 1: this group's incidence items -- (edge * 2 + "this pose is the destination", pose at the other end)
      int2 it2[kBatch];
      int item[kBatch];
#pragma unroll
      for (int b = 0; b < kBatch; ++b) {
        const int k = base + 4 * b + g;
        it2[b] = k < hi ? a.inc_item[k] : make_int2(-1, 0);
        item[b] = it2[b].x;
      }
      // trip 2: the records (8 lanes x 16 B = one 128-byte line each) AND the pose at the other end
      double2 rq[kBatch];
      double xo[kBatch][4];
      bool remote[kBatch];
#pragma unroll
      for (int b = 0; b < kBatch; ++b) {
        rq[b] = item[b] >= 0 ? reinterpret_cast<const double2 *>(a.rec + (size_t)(item[b] >> 1) * 16)[row]
                             : make_double2(0.0, 0.0);
        const int other = it2[b].y;   // remote end: -(inbox slot + 1)
        remote[b] = other < 0;
        const double *src = remote[b] ? a.inbox + (size_t)(-other - 1) * 4 * r : a.Xin + (size_t)other * 4 * r;
        ld4(src, r, row, item[b] >= 0 && act_row && !(remote[b] && !a.build_g), xo[b]);
      }
#pragma unroll
      for (int b = 0; b < kBatch; ++b) reinterpret_cast<double2 *>(stage[b])[row] = rq[b];
      __syncwarp();
#pragma unroll
      for (int b = 0; b < kBatch; ++b) {
        if (item[b] < 0) continue;
        const double *rec = stage[b];
        const double k = rec[12], tau = rec[13];
        if (item[b] & 1) {
          // this pose is the destination j:  egrad_j += (X_j - X_i T) Om
          double at[4];
          row_times_T(xo[b], rec, at);
          accq[0] += x[0] * k; accq[1] += x[1] * k; accq[2] += x[2] * k; accq[3] += x[3] * tau;
          if (!remote[b]) {
            accq[0] -= at[0] * k; accq[1] -= at[1] * k; accq[2] -= at[2] * k; accq[3] -= at[3] * tau;
          } else if (a.build_g) {
            accg[0] -= at[0] * k; accg[1] -= at[1] * k; accg[2] -= at[2] * k; accg[3] -= at[3] * tau;
          }
        } else {
          // this pose is the source i:  egrad_i -= ((X_j - X_i T) Om) T^T
          double at[4];
          row_times_T(x, rec, at);
          add_row_times_OmTt(at, rec, 1.0, accq);
          if (!remote[b])
            add_row_times_OmTt(xo[b], rec, -1.0, accq);
          else if (a.build_g)
            add_row_times_OmTt(xo[b], rec, -1.0, accg);
        }
      }
      __syncwarp();
    }
    // the four groups' partial sums, in a fixed order: (g0 + g1) + (g2 + g3)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      accq[c] += __shfl_xor_sync(0xffffffffu, accq[c], 8);
      accq[c] += __shfl_xor_sync(0xffffffffu, accq[c], 16);
      accg[c] += __shfl_xor_sync(0xffffffffu, accg[c], 8);
      accg[c] += __shfl_xor_sync(0xffffffffu, accg[c], 16);
    }
    if (!a.build_g) ld4(a.G + off, r, row, act_row, accg);
    double eg[4];
    double f_part = 0;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      eg[c] = accq[c] + accg[c];
      f_part += (0.5 * accq[c] + accg[c]) * x[c];
    }
    // tangent projection of the rotation block: eg_Y -= Y sym(Y^T eg_Y)   (every group holds the same data)
    const double s00 = gsum8e(x[0] * eg[0]), s11 = gsum8e(x[1] * eg[1]), s22 = gsum8e(x[2] * eg[2]);
    const double s01 = 0.5 * gsum8e(x[0] * eg[1] + x[1] * eg[0]);
    const double s02 = 0.5 * gsum8e(x[0] * eg[2] + x[2] * eg[0]);
    const double s12 = 0.5 * gsum8e(x[1] * eg[2] + x[2] * eg[1]);
    double rg[4];
    rg[0] = eg[0] - (x[0] * s00 + x[1] * s01 + x[2] * s02);
    rg[1] = eg[1] - (x[0] * s01 + x[1] * s11 + x[2] * s12);
    rg[2] = eg[2] - (x[0] * s02 + x[1] * s12 + x[2] * s22);
    rg[3] = eg[3];
    if (g == 0 && act_row) {
      pf += f_part;
#pragma unroll
      for (int c = 0; c < 4; ++c) pg2 += rg[c] * rg[c];
      if (a.build_g) st4(a.G + off, r, row, true, accg);
      if (a.Rg) st4(a.Rg + off, r, row, true, rg);
      if (a.RgT) {
        const size_t n4 = (size_t)4 * a.n;
#pragma unroll
        for (int c = 0; c < 4; ++c) a.RgT[(size_t)row * n4 + 4 * j + c] = rg[c];
      }
    }
  }
  // per-CTA partials of f and |rgrad|^2 (summed over the CTAs in index order by whoever reads them)
  pf = wsum32(pf);
  pg2 = wsum32(pg2);
  if (lane == 0) {
    sm_part[warp][0] = pf;
    sm_part[warp][1] = pg2;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s0 = 0, s1 = 0;
    for (int w = 0; w < 8; ++w) {
      s0 += sm_part[w][0];
      s1 += sm_part[w][1];
    }
    a.partials[blockIdx.x * 2] = s0;
    a.partials[blockIdx.x * 2 + 1] = s1;
    if (a.tmarks) atomicMax(a.tmarks + 1, globaltimer_ns());
  }
}

cudaError_t launch_pack_edge_records(const MeasDev &M, double *rec, cudaStream_t s) {
  if (M.count == 0) return cudaSuccess;
  count_launch();
  k_pack_edge_records<<<(M.count + 127) / 128, 128, 0, s>>>(M, rec);
  return cudaGetLastError();
}

int edge_grad_grid(int n) { return std::max(1, (n + 7) / 8); }

cudaError_t launch_edge_grad(const EdgeGradArgs &a, int r, cudaStream_t s) {
  count_launch();
  const int grid = edge_grad_grid(a.n);
  switch (r) {
    case 3: k_edge_grad<3><<<grid, 256, 0, s>>>(a); break;
    case 4: k_edge_grad<4><<<grid, 256, 0, s>>>(a); break;
    case 5: k_edge_grad<5><<<grid, 256, 0, s>>>(a); break;
    case 6: k_edge_grad<6><<<grid, 256, 0, s>>>(a); break;
    case 7: k_edge_grad<7><<<grid, 256, 0, s>>>(a); break;
    case 8: k_edge_grad<8><<<grid, 256, 0, s>>>(a); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

}  // namespace dpgo
