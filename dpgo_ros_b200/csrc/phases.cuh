// Phases of one RBCD iteration as grid-cooperative device functions.  Each
// phase is called by every CTA of the grid; phases are separated by
// grid_barrier / grid_reduce in the persistent kernel (kernels.cu) or by kernel
// boundaries in the single-op kernels that back the parity hooks.
//
// Reference call sites (relative to the reference repo): the arithmetic lives
// in the un-vendored mit-acl/dpgo; what is cited is the wrapper line that
// triggers it.  iterate(): src/PGOAgentROS.cpp:160,1185.
#pragma once
#include "device.cuh"

namespace dpgo {

// ---- pose -> 8-lane-group mapping --------------------------------------------
// item (pose) = blockIdx.x + gridDim.x * (local_group + 32 k): consecutive poses
// land on different SMs so a 300-pose agent is spread over the whole chip.
struct PoseIter {
  int a;       // row handled by this lane
  int lg;      // local group id 0..31
  int k;       // loop counter
  __device__ __forceinline__ PoseIter() : a(threadIdx.x & 7), lg(threadIdx.x >> 3), k(0) {}
  // warp-uniform "any lane of my warp still has work" + my item
  __device__ __forceinline__ bool next(int total, int &item) {
    const int g0 = (lg & ~3) + 32 * k;  // first group of my warp at this step
    if ((int)blockIdx.x + (int)gridDim.x * g0 >= total) return false;
    item = (int)blockIdx.x + (int)gridDim.x * (lg + 32 * k);
    ++k;
    return true;
  }
};

__device__ __forceinline__ void publish(const int *rowptr, double *const *dst, int j, int r, int a, bool act,
                                        const double (&x)[4]) {
  const int e0 = rowptr[j], e1 = rowptr[j + 1];
  for (int e = e0; e < e1; ++e) st4(dst[e], r, a, act, x);
}

// out_row(1x4) += x_row(1x4) * B(4x4 col-major)
__device__ __forceinline__ void row_times_block(const double (&x)[4], const double *__restrict__ B, double (&acc)[4]) {
#pragma unroll
  for (int cp = 0; cp < 4; ++cp) {
    const double b0 = B[cp * 4 + 0], b1 = B[cp * 4 + 1], b2 = B[cp * 4 + 2], b3 = B[cp * 4 + 3];
    acc[cp] = fma(x[0], b0, fma(x[1], b1, fma(x[2], b2, fma(x[3], b3, acc[cp]))));
  }
}

// ---------------------------------------------------------------------------
// Phase A -- Nesterov bookkeeping of iterate() for every local agent (a7):
//   Y = proj((1-alpha) X + alpha V); non-selected agents: X = Y (V = proj(V) = V);
//   restart iterations: non-selected agents V = Y = X.
// Publishes Y (aux) and, for non-selected agents, X (reg) into the neighbours'
// inboxes (a9: getAuxSharedPoseDictWithNeighbor :666 / updateAuxNeighborPoses :1278).
// ---------------------------------------------------------------------------
__device__ __forceinline__ void phase_nesterov(const TeamDev &T, int sel_local, bool restart, double alpha) {
  PoseIter it;
  const int total = T.pose_prefix[T.num_local];
  int item;
  while (it.next(total, item)) {
    const bool valid = item < total;
    int ai = 0;
    if (valid) {
      while (item >= T.pose_prefix[ai + 1]) ++ai;
    }
    const AgentDev &A = T.ag[ai];
    const int j = valid ? item - T.pose_prefix[ai] : 0;
    const int r = A.r;
    const bool act = valid && it.a < r;
    const size_t off = (size_t)j * 4 * r;
    double x[4], v[4], m[4];
    ld4(A.X + off, r, it.a, act, x);
    if (restart) {
      if (ai != sel_local && valid) {
        st4(A.V + off, r, it.a, act, x);
        st4(A.Y + off, r, it.a, act, x);
        publish(A.pub_rowptr, A.pub_dst_aux, j, r, it.a, act, x);
        publish(A.pub_rowptr, A.pub_dst_reg, j, r, it.a, act, x);
      }
      continue;
    }
    ld4(A.V + off, r, it.a, act, v);
#pragma unroll
    for (int c = 0; c < 4; ++c) m[c] = (1.0 - alpha) * x[c] + alpha * v[c];
    if (!valid) {  // keep idle groups on the fast path of sym3_invsqrt
      m[0] = (it.a == 0);
      m[1] = (it.a == 1);
      m[2] = (it.a == 2);
    }
    stiefel_project_row(m);
    if (valid) {
      st4(A.Y + off, r, it.a, act, m);
      publish(A.pub_rowptr, A.pub_dst_aux, j, r, it.a, act, m);
      if (ai != sel_local) {
        st4(A.X + off, r, it.a, act, m);
        publish(A.pub_rowptr, A.pub_dst_reg, j, r, it.a, act, m);
      }
    }
  }
}

// ---------------------------------------------------------------------------
// Cost / gradient (a3, a4): egrad = Xin Q + G, rgrad = Proj_Xin(egrad),
// f = 0.5 <Xin Q, Xin> + <G, Xin>.  G is (re)assembled from the inbox when
// build_g (a4: "G rebuilt every iteration", updateNeighborPoses :1276).
// Writes G (if build_g), S = sym(Y^T egrad_Y) per pose, Rg and its row-major
// copy RgT.  Accumulates partial f and |rgrad|^2 into part[0], part[1].
// ---------------------------------------------------------------------------
__device__ __forceinline__ void phase_grad(const AgentDev &A, const double *Xin,
                                           const double *inbox, bool build_g, double *Sout,
                                           double *Rgout, double *RgTout, double *egrad_out, double &pf, double &pg2) {
  PoseIter it;
  const int n = A.n, r = A.r;
  const size_t n4 = (size_t)4 * n;
  int j;
  while (it.next(n, j)) {
    const bool valid = j < n;
    const bool act = valid && it.a < r;
    const size_t off = (size_t)(valid ? j : 0) * 4 * r;
    double x[4], accq[4] = {0, 0, 0, 0}, accg[4] = {0, 0, 0, 0};
    ld4(Xin + off, r, it.a, act, x);
    if (valid) {
      const int e0 = A.q_rowptr[j], e1 = A.q_rowptr[j + 1];
      for (int e = e0; e < e1; ++e) {
        double xi[4];
        ld4(Xin + (size_t)A.q_col[e] * 4 * r, r, it.a, act, xi);
        row_times_block(xi, A.q_val + (size_t)e * 16, accq);
      }
      if (build_g) {
        const int s0 = A.s_rowptr[j], s1 = A.s_rowptr[j + 1];
        for (int e = s0; e < s1; ++e) {
          double xi[4];
          ld4(inbox + (size_t)A.s_slot[e] * 4 * r, r, it.a, act, xi);
          row_times_block(xi, A.s_val + (size_t)e * 16, accg);
        }
        st4(A.G + off, r, it.a, act, accg);
      } else {
        ld4(A.G + off, r, it.a, act, accg);
      }
    }
    double eg[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      eg[c] = accq[c] + accg[c];
      pf += (0.5 * accq[c] + accg[c]) * x[c];
    }
    if (egrad_out) st4(egrad_out + off, r, it.a, act, eg);
    const Sym3 S = sym_ytz(x, eg);
    sub_y_sym(x, S, eg);
#pragma unroll
    for (int c = 0; c < 4; ++c) pg2 += eg[c] * eg[c];
    if (valid) {
      if (Sout && it.a == 0) {
        double *s = Sout + (size_t)j * 6;
        s[0] = S.a00; s[1] = S.a01; s[2] = S.a02; s[3] = S.a11; s[4] = S.a12; s[5] = S.a22;
      }
      if (Rgout) st4(Rgout + off, r, it.a, act, eg);
      if (RgTout && act) {
#pragma unroll
        for (int c = 0; c < 4; ++c) RgTout[(size_t)it.a * n4 + 4 * j + c] = eg[c];
      }
    }
  }
}

// ---------------------------------------------------------------------------
// Riemannian Hessian-vector product at Xbase (a3/a5):
//   H = Proj_Xbase( V Q - V_Y * S ),  S = sym(Y^T egrad_Y) cached per pose.
// pvh accumulates <V, H>.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void phase_hess(const AgentDev &A, const double *Xbase,
                                           const double *S, const double *Vin,
                                           double *Hout, double &pvh) {
  PoseIter it;
  const int n = A.n, r = A.r;
  int j;
  while (it.next(n, j)) {
    const bool valid = j < n;
    const bool act = valid && it.a < r;
    const size_t off = (size_t)(valid ? j : 0) * 4 * r;
    double x[4], v[4], h[4] = {0, 0, 0, 0};
    ld4(Xbase + off, r, it.a, act, x);
    ld4(Vin + off, r, it.a, act, v);
    Sym3 Sj = {0, 0, 0, 0, 0, 0};
    if (valid) {
      const int e0 = A.q_rowptr[j], e1 = A.q_rowptr[j + 1];
      for (int e = e0; e < e1; ++e) {
        double vi[4];
        ld4(Vin + (size_t)A.q_col[e] * 4 * r, r, it.a, act, vi);
        row_times_block(vi, A.q_val + (size_t)e * 16, h);
      }
      const double *s = S + (size_t)j * 6;
      Sj.a00 = s[0]; Sj.a01 = s[1]; Sj.a02 = s[2]; Sj.a11 = s[3]; Sj.a12 = s[4]; Sj.a22 = s[5];
    }
    sub_y_sym(v, Sj, h);
    tangent_project_row(x, h);
#pragma unroll
    for (int c = 0; c < 4; ++c) pvh += v[c] * h[c];
    if (valid) st4(Hout + off, r, it.a, act, h);
  }
}

// ---------------------------------------------------------------------------
// Dense preconditioner slab (a6):  Z[:, cols] = V * Pinv[:, cols] for the
// columns of poses [p0, p0+np) -- this CTA's share.  V is read through its
// row-major copy VT ([R][n4]) so lanes read consecutive q; Pinv is column-major
// with leading dimension ldp, so the same holds for it.  Result goes to shared
// memory zs[pose_local][c][8].  Thread q-strided accumulation, then a
// reduce-scatter over the warp and a cross-warp sum.
// ---------------------------------------------------------------------------
template <int R>
__device__ __forceinline__ void dense_slab(const double *__restrict__ Pinv, size_t ldp,
                                           const double *VT, int r, int n4, int p0, int np, double *zs,
                                           double *red /* smem [8 warps][8 cols][8] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int sub = 0; sub < np; sub += 2) {
    const int ncols = min(2, np - sub) * 4;
    const size_t col0 = (size_t)4 * (p0 + sub);
    double acc[R][8];
#pragma unroll
    for (int a = 0; a < R; ++a)
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[a][c] = 0.0;
    for (int q = threadIdx.x; q < n4; q += kThreads) {
      double vr[R];
#pragma unroll
      for (int a = 0; a < R; ++a) vr[a] = (a < r) ? VT[(size_t)a * n4 + q] : 0.0;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        if (c < ncols) {
          const double pv = Pinv[(col0 + c) * ldp + q];
#pragma unroll
          for (int a = 0; a < R; ++a) acc[a][c] = fma(vr[a], pv, acc[a][c]);
        }
      }
    }
    // reduce-scatter over the 8 columns (xor 1, 2, 4), then butterfly (8, 16)
    double h4[R][4], h2[R][2], h1[R];
    {
      const bool hi = lane & 1;
#pragma unroll
      for (int a = 0; a < R; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const double keep = hi ? acc[a][c + 4] : acc[a][c];
          const double send = hi ? acc[a][c] : acc[a][c + 4];
          h4[a][c] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
        }
    }
    {
      const bool hi = lane & 2;
#pragma unroll
      for (int a = 0; a < R; ++a)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const double keep = hi ? h4[a][c + 2] : h4[a][c];
          const double send = hi ? h4[a][c] : h4[a][c + 2];
          h2[a][c] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
        }
    }
    {
      const bool hi = lane & 4;
#pragma unroll
      for (int a = 0; a < R; ++a) {
        const double keep = hi ? h2[a][1] : h2[a][0];
        const double send = hi ? h2[a][0] : h2[a][1];
        double v = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        h1[a] = v;
      }
    }
    // lane l < 8 holds column ((l&1)<<2 | (l&2) | (l&4)>>2)
    __syncthreads();  // red reuse
    if (lane < 8) {
      const int col = ((lane & 1) << 2) | (lane & 2) | ((lane & 4) >> 2);
#pragma unroll
      for (int a = 0; a < R; ++a) red[(warp * 8 + col) * 8 + a] = h1[a];
    }
    __syncthreads();
    if (threadIdx.x < 64) {
      const int col = threadIdx.x >> 3, a = threadIdx.x & 7;
      if (a < r && col < ncols) {
        double s = 0;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) s += red[(w * 8 + col) * 8 + a];
        zs[((sub + (col >> 2)) * 4 + (col & 3)) * 8 + a] = s;
      }
    }
  }
  __syncthreads();
}

// this CTA's balanced share of an agent's poses for the dense phase
__device__ __forceinline__ void cta_pose_chunk(int n, int &p0, int &np) {
  const int base = n / (int)gridDim.x, rem = n % (int)gridDim.x;
  const int b = (int)blockIdx.x;
  np = base + (b < rem ? 1 : 0);
  p0 = b * base + min(b, rem);
}

// ---------------------------------------------------------------------------
// Commit a new iterate for pose j (group-collective): relative-change partial,
// X <- xnew, publish, and the Nesterov V update (a7):
//   V = proj(V + gamma (X+ - Y))    or, on restart iterations, V = Y = X+.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void finish_pose(const AgentDev &A, int j, bool valid, int a, const double (&xnew)[4],
                                            bool accel, bool restart, double gamma, double &prel) {
  const int r = A.r;
  const bool act = valid && a < r;
  const size_t off = (size_t)(valid ? j : 0) * 4 * r;
  double xold[4];
  ld4(A.X + off, r, a, act, xold);
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const double d = xnew[c] - xold[c];
    prel += act ? d * d : 0.0;
  }
  if (valid) {
    st4(A.X + off, r, a, act, xnew);
    publish(A.pub_rowptr, A.pub_dst_reg, j, r, a, act, xnew);
  }
  if (accel) {
    if (restart) {
      if (valid) {
        st4(A.V + off, r, a, act, xnew);
        st4(A.Y + off, r, a, act, xnew);
        publish(A.pub_rowptr, A.pub_dst_aux, j, r, a, act, xnew);
      }
    } else {
      double y[4], v[4], m[4];
      ld4(A.Y + off, r, a, act, y);
      ld4(A.V + off, r, a, act, v);
#pragma unroll
      for (int c = 0; c < 4; ++c) m[c] = v[c] + gamma * (xnew[c] - y[c]);
      if (!valid) {
        m[0] = (a == 0);
        m[1] = (a == 1);
        m[2] = (a == 2);
      }
      stiefel_project_row(m);
      if (valid) st4(A.V + off, r, a, act, m);
    }
  }
}

// ---------------------------------------------------------------------------
// RGD step (a2, src/PGOAgentROSNode.cpp:96-97):
//   X+ = Retr_Xs( -eta * Proj_Xs( P^-1 rgrad ) ), fused with finish_pose.
// With the preconditioner the CTA first computes its dense slab.
// ---------------------------------------------------------------------------
template <int R>
__device__ __forceinline__ void phase_rgd_step(const AgentDev &A, const SolverParams &P, const double *Xs,
                                               bool accel, bool restart, double gamma, double *zs, double *red,
                                               double &prel) {
  const int n = A.n, r = A.r;
  const int a = threadIdx.x & 7, lg = threadIdx.x >> 3;
  if (P.rgd_use_precond) {
    int p0, np;
    cta_pose_chunk(n, p0, np);
    const size_t ldp = ((size_t)4 * n + 31) / 32 * 32;
    dense_slab<R>(A.Pinv, ldp, A.RgT, r, 4 * n, p0, np, zs, red);
    for (int k0 = 0; k0 < np; k0 += kGroupsPerCta) {
      const int k = k0 + lg;
      const bool valid = k < np;
      const int j = p0 + (valid ? k : 0);
      const bool act = valid && a < r;
      double y[4], z[4];
      ld4(Xs + (size_t)j * 4 * r, r, a, act, y);
#pragma unroll
      for (int c = 0; c < 4; ++c) z[c] = act ? zs[((valid ? k : 0) * 4 + c) * 8 + a] : 0.0;
      tangent_project_row(y, z);
      double xn[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) xn[c] = y[c] - P.rgd_stepsize * z[c];
      if (!valid) {
        xn[0] = (a == 0); xn[1] = (a == 1); xn[2] = (a == 2);
      }
      qf_row(xn);
      finish_pose(A, j, valid, a, xn, accel, restart, gamma, prel);
    }
  } else {
    PoseIter it;
    int j;
    while (it.next(n, j)) {
      const bool valid = j < n;
      const bool act = valid && it.a < r;
      const size_t off = (size_t)(valid ? j : 0) * 4 * r;
      double y[4], z[4], xn[4];
      ld4(Xs + off, r, it.a, act, y);
      ld4(A.Rg + off, r, it.a, act, z);
#pragma unroll
      for (int c = 0; c < 4; ++c) xn[c] = y[c] - P.rgd_stepsize * z[c];
      if (!valid) {
        xn[0] = (it.a == 0); xn[1] = (it.a == 1); xn[2] = (it.a == 2);
      }
      qf_row(xn);
      finish_pose(A, j, valid, it.a, xn, accel, restart, gamma, prel);
    }
  }
}

// Z = Proj_Xbase( V Pinv ) for the whole agent (tCG preconditioner, a6).  Writes
// Z (+ optional row-major copy) and optionally dlt = -Z; pzr accumulates <Z, Rin>.
template <int R>
__device__ __forceinline__ void phase_precond(const AgentDev &A, const double *Xbase, const double *Rin,
                                              const double *RinT, double *Zout, double *neg_out, double *zs,
                                              double *red, double &pzr) {
  const int n = A.n, r = A.r;
  const int a = threadIdx.x & 7, lg = threadIdx.x >> 3;
  int p0, np;
  cta_pose_chunk(n, p0, np);
  const size_t ldp = ((size_t)4 * n + 31) / 32 * 32;
  dense_slab<R>(A.Pinv, ldp, RinT, r, 4 * n, p0, np, zs, red);
  for (int k0 = 0; k0 < np; k0 += kGroupsPerCta) {
    const int k = k0 + lg;
    const bool valid = k < np;
    const int j = p0 + (valid ? k : 0);
    const bool act = valid && a < r;
    const size_t off = (size_t)j * 4 * r;
    double y[4], z[4], rr[4];
    ld4(Xbase + off, r, a, act, y);
    ld4(Rin + off, r, a, act, rr);
#pragma unroll
    for (int c = 0; c < 4; ++c) z[c] = act ? zs[((valid ? k : 0) * 4 + c) * 8 + a] : 0.0;
    tangent_project_row(y, z);
#pragma unroll
    for (int c = 0; c < 4; ++c) pzr += z[c] * rr[c];
    if (valid) {
      st4(Zout + off, r, a, act, z);
      if (neg_out) {
        double nz[4] = {-z[0], -z[1], -z[2], -z[3]};
        st4(neg_out + off, r, a, act, nz);
      }
    }
  }
}

}  // namespace dpgo
