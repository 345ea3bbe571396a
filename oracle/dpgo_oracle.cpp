// CPU ORACLE -- test infrastructure only; see the header for scope and the
// "parity unpinned" statement.  Every function cites the reference call site
// (relative to /root/reference) whose behaviour it restates.
#include "dpgo_oracle.hpp"

#include <algorithm>
#include <array>
#include <atomic>
#include <cassert>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <stdexcept>
#include <thread>

namespace dpgo_oracle {

// ============================================================================
// dense helpers
// ============================================================================
double dot(const Mat &A, const Mat &B) {
  assert(A.size() == B.size());
  double s = 0;
  for (size_t i = 0; i < A.a.size(); ++i) s += A.a[i] * B.a[i];
  return s;
}
double squaredNorm(const Mat &A) { return dot(A, A); }
void axpy(double alpha, const Mat &X, Mat &Y) {
  assert(X.size() == Y.size());
  for (size_t i = 0; i < X.a.size(); ++i) Y.a[i] += alpha * X.a[i];
}

// 4x4 column-major block helpers: C (+)= A*B, etc.
static inline void mm44(const double *A, const double *B, double *C) {  // C = A B
  for (int j = 0; j < 4; ++j)
    for (int i = 0; i < 4; ++i) {
      double s = 0;
      for (int k = 0; k < 4; ++k) s += A[k * 4 + i] * B[j * 4 + k];
      C[j * 4 + i] = s;
    }
}
static inline void mm44_nt(const double *A, const double *B, double *C) {  // C = A B^T
  for (int j = 0; j < 4; ++j)
    for (int i = 0; i < 4; ++i) {
      double s = 0;
      for (int k = 0; k < 4; ++k) s += A[k * 4 + i] * B[k * 4 + j];
      C[j * 4 + i] = s;
    }
}
// out(r x 4) += alpha * Xi(r x 4) * B(4 x 4)
static inline void rmm(int r, double alpha, const double *Xi, const double *B, double *out) {
  for (int j = 0; j < 4; ++j)
    for (int k = 0; k < 4; ++k) {
      const double b = alpha * B[j * 4 + k];
      if (b == 0.0) continue;
      const double *x = Xi + (size_t)k * r;
      double *o = out + (size_t)j * r;
      for (int a = 0; a < r; ++a) o[a] += x[a] * b;
    }
}

// T~ = [R t; 0 1] and Omega = w diag(kappa,kappa,kappa,tau)  
static void edgeBlocks(const Measurement &m, double *T /*4x4*/, double *Om /*4 diag*/) {
  std::memset(T, 0, 16 * sizeof(double));
  for (int j = 0; j < 3; ++j)
    for (int i = 0; i < 3; ++i) T[j * 4 + i] = m.R[j * 3 + i];
  for (int i = 0; i < 3; ++i) T[3 * 4 + i] = m.t[i];
  T[15] = 1.0;
  Om[0] = Om[1] = Om[2] = m.weight * m.kappa;
  Om[3] = m.weight * m.tau;
}

// ============================================================================
// RobustCost (a8)  -- weight() is called by the wrapper at
// src/PGOAgentROS.cpp:1050; parameters from src/PGOAgentROSNode.cpp:174-221.
// ============================================================================
void RobustCost::configure(const Params &p) {
  type = p.costType;
  barcSq = p.GNCBarc * p.GNCBarc;
  muStep = p.GNCMuStep;
  initMu = p.GNCInitMu;
  mu = initMu;
}
void RobustCost::update() {
  if (type == CostType::GNC_TLS) mu = mu * muStep;
}
double RobustCost::weight(double r) const {
  switch (type) {
    case CostType::L2:
      return 1.0;
    case CostType::GNC_TLS: {
      // GNC-TLS weight, eq. (14) of Yang et al. 2020
      const double upper = (mu + 1.0) / mu * barcSq;
      const double lower = mu / (mu + 1.0) * barcSq;
      const double rSq = r * r;
      if (rSq >= upper) return 0.0;
      if (rSq <= lower) return 1.0;
      return std::sqrt(barcSq * mu * (mu + 1.0) / rSq) - mu;
    }
    default:
      throw std::runtime_error("RobustCost: cost type out of scope (SURVEY §2 #4: only L2 and GNC_TLS)");
  }
}

// ============================================================================
// BlockCholesky -- stands in for CHOLMOD on Q + lambda I (a6)
// ============================================================================
static void chol44(double *A) {  // in-place lower Cholesky of a 4x4 SPD block (column-major)
  for (int j = 0; j < 4; ++j) {
    double d = A[j * 4 + j];
    for (int k = 0; k < j; ++k) d -= A[k * 4 + j] * A[k * 4 + j];
    if (!(d > 0)) throw std::runtime_error("BlockCholesky: matrix not positive definite");
    d = std::sqrt(d);
    A[j * 4 + j] = d;
    for (int i = j + 1; i < 4; ++i) {
      double s = A[j * 4 + i];
      for (int k = 0; k < j; ++k) s -= A[k * 4 + i] * A[k * 4 + j];
      A[j * 4 + i] = s / d;
    }
    for (int i = 0; i < j; ++i) A[j * 4 + i] = 0.0;  // zero strict upper
  }
}
// B <- B * L^{-T}  (solve X L^T = B), L lower 4x4
static void trsmRightLT(const double *L, double *B) {
  for (int j = 0; j < 4; ++j) {
    for (int i = 0; i < 4; ++i) {
      double s = B[j * 4 + i];
      for (int k = 0; k < j; ++k) s -= B[k * 4 + i] * L[k * 4 + j];
      B[j * 4 + i] = s / L[j * 4 + j];
    }
  }
}

void BlockCholesky::factor(int n, const std::vector<std::vector<int>> &rows,
                           const std::vector<std::vector<double>> &vals) {
  n_ = n;
  // --- minimum-degree ordering on the pose graph (explicit elimination graph)
  std::vector<std::vector<int>> adj(n);
  for (int j = 0; j < n; ++j)
    for (int i : rows[j])
      if (i != j) adj[j].push_back(i);
  for (auto &v : adj) {
    std::sort(v.begin(), v.end());
    v.erase(std::unique(v.begin(), v.end()), v.end());
  }
  std::vector<char> done(n, 0);
  perm_.assign(n, 0);
  iperm_.assign(n, 0);
  std::vector<std::vector<int>> structOf(n);
  std::vector<int> tmp;
  for (int k = 0; k < n; ++k) {
    int best = -1;
    size_t bestDeg = (size_t)-1;
    for (int v = 0; v < n; ++v)
      if (!done[v] && adj[v].size() < bestDeg) {
        bestDeg = adj[v].size();
        best = v;
      }
    const int v = best;
    done[v] = 1;
    perm_[k] = v;
    iperm_[v] = k;
    structOf[v] = adj[v];
    for (int u : adj[v]) {
      // adj[u] = (adj[u] U adj[v]) \ {u, v}
      tmp.clear();
      std::set_union(adj[u].begin(), adj[u].end(), adj[v].begin(), adj[v].end(), std::back_inserter(tmp));
      tmp.erase(std::remove_if(tmp.begin(), tmp.end(), [&](int w) { return w == u || w == v; }), tmp.end());
      adj[u].swap(tmp);
    }
    adj[v].clear();
    adj[v].shrink_to_fit();
  }
  // --- permuted structure
  lrows_.assign(n, {});
  lvals_.assign(n, {});
  ldiag_.assign((size_t)n * 16, 0.0);
  for (int k = 0; k < n; ++k) {
    auto &lr = lrows_[k];
    for (int u : structOf[perm_[k]]) lr.push_back(iperm_[u]);
    std::sort(lr.begin(), lr.end());
    lvals_[k].assign(lr.size() * 16, 0.0);
  }
  // --- scatter A (lower part in permuted order)
  for (int j = 0; j < n; ++j) {
    const int pj = iperm_[j];
    for (size_t e = 0; e < rows[j].size(); ++e) {
      const int i = rows[j][e];
      const int pi = iperm_[i];
      const double *blk = &vals[j][e * 16];
      if (pi == pj) {
        for (int q = 0; q < 16; ++q) ldiag_[(size_t)pj * 16 + q] += blk[q];
      } else if (pi > pj) {
        auto &lr = lrows_[pj];
        auto it = std::lower_bound(lr.begin(), lr.end(), pi);
        assert(it != lr.end() && *it == pi);
        double *dst = &lvals_[pj][(size_t)(it - lr.begin()) * 16];
        for (int q = 0; q < 16; ++q) dst[q] += blk[q];
      }
    }
  }
  // --- right-looking numeric factorisation
  double upd[16];
  for (int k = 0; k < n; ++k) {
    double *Lkk = &ldiag_[(size_t)k * 16];
    chol44(Lkk);
    auto &lr = lrows_[k];
    auto &lv = lvals_[k];
    for (size_t e = 0; e < lr.size(); ++e) trsmRightLT(Lkk, &lv[e * 16]);
    for (size_t b = 0; b < lr.size(); ++b) {
      const int ib = lr[b];
      const double *Lb = &lv[b * 16];
      // diagonal update
      mm44_nt(Lb, Lb, upd);
      double *D = &ldiag_[(size_t)ib * 16];
      for (int q = 0; q < 16; ++q) D[q] -= upd[q];
      auto &tr = lrows_[ib];
      auto &tv = lvals_[ib];
      size_t pos = 0;
      for (size_t a = b + 1; a < lr.size(); ++a) {
        const int ia = lr[a];
        while (pos < tr.size() && tr[pos] < ia) ++pos;
        assert(pos < tr.size() && tr[pos] == ia);
        mm44_nt(&lv[a * 16], Lb, upd);  // L_ia,k * L_ib,k^T  -> block (ia, ib)
        double *dst = &tv[pos * 16];
        for (int q = 0; q < 16; ++q) dst[q] -= upd[q];
      }
    }
  }
}

size_t BlockCholesky::nnzBlocks() const {
  size_t s = n_;
  for (auto &v : lrows_) s += v.size();
  return s;
}

void BlockCholesky::solveRows(Mat &V) const {
  const int r = V.rows;
  assert(V.cols == 4 * n_);
  std::vector<double> W((size_t)n_ * 4 * r);
  for (int k = 0; k < n_; ++k) std::memcpy(&W[(size_t)k * 4 * r], V.col(4 * perm_[k]), sizeof(double) * 4 * r);
  // W block k is laid out [c][a] = (4 x r)^T i.e. element (c, a) at (k*4 + c)*r + a
  // forward: L y = b
  for (int k = 0; k < n_; ++k) {
    const double *L = &ldiag_[(size_t)k * 16];
    double *b = &W[(size_t)k * 4 * r];
    for (int c = 0; c < 4; ++c) {
      for (int q = 0; q < c; ++q) {
        const double l = L[q * 4 + c];
        for (int a = 0; a < r; ++a) b[c * r + a] -= l * b[q * r + a];
      }
      const double inv = 1.0 / L[c * 4 + c];
      for (int a = 0; a < r; ++a) b[c * r + a] *= inv;
    }
    const auto &lr = lrows_[k];
    const auto &lv = lvals_[k];
    for (size_t e = 0; e < lr.size(); ++e) {
      double *bi = &W[(size_t)lr[e] * 4 * r];
      const double *Lik = &lv[e * 16];
      for (int c = 0; c < 4; ++c)      // row c of L_ik
        for (int q = 0; q < 4; ++q) {  // col q
          const double l = Lik[q * 4 + c];
          for (int a = 0; a < r; ++a) bi[c * r + a] -= l * b[q * r + a];
        }
    }
  }
  // backward: L^T x = y
  for (int k = n_ - 1; k >= 0; --k) {
    double *b = &W[(size_t)k * 4 * r];
    const auto &lr = lrows_[k];
    const auto &lv = lvals_[k];
    for (size_t e = 0; e < lr.size(); ++e) {
      const double *xi = &W[(size_t)lr[e] * 4 * r];
      const double *Lik = &lv[e * 16];
      for (int q = 0; q < 4; ++q)      // (L_ik^T)(q, c) = L_ik(c, q)
        for (int c = 0; c < 4; ++c) {
          const double l = Lik[q * 4 + c];
          for (int a = 0; a < r; ++a) b[q * r + a] -= l * xi[c * r + a];
        }
    }
    const double *L = &ldiag_[(size_t)k * 16];
    for (int c = 3; c >= 0; --c) {
      for (int q = c + 1; q < 4; ++q) {
        const double l = L[c * 4 + q];  // L(q, c)
        for (int a = 0; a < r; ++a) b[c * r + a] -= l * b[q * r + a];
      }
      const double inv = 1.0 / L[c * 4 + c];
      for (int a = 0; a < r; ++a) b[c * r + a] *= inv;
    }
  }
  for (int k = 0; k < n_; ++k) std::memcpy(V.col(4 * perm_[k]), &W[(size_t)k * 4 * r], sizeof(double) * 4 * r);
}

// ============================================================================
// Manifold ops (a5)
// ============================================================================
// U V^T of the thin SVD of an r x 3 matrix, one-sided (Hestenes) Jacobi.
// Used by the Nesterov Y / V projections (a7).
void projectToStiefel(const double *M, int r, double *out) {
  double A[3][16];  // columns (r <= 16)
  if (r > 16) throw std::runtime_error("projectToStiefel: r > 16 unsupported");
  double V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};  // V[j] = j-th column
  for (int j = 0; j < 3; ++j)
    for (int a = 0; a < r; ++a) A[j][a] = M[(size_t)j * r + a];
  for (int sweep = 0; sweep < 60; ++sweep) {
    bool rotated = false;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        double alpha = 0, beta = 0, gamma = 0;
        for (int a = 0; a < r; ++a) {
          alpha += A[p][a] * A[p][a];
          beta += A[q][a] * A[q][a];
          gamma += A[p][a] * A[q][a];
        }
        if (std::fabs(gamma) <= 2.220446049250313e-16 * std::sqrt(alpha * beta)) continue;
        rotated = true;
        const double zeta = (beta - alpha) / (2.0 * gamma);
        const double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / std::sqrt(1.0 + t * t), s = c * t;
        for (int a = 0; a < r; ++a) {
          const double ap = A[p][a], aq = A[q][a];
          A[p][a] = c * ap - s * aq;
          A[q][a] = s * ap + c * aq;
        }
        for (int a = 0; a < 3; ++a) {
          const double vp = V[p][a], vq = V[q][a];
          V[p][a] = c * vp - s * vq;
          V[q][a] = s * vp + c * vq;
        }
      }
    if (!rotated) break;
  }
  for (int c = 0; c < 3; ++c)
    for (int a = 0; a < r; ++a) out[(size_t)c * r + a] = 0.0;
  for (int j = 0; j < 3; ++j) {
    double s = 0;
    for (int a = 0; a < r; ++a) s += A[j][a] * A[j][a];
    s = std::sqrt(s);
    if (s < 1e-300) continue;
    for (int c = 0; c < 3; ++c)
      for (int a = 0; a < r; ++a) out[(size_t)c * r + a] += (A[j][a] / s) * V[j][c];
  }
}

void manifoldProject(const Mat &M, Mat &out) {
  const int r = M.rows, n = M.cols / 4;
  if (out.rows != M.rows || out.cols != M.cols) out = Mat(M.rows, M.cols);
  for (int i = 0; i < n; ++i) {
    projectToStiefel(M.col(4 * i), r, out.col(4 * i));
    for (int a = 0; a < r; ++a) out(a, 4 * i + 3) = M(a, 4 * i + 3);
  }
}

void tangentProject(const Mat &X, const Mat &Z, Mat &out) {
  const int r = X.rows, n = X.cols / 4;
  if (out.rows != X.rows || out.cols != X.cols) out = Mat(X.rows, X.cols);
  for (int i = 0; i < n; ++i) {
    const double *Y = X.col(4 * i);
    const double *Zy = Z.col(4 * i);
    double S[3][3];
    for (int p = 0; p < 3; ++p)
      for (int q = 0; q < 3; ++q) {
        double s = 0;
        for (int a = 0; a < r; ++a) s += Y[(size_t)p * r + a] * Zy[(size_t)q * r + a];
        S[p][q] = s;
      }
    double res[3 * 16];
    for (int q = 0; q < 3; ++q)
      for (int a = 0; a < r; ++a) {
        double s = Zy[(size_t)q * r + a];
        for (int p = 0; p < 3; ++p) s -= Y[(size_t)p * r + a] * 0.5 * (S[p][q] + S[q][p]);
        res[q * r + a] = s;
      }
    double *o = out.col(4 * i);
    for (int q = 0; q < 3 * r; ++q) o[q] = res[q];
    for (int a = 0; a < r; ++a) out(a, 4 * i + 3) = Z(a, 4 * i + 3);
  }
}

// QF retraction (ROPTLIB Stiefel "params set 3": Q factor of the thin QR with
// positive diagonal R) [UPSTREAM-RECALL]; translations add.
void retract(const Mat &X, const Mat &xi, Mat &out) {
  const int r = X.rows, n = X.cols / 4;
  if (out.rows != X.rows || out.cols != X.cols) out = Mat(X.rows, X.cols);
  for (int i = 0; i < n; ++i) {
    double A[3][16];
    for (int j = 0; j < 3; ++j)
      for (int a = 0; a < r; ++a) A[j][a] = X(a, 4 * i + j) + xi(a, 4 * i + j);
    for (int j = 0; j < 3; ++j) {
      for (int pass = 0; pass < 2; ++pass)  // Gram-Schmidt with re-orthogonalisation
        for (int k = 0; k < j; ++k) {
          double s = 0;
          for (int a = 0; a < r; ++a) s += A[k][a] * A[j][a];
          for (int a = 0; a < r; ++a) A[j][a] -= s * A[k][a];
        }
      double nn = 0;
      for (int a = 0; a < r; ++a) nn += A[j][a] * A[j][a];
      nn = std::sqrt(nn);
      for (int a = 0; a < r; ++a) A[j][a] /= nn;
    }
    for (int j = 0; j < 3; ++j)
      for (int a = 0; a < r; ++a) out(a, 4 * i + j) = A[j][a];
    for (int a = 0; a < r; ++a) out(a, 4 * i + 3) = X(a, 4 * i + 3) + xi(a, 4 * i + 3);
  }
}

// ============================================================================
// PoseGraph (a4)
// ============================================================================
bool PoseGraph::hasMeasurement(int r1, int p1, int r2, int p2) const {
  return have_.count({{r1, p1}, {r2, p2}}) > 0;
}

void PoseGraph::addMeasurement(const Measurement &m) {
  if (m.r1 != id_ && m.r2 != id_) return;  // "irrelevant measurement" src/PGOAgentROS.cpp:273-275
  if (hasMeasurement(m.r1, m.p1, m.r2, m.p2)) return;
  have_.insert({{m.r1, m.p1}, {m.r2, m.p2}});
  if (m.r1 == id_ && m.r2 == id_) {
    if (m.p1 + 1 == m.p2)
      odom_.push_back(m);
    else
      privateLC_.push_back(m);
    n_ = std::max(n_, std::max(m.p1, m.p2) + 1);
  } else {
    sharedLC_.push_back(m);
    if (m.r1 == id_) {
      n_ = std::max(n_, m.p1 + 1);
      nbrs_.insert(m.r2);
    } else {
      n_ = std::max(n_, m.p2 + 1);
      nbrs_.insert(m.r1);
    }
  }
  clearDataMatrices();
}

Measurement *PoseGraph::findMeasurement(int r1, int p1, int r2, int p2) {
  for (auto *vec : {&odom_, &privateLC_, &sharedLC_})
    for (auto &m : *vec)
      if (m.r1 == r1 && m.p1 == p1 && m.r2 == r2 && m.p2 == p2) return &m;
  return nullptr;
}

std::vector<int> PoseGraph::myPublicPoseIDs(int nbr) const {
  std::set<int> s;
  for (const auto &m : sharedLC_) {
    if (m.r1 == id_ && m.r2 == nbr) s.insert(m.p1);
    if (m.r2 == id_ && m.r1 == nbr) s.insert(m.p2);
  }
  return std::vector<int>(s.begin(), s.end());
}

std::vector<PoseKey> PoseGraph::neighborPublicPoseIDs() const {
  std::set<PoseKey> s;
  for (const auto &m : sharedLC_) {
    if (m.r1 == id_)
      s.insert({m.r2, m.p2});
    else
      s.insert({m.r1, m.p1});
  }
  return std::vector<PoseKey>(s.begin(), s.end());
}

void PoseGraph::buildQ() {
  std::map<std::pair<int, int>, std::vector<double>> blk;  // (col j, row i) -> 16
  auto add = [&](int i, int j, const double *B, double sgn) {
    auto &v = blk[{j, i}];
    if (v.empty()) v.assign(16, 0.0);
    for (int q = 0; q < 16; ++q) v[q] += sgn * B[q];
  };
  double T[16], Om[4], TOm[16], TOmTt[16], OmTt[16], OmB[16];
  auto prep = [&](const Measurement &m) {
    edgeBlocks(m, T, Om);
    for (int j = 0; j < 4; ++j)
      for (int i = 0; i < 4; ++i) TOm[j * 4 + i] = T[j * 4 + i] * Om[j];   // T Omega
    mm44_nt(TOm, T, TOmTt);                                                 // T Omega T^T
    for (int j = 0; j < 4; ++j)
      for (int i = 0; i < 4; ++i) OmTt[j * 4 + i] = TOm[i * 4 + j];         // Omega T^T = (T Omega)^T
    std::memset(OmB, 0, sizeof(OmB));
    for (int i = 0; i < 4; ++i) OmB[i * 4 + i] = Om[i];
  };
  for (int i = 0; i < n_; ++i) {  // make sure every diagonal block exists
    auto &v = blk[{i, i}];
    if (v.empty()) v.assign(16, 0.0);
  }
  for (auto *vec : {&odom_, &privateLC_})
    for (const auto &m : *vec) {
      prep(m);
      add(m.p1, m.p1, TOmTt, 1.0);   // Q_ii += T Om T^T
      add(m.p2, m.p2, OmB, 1.0);     // Q_jj += Om
      add(m.p1, m.p2, TOm, -1.0);    // Q_ij  = -T Om
      add(m.p2, m.p1, OmTt, -1.0);   // Q_ji  = -Om T^T
    }
  for (const auto &m : sharedLC_) {  // only my diagonal block
    if (!neighborActive(m.r1 == id_ ? m.r2 : m.r1)) continue;
    prep(m);
    if (m.r1 == id_)
      add(m.p1, m.p1, TOmTt, 1.0);
    else
      add(m.p2, m.p2, OmB, 1.0);
  }
  qrows_.assign(n_, {});
  qvals_.assign(n_, {});
  for (auto &kv : blk) {
    const int j = kv.first.first, i = kv.first.second;
    qrows_[j].push_back(i);
    qvals_[j].insert(qvals_[j].end(), kv.second.begin(), kv.second.end());
  }
  haveQ_ = true;
  havePrecon_ = false;
}

bool PoseGraph::constructDataMatrices(const PoseDict &nbrPoses, bool needPreconditioner, double lambda) {
  if (!haveQ_) buildQ();
  if (needPreconditioner && !havePrecon_) {
    auto vals = qvals_;
    for (int j = 0; j < n_; ++j) {
      auto it = std::lower_bound(qrows_[j].begin(), qrows_[j].end(), j);
      double *d = &vals[j][(size_t)(it - qrows_[j].begin()) * 16];
      for (int c = 0; c < 4; ++c) d[c * 4 + c] += lambda;
    }
    chol_.factor(n_, qrows_, vals);
    havePrecon_ = true;
  }
  // G: rebuilt every call from the latest neighbour poses
  G_ = Mat(r_, 4 * n_);
  double T[16], Om[4], M[16];
  for (const auto &m : sharedLC_) {
    if (!neighborActive(m.r1 == id_ ? m.r2 : m.r1)) continue;
    edgeBlocks(m, T, Om);
    if (m.r1 == id_) {  // outgoing: G_i -= X_j Om T^T
      auto it = nbrPoses.find({m.r2, m.p2});
      if (it == nbrPoses.end()) return false;
      for (int j = 0; j < 4; ++j)
        for (int i = 0; i < 4; ++i) M[j * 4 + i] = Om[i] * T[i * 4 + j];  // (Om T^T)(i,j) = Om_i T(j,i)
      rmm(r_, -1.0, it->second.data(), M, G_.col(4 * m.p1));
    } else {  // incoming: G_j -= X_i T Om
      auto it = nbrPoses.find({m.r1, m.p1});
      if (it == nbrPoses.end()) return false;
      for (int j = 0; j < 4; ++j)
        for (int i = 0; i < 4; ++i) M[j * 4 + i] = T[j * 4 + i] * Om[j];
      rmm(r_, -1.0, it->second.data(), M, G_.col(4 * m.p2));
    }
  }
  return true;
}

void PoseGraph::applyQ(const Mat &X, Mat &out) const {
  if (out.rows != X.rows || out.cols != X.cols) out = Mat(X.rows, X.cols);
  out.setZero();
  const int r = X.rows;
  for (int j = 0; j < n_; ++j) {
    double *o = out.col(4 * j);
    for (size_t e = 0; e < qrows_[j].size(); ++e) rmm(r, 1.0, X.col(4 * qrows_[j][e]), &qvals_[j][e * 16], o);
  }
}

Mat PoseGraph::denseQ() const {
  Mat Q(4 * n_, 4 * n_);
  for (int j = 0; j < n_; ++j)
    for (size_t e = 0; e < qrows_[j].size(); ++e) {
      const int i = qrows_[j][e];
      for (int c = 0; c < 4; ++c)
        for (int a = 0; a < 4; ++a) Q(4 * i + a, 4 * j + c) = qvals_[j][e * 16 + c * 4 + a];
    }
  return Q;
}

// ============================================================================
// QuadraticProblem (a3)
// ============================================================================
double QuadraticProblem::f(const Mat &X) const {
  Mat XQ;
  pg_->applyQ(X, XQ);
  return 0.5 * dot(XQ, X) + dot(pg_->G(), X);
}
void QuadraticProblem::eucGrad(const Mat &X, Mat &g) const {
  pg_->applyQ(X, g);
  axpy(1.0, pg_->G(), g);
}
void QuadraticProblem::rieGrad(const Mat &X, Mat &g) const {
  Mat e;
  eucGrad(X, e);
  tangentProject(X, e, g);
}
double QuadraticProblem::rieGradNorm(const Mat &X) const {
  Mat g;
  rieGrad(X, g);
  return std::sqrt(squaredNorm(g));
}
void QuadraticProblem::rieHess(const Mat &X, const Mat &egrad, const Mat &V, Mat &out) const {
  const int r = X.rows, n = X.cols / 4;
  Mat H;
  pg_->applyQ(V, H);
  for (int i = 0; i < n; ++i) {
    const double *Y = X.col(4 * i);
    const double *E = egrad.col(4 * i);
    double S[3][3];
    for (int p = 0; p < 3; ++p)
      for (int q = 0; q < 3; ++q) {
        double s = 0;
        for (int a = 0; a < r; ++a) s += Y[(size_t)p * r + a] * E[(size_t)q * r + a];
        S[p][q] = s;
      }
    for (int q = 0; q < 3; ++q)
      for (int a = 0; a < r; ++a) {
        double s = 0;
        for (int p = 0; p < 3; ++p) s += V(a, 4 * i + p) * 0.5 * (S[p][q] + S[q][p]);
        H(a, 4 * i + q) -= s;
      }
  }
  tangentProject(X, H, out);
}
void QuadraticProblem::precondition(const Mat &X, const Mat &V, Mat &out) const {
  Mat Z = V;
  pg_->preconditioner().solveRows(Z);
  tangentProject(X, Z, out);
}

// ============================================================================
// QuadraticOptimizer (a2): selected at src/PGOAgentROSNode.cpp:82-93, params :96-100
// ============================================================================
Mat QuadraticOptimizer::optimize(const Mat &Y) {
  result_ = OptResult();
  result_.fInit = prob_->f(Y);
  result_.gradNormInit = prob_->rieGradNorm(Y);
  Mat Yopt = (params_.method == OptMethod::RTR) ? trustRegion(Y) : gradientDescent(Y);
  result_.fOpt = prob_->f(Yopt);
  result_.gradNormOpt = prob_->rieGradNorm(Yopt);
  result_.success = true;
  Mat D = Yopt;
  axpy(-1.0, Y, D);
  result_.relativeChange = std::sqrt(squaredNorm(D) / prob_->n());
  return Yopt;
}

Mat QuadraticOptimizer::gradientDescent(const Mat &Yinit) {
  Mat g, step;
  prob_->rieGrad(Yinit, g);
  if (params_.RGD_use_preconditioner) {
    Mat z;
    prob_->precondition(Yinit, g, z);
    g = z;
  }
  step = g;
  for (auto &v : step.a) v *= -params_.RGD_stepsize;
  Mat out;
  retract(Yinit, step, out);
  return out;
}

// ROPTLIB RTRNewton + tCG as dpgo configures it [UPSTREAM-RECALL]:
// accept rho > 0.1; rho < 0.25 => Delta/4; rho > 0.75 and (boundary | negative
// curvature) => min(2 Delta, Delta_max); tCG stop ||r|| <= ||r0|| min(||r0||^theta,
// kappa), theta = 1, kappa = 0.1; no randomisation; preconditioned.
Mat QuadraticOptimizer::rtrRun(const Mat &x0, int maxOuter, double initialDelta, double maxDelta,
                               bool *lastAccepted) {
  const int maxInner = params_.RTR_tCG_iterations;
  const double tol = params_.gradnorm_tol;
  const double theta = 1.0, kappa = 0.1;
  Mat x1 = x0, x2, eg1, gf1, eta2, r, z, delta, Hd, zeta, tmp;
  double f1 = prob_->f(x1);
  prob_->eucGrad(x1, eg1);
  tangentProject(x1, eg1, gf1);
  double ngf = std::sqrt(squaredNorm(gf1));
  double Delta = initialDelta;
  // ROPTLIB's Run() evaluates the stopping criterion BEFORE the first iteration (isstop = IsStopped(); while (!isstop
  // && iter < Max_Iteration)): a start whose gradient norm is already below the tolerance is returned untouched
  // ("early-return if initial ||grad|| already below tol"; round 1 always took one step)
  bool isstop = ngf < tol;
  *lastAccepted = false;
  int iter = 0;
  enum { TR_NEGCURV, TR_EXCREGION, TR_LCON, TR_SCON, TR_MAXITER } status;
  while (!isstop && iter < maxOuter) {
    // ---- tCG_TR (eta1 = 0)
    r = gf1;
    double e_Pe = 0.0;
    double r_r = squaredNorm(r);
    double norm_r = std::sqrt(r_r);
    const double norm_r0 = norm_r;
    prob_->precondition(x1, r, z);
    double z_r = dot(z, r);
    double d_Pd = z_r;
    delta = z;
    for (auto &v : delta.a) v = -v;
    double e_Pd = 0.0;
    eta2 = Mat(x1.rows, x1.cols);
    status = TR_MAXITER;
    int j = 0;
    for (j = 0; j < maxInner; ++j) {
      prob_->rieHess(x1, eg1, delta, Hd);
      const double d_Hd = dot(delta, Hd);
      const double alpha = z_r / d_Hd;
      const double e_Pe_new = e_Pe + 2.0 * alpha * e_Pd + alpha * alpha * d_Pd;
      if (d_Hd <= 0 || e_Pe_new >= Delta * Delta) {
        const double tau = (-e_Pd + std::sqrt(e_Pd * e_Pd + d_Pd * (Delta * Delta - e_Pe))) / d_Pd;
        axpy(tau, delta, eta2);
        status = (d_Hd <= 0) ? TR_NEGCURV : TR_EXCREGION;
        break;
      }
      e_Pe = e_Pe_new;
      axpy(alpha, delta, eta2);
      axpy(alpha, Hd, r);
      r_r = squaredNorm(r);
      norm_r = std::sqrt(r_r);
      const double tempnum = std::pow(norm_r0, theta);
      if (norm_r <= norm_r0 * std::min(tempnum, kappa)) {
        status = (kappa < tempnum) ? TR_LCON : TR_SCON;
        break;
      }
      prob_->precondition(x1, r, z);
      const double zold_rold = z_r;
      z_r = dot(z, r);
      const double beta = z_r / zold_rold;
      for (size_t q = 0; q < delta.a.size(); ++q) delta.a[q] = -z.a[q] + beta * delta.a[q];
      e_Pd = beta * (e_Pd + alpha * d_Pd);
      d_Pd = z_r + beta * beta * d_Pd;
    }
    result_.tcgIters += std::min(j + 1, maxInner);
    // ---- candidate + ratio
    retract(x1, eta2, x2);
    const double f2 = prob_->f(x2);
    prob_->rieHess(x1, eg1, eta2, zeta);
    tmp = gf1;
    axpy(0.5, zeta, tmp);
    const double rho = (f1 - f2) / (-dot(eta2, tmp));
    if (rho > 0.75) {
      if (status == TR_EXCREGION || status == TR_NEGCURV) Delta = std::min(2.0 * Delta, maxDelta);
    } else if (rho < 0.25) {
      Delta = 0.25 * Delta;
    }
    const bool accept =
        (rho > 0.1) || (std::fabs(f1 - f2) / (std::fabs(f1) + 1.0) < std::sqrt(2.220446049250313e-16) && f2 < f1);
    if (accept) {
      x1 = x2;
      f1 = f2;
      prob_->eucGrad(x1, eg1);
      tangentProject(x1, eg1, gf1);
      ngf = std::sqrt(squaredNorm(gf1));
      *lastAccepted = true;
    } else {
      *lastAccepted = false;
      result_.rtrRejections++;
    }
    ++iter;
    result_.rtrOuterIters++;
    isstop = ngf < tol;
  }
  return x1;
}

Mat QuadraticOptimizer::trustRegion(const Mat &Yinit) {
  const double initial = params_.RTR_initial_radius;
  if (params_.RTR_iterations == 1) {
    // single-step mode: shrink the radius until the step is accepted (a start that already meets the gradient
    // tolerance takes no step at all and comes back as it is)
    if (prob_->rieGradNorm(Yinit) < params_.gradnorm_tol) return Yinit;
    double radius = initial;
    int total = 0;
    while (true) {
      bool accepted = false;
      Mat out = rtrRun(Yinit, 1, radius, radius, &accepted);
      if (accepted) return out;
      if (total > 10) return Yinit;
      radius /= 4.0;
      ++total;
    }
  }
  bool accepted = false;
  return rtrRun(Yinit, params_.RTR_iterations, initial, 5.0 * initial, &accepted);
}

// ============================================================================
// Agent (a1, a7-a10)
// ============================================================================
Agent::Agent(int id, const Params &p) : id_(id), params_(p) {
  pg_ = std::make_shared<PoseGraph>(id, p.r, p.d);
  robust_.configure(p);
  status_.agentID = id;
}

void Agent::addMeasurement(const Measurement &m) { pg_->addMeasurement(m); }

void Agent::setLiftingMatrix(const double *Y) {
  YLift_ = Mat(params_.r, params_.d);
  std::memcpy(YLift_.a.data(), Y, sizeof(double) * params_.r * params_.d);
  haveLift_ = true;
}

void Agent::initialize(const double *T_local) {
  const int n = pg_->n();
  if (n == 0) return;
  Tlocal_ = Mat(3, 4 * n);
  if (T_local) {
    std::memcpy(Tlocal_.a.data(), T_local, sizeof(double) * 12 * n);
  } else {
    // Odometry initialisation (local_initialization_method "Odometry",
    // src/PGOAgentROSNode.cpp:106-108): chain from identity.
    std::vector<const Measurement *> bySrc(n, nullptr);
    for (const auto &m : pg_->odometry()) bySrc[m.p1] = &m;
    for (int c = 0; c < 3; ++c) Tlocal_(c, c) = 1.0;
    for (int i = 0; i + 1 < n; ++i) {
      const Measurement *m = bySrc[i];
      if (!m) throw std::runtime_error("initialize: missing odometry edge");
      // R_{i+1} = R_i R_m ; t_{i+1} = R_i t_m + t_i
      for (int c = 0; c < 3; ++c)
        for (int a = 0; a < 3; ++a) {
          double s = 0;
          for (int k = 0; k < 3; ++k) s += Tlocal_(a, 4 * i + k) * m->R[c * 3 + k];
          Tlocal_(a, 4 * (i + 1) + c) = s;
        }
      for (int a = 0; a < 3; ++a) {
        double s = Tlocal_(a, 4 * i + 3);
        for (int k = 0; k < 3; ++k) s += Tlocal_(a, 4 * i + k) * m->t[k];
        Tlocal_(a, 4 * (i + 1) + 3) = s;
      }
    }
  }
  state_ = AgentState::WAIT_FOR_INITIALIZATION;
}

void Agent::initializeChordal() {
  const int n = pg_->n();
  if (n == 0) return;
  Tlocal_ = Mat(3, 4 * n);
  for (int c = 0; c < 3; ++c) Tlocal_(c, c) = 1.0;  // pose 0: the anchor
  const int N = n - 1;                               // unknown poses 1 .. n-1 -> blocks 0 .. N-1
  if (N > 0) {
    // One block matrix serves both stages: block (i, j) = [ -k R_ij  0 ; 0  -t ] , diagonal [ sum k I3  0 ; 0  sum t ]
    // (k = w kappa, t = w tau).  Row-vector convention: unknown rows x_i (1x3 rows of R_i), cost |x_j - x_i R_ij|^2.
    std::vector<std::map<int, std::array<double, 16>>> cols(N);
    auto blk = [&](int i, int j) -> std::array<double, 16> & {
      auto it = cols[j].find(i);
      if (it == cols[j].end()) it = cols[j].emplace(i, std::array<double, 16>{}).first;
      return it->second;
    };
    Mat B1(3, 4 * N);
    std::vector<const Measurement *> edges;
    for (const auto &m : pg_->odometry()) edges.push_back(&m);
    for (const auto &m : pg_->privateLoopClosures()) edges.push_back(&m);
    for (const Measurement *m : edges) {
      const double k = m->weight * m->kappa, t = m->weight * m->tau;
      const int i = m->p1 - 1, j = m->p2 - 1;
      if (i >= 0) {
        auto &D = blk(i, i);
        D[0] += k; D[5] += k; D[10] += k; D[15] += t;
      }
      if (j >= 0) {
        auto &D = blk(j, j);
        D[0] += k; D[5] += k; D[10] += k; D[15] += t;
      }
      if (i >= 0 && j >= 0) {
        auto &U = blk(i, j), &L = blk(j, i);
        for (int c = 0; c < 3; ++c)
          for (int a = 0; a < 3; ++a) {
            U[c * 4 + a] -= k * m->R[c * 3 + a];   // (i, j) = -k R
            L[c * 4 + a] -= k * m->R[a * 3 + c];   // (j, i) = -k R^T
          }
        U[15] -= t;
        L[15] -= t;
      } else if (i < 0 && j >= 0) {  // 0 -> j : x_j should equal I * R
        for (int c = 0; c < 3; ++c)
          for (int a = 0; a < 3; ++a) B1(a, 4 * j + c) += k * m->R[c * 3 + a];
      } else if (j < 0 && i >= 0) {  // i -> 0 : x_i R should equal I
        for (int c = 0; c < 3; ++c)
          for (int a = 0; a < 3; ++a) B1(a, 4 * i + c) += k * m->R[a * 3 + c];
      }
    }
    std::vector<std::vector<int>> rows(N);
    std::vector<std::vector<double>> vals(N);
    for (int j = 0; j < N; ++j)
      for (const auto &kv : cols[j]) {
        rows[j].push_back(kv.first);
        vals[j].insert(vals[j].end(), kv.second.begin(), kv.second.end());
      }
    BlockCholesky chol;
    chol.factor(N, rows, vals);
    chol.solveRows(B1);
    for (int i = 0; i < N; ++i) {
      double M[9], Rp[9];
      for (int c = 0; c < 3; ++c)
        for (int a = 0; a < 3; ++a) M[c * 3 + a] = B1(a, 4 * i + c);
      projectToStiefel(M, 3, Rp);  // polar factor of the 3x3 block
      for (int c = 0; c < 3; ++c)
        for (int a = 0; a < 3; ++a) Tlocal_(a, 4 * (i + 1) + c) = Rp[c * 3 + a];
    }
    // translations: sum t |p_j - p_i - R_i t_ij|^2, p_0 = 0
    Mat B2(3, 4 * N);
    for (const Measurement *m : edges) {
      const double t = m->weight * m->tau;
      const int i = m->p1 - 1, j = m->p2 - 1;
      double v[3];
      for (int a = 0; a < 3; ++a) {
        double sum = 0;
        for (int kk = 0; kk < 3; ++kk) sum += Tlocal_(a, 4 * m->p1 + kk) * m->t[kk];
        v[a] = t * sum;
      }
      for (int a = 0; a < 3; ++a) {
        if (j >= 0) B2(a, 4 * j + 3) += v[a];
        if (i >= 0) B2(a, 4 * i + 3) -= v[a];
      }
    }
    chol.solveRows(B2);
    for (int i = 0; i < N; ++i)
      for (int a = 0; a < 3; ++a) Tlocal_(a, 4 * (i + 1) + 3) = B2(a, 4 * i + 3);
  }
  state_ = AgentState::WAIT_FOR_INITIALIZATION;
}

void Agent::initializeInGlobalFrame(const double *Tw) {
  if (state_ == AgentState::WAIT_FOR_DATA) throw std::runtime_error("initializeInGlobalFrame before initialize");
  if (!haveLift_) throw std::runtime_error("initializeInGlobalFrame: lifting matrix not set");
  const int n = pg_->n(), r = params_.r;
  X_ = Mat(r, 4 * n);
  for (int i = 0; i < n; ++i) {
    double Tg[12];  // 3x4 column-major: R = Rw Ri ; t = Rw ti + tw
    for (int c = 0; c < 4; ++c)
      for (int a = 0; a < 3; ++a) {
        double s = (c == 3) ? Tw[9 + a] : 0.0;
        for (int k = 0; k < 3; ++k) s += Tw[k * 3 + a] * Tlocal_(k, 4 * i + c);
        Tg[c * 3 + a] = s;
      }
    for (int c = 0; c < 4; ++c)
      for (int a = 0; a < r; ++a) {
        double s = 0;
        for (int k = 0; k < 3; ++k) s += YLift_(a, k) * Tg[c * 3 + k];
        X_(a, 4 * i + c) = s;
      }
  }
  Xinit_ = X_;
  Xprev_ = X_;
  state_ = AgentState::INITIALIZED;
  status_.state = state_;
  if (params_.acceleration) initializeAcceleration();
}

void Agent::reset() {
  // PGOAgent::reset() as called from src/PGOAgentROS.cpp:223 -- back to
  // WAIT_FOR_DATA, new instance, measurements kept unless the wrapper swaps
  // mPoseGraph (:237).
  instance_++;
  iter_ = 0;
  state_ = AgentState::WAIT_FOR_DATA;
  status_ = Status();
  status_.agentID = id_;
  status_.instanceNumber = instance_;
  teamStatus_.clear();
  for (int rid : std::set<int>(inactive_)) setRobotActive(rid, true);
  nbrPoses_.clear();
  nbrAuxPoses_.clear();
  weightUpdateCount_ = 0;
  robustInnerIter_ = 0;
  robust_.reset();
  gamma_ = alpha_ = 0;
  optResult_ = OptResult();
}

void Agent::initializeAcceleration() {
  V_ = X_;
  Y_ = X_;
  gamma_ = 0;
  alpha_ = 0;
}
// Nesterov sequences of RBCD++; N = number of robots
void Agent::updateGamma() {
  const double N = params_.numRobots;
  gamma_ = (1.0 + std::sqrt(1.0 + 4.0 * N * N * gamma_ * gamma_)) / (2.0 * N);
}
void Agent::updateAlpha() { alpha_ = 1.0 / (gamma_ * params_.numRobots); }
void Agent::updateY() {
  Mat M = X_;
  for (size_t q = 0; q < M.a.size(); ++q) M.a[q] = (1.0 - alpha_) * X_.a[q] + alpha_ * V_.a[q];
  manifoldProject(M, Y_);
}
void Agent::updateV() {
  Mat M = V_;
  for (size_t q = 0; q < M.a.size(); ++q) M.a[q] = V_.a[q] + gamma_ * (X_.a[q] - Y_.a[q]);
  manifoldProject(M, V_);
}
bool Agent::shouldRestart() const {
  if (params_.acceleration) return ((iter_ + 1) % params_.restartInterval) == 0;
  return false;
}
void Agent::restartNesterov(bool doOptimization) {
  if (params_.acceleration && state_ == AgentState::INITIALIZED) {
    X_ = Xprev_;
    updateX(doOptimization, false);
    V_ = X_;
    Y_ = X_;
    gamma_ = 0;
    alpha_ = 0;
  }
}

bool Agent::updateX(bool doOptimization, bool acceleration) {
  if (!doOptimization) {
    if (acceleration) X_ = Y_;
    return true;
  }
  const PoseDict &nbr = acceleration ? nbrAuxPoses_ : nbrPoses_;
  const bool needPre = (params_.method == OptMethod::RTR) || params_.RGD_use_preconditioner;
  if (!pg_->constructDataMatrices(nbr, needPre, params_.precondLambda)) return false;
  QuadraticProblem problem(pg_.get(), params_.r);
  QuadraticOptimizer optimizer(&problem, params_);
  const Mat &Xinit = acceleration ? Y_ : X_;
  X_ = optimizer.optimize(Xinit);
  optResult_ = optimizer.result();
  return true;
}

// PGOAgent::iterate -- src/PGOAgentROS.cpp:160 (true) and :1185 (false)
bool Agent::iterate(bool doOptimization) {
  iter_++;
  if (params_.costType != CostType::L2) robustInnerIter_++;
  bool success = false;
  if (state_ == AgentState::INITIALIZED) {
    Xprev_ = X_;
    if (params_.acceleration) {
      updateGamma();
      updateAlpha();
      updateY();
      success = updateX(doOptimization, true);
      updateV();
      if (shouldRestart()) restartNesterov(doOptimization);
    } else {
      success = updateX(doOptimization, false);
    }
    if (doOptimization) {
      Mat D = X_;
      axpy(-1.0, Xprev_, D);
      status_.relativeChange = std::sqrt(squaredNorm(D) / pg_->n());
      bool ready = success;
      if (status_.relativeChange > params_.relChangeTol) ready = false;
      // robustOptMinConvergenceRatio (src/PGOAgentROSNode.cpp:214; launch default 0, launch/PGOAgent.launch:34): not
      // ready while too few loop-closure weights have settled at 0 or 1 [UPSTREAM-RECALL; counted like the
      // statistics the TERMINATE handler prints, src/PGOAgentROS.cpp:1058-1067]
      if (params_.robustOptMinConvergenceRatio > 0.0) {
        size_t total = 0, settled = 0;
        for (auto *vec : {&pg_->privateLoopClosures(), &pg_->sharedLoopClosures()})
          for (const auto &m : *vec) {
            ++total;
            if (m.weight == 1.0 || m.weight == 0.0) ++settled;
          }
        if (total > 0 && (double)settled < params_.robustOptMinConvergenceRatio * (double)total) ready = false;
      }
      status_.readyToTerminate = ready;
    }
  }
  status_.agentID = id_;
  status_.state = state_;
  status_.instanceNumber = instance_;
  status_.iterationNumber = iter_;
  teamStatus_[id_] = status_;
  return success;
}

std::vector<int> Agent::getNeighbors() const {
  return std::vector<int>(pg_->neighbors().begin(), pg_->neighbors().end());
}

static bool packDict(const PoseGraph &pg, int id, const Mat &X, int nbr, PoseDict &out) {
  const int r = X.rows;
  for (int f : pg.myPublicPoseIDs(nbr)) {
    std::vector<double> v(X.col(4 * f), X.col(4 * f) + 4 * r);
    out[{id, f}] = std::move(v);
  }
  return true;
}
bool Agent::getSharedPoseDictWithNeighbor(PoseDict &out, int nbr) const {
  if (state_ != AgentState::INITIALIZED) return false;
  return packDict(*pg_, id_, X_, nbr, out);
}
bool Agent::getAuxSharedPoseDictWithNeighbor(PoseDict &out, int nbr) const {
  if (state_ != AgentState::INITIALIZED || !params_.acceleration) return false;
  return packDict(*pg_, id_, Y_, nbr, out);
}
void Agent::updateNeighborPoses(int nbr, const PoseDict &poses) {
  for (const auto &kv : poses)
    if (kv.first.robot == nbr) nbrPoses_[kv.first] = kv.second;
}
void Agent::updateAuxNeighborPoses(int nbr, const PoseDict &poses) {
  for (const auto &kv : poses)
    if (kv.first.robot == nbr) nbrAuxPoses_[kv.first] = kv.second;
}

Status Agent::getStatus() const {
  Status s = status_;
  s.agentID = id_;
  s.state = state_;
  s.instanceNumber = instance_;
  s.iterationNumber = iter_;
  return s;
}

// [UPSTREAM-RECALL] leader-side tests evaluated at src/PGOAgentROS.cpp:208,210
bool Agent::shouldTerminate() const {
  if (iter_ > params_.maxNumIters) return true;
  if (params_.costType != CostType::L2 && weightUpdateCount_ < params_.robustOptNumWeightUpdates) return false;
  for (int rid = 0; rid < params_.numRobots; ++rid) {
    if (inactive_.count(rid)) continue;
    auto it = teamStatus_.find(rid);
    if (it == teamStatus_.end()) return false;
    if (it->second.state != AgentState::INITIALIZED) return false;
    if (!it->second.readyToTerminate) return false;
  }
  return true;
}
bool Agent::shouldUpdateMeasurementWeights() const {
  if (params_.costType == CostType::L2) return false;
  if (weightUpdateCount_ >= params_.robustOptNumWeightUpdates) return false;
  if (robustInnerIter_ >= params_.robustOptInnerIters) return true;
  // otherwise only when every robot reports readyToTerminate
  for (int rid = 0; rid < params_.numRobots; ++rid) {
    if (inactive_.count(rid)) continue;
    auto it = teamStatus_.find(rid);
    if (it == teamStatus_.end()) return false;
    if (it->second.state != AgentState::INITIALIZED || !it->second.readyToTerminate) return false;
  }
  return true;
}

bool Agent::computeMeasurementResidual(const Measurement &m, double *residual) const {
  if (state_ != AgentState::INITIALIZED) return false;
  const int r = params_.r;
  const double *Xi = nullptr, *Xj = nullptr;
  if (m.r1 == id_)
    Xi = X_.col(4 * m.p1);
  else {
    auto it = nbrPoses_.find({m.r1, m.p1});
    if (it == nbrPoses_.end()) return false;
    Xi = it->second.data();
  }
  if (m.r2 == id_)
    Xj = X_.col(4 * m.p2);
  else {
    auto it = nbrPoses_.find({m.r2, m.p2});
    if (it == nbrPoses_.end()) return false;
    Xj = it->second.data();
  }
  double rot = 0, tr = 0;
  for (int c = 0; c < 3; ++c)
    for (int a = 0; a < r; ++a) {
      double s = -Xj[(size_t)c * r + a];
      for (int k = 0; k < 3; ++k) s += Xi[(size_t)k * r + a] * m.R[c * 3 + k];
      rot += s * s;
    }
  for (int a = 0; a < r; ++a) {
    double s = Xj[(size_t)3 * r + a] - Xi[(size_t)3 * r + a];
    for (int k = 0; k < 3; ++k) s -= Xi[(size_t)k * r + a] * m.t[k];
    tr += s * s;
  }
  *residual = std::sqrt(m.kappa * rot + m.tau * tr);
  return true;
}

// PGOAgent::updateMeasurementWeights -- src/PGOAgentROS.cpp:1218.  Shared edges
// are re-weighted by the lower-ID robot only (ownership rule visible in the
// wrapper at :732 and :1340); the higher-ID robot receives the weight.
void Agent::updateMeasurementWeights() {
  if (state_ != AgentState::INITIALIZED) return;
  double res = 0;
  for (auto &m : pg_->privateLoopClosures()) {
    if (m.fixedWeight) continue;
    if (computeMeasurementResidual(m, &res)) m.weight = robust_.weight(res);
  }
  for (auto &m : pg_->sharedLoopClosures()) {
    if (m.fixedWeight) continue;
    const int other = (m.r1 == id_) ? m.r2 : m.r1;
    if (other < id_) continue;
    if (computeMeasurementResidual(m, &res)) m.weight = robust_.weight(res);
  }
  robust_.update();
  weightUpdateCount_++;
  robustInnerIter_ = 0;
  pg_->clearDataMatrices();
  if (weightUpdateCount_ <= params_.robustOptNumResets) X_ = Xinit_;
  if (params_.acceleration) initializeAcceleration();
}

bool Agent::setMeasurementWeight(int r1, int p1, int r2, int p2, double w, bool fixed) {
  Measurement *m = pg_->findMeasurement(r1, p1, r2, p2);
  if (!m) return false;
  m->weight = w;
  m->fixedWeight = fixed;
  return true;
}

// ============================================================================
// Team: in-process replay of the synchronous protocol
// ============================================================================
Team::Team(const Params &p) : params_(p) {
  for (int i = 0; i < p.numRobots; ++i) agents_.emplace_back(new Agent(i, p));
}

void Team::deliver(int from) {
  Agent &a = *agents_[from];
  for (int nbr : a.getNeighbors()) {
    PoseDict d;
    if (a.getSharedPoseDictWithNeighbor(d, nbr)) agents_[nbr]->updateNeighborPoses(from, d);
    if (params_.acceleration) {
      PoseDict da;
      if (a.getAuxSharedPoseDictWithNeighbor(da, nbr)) agents_[nbr]->updateAuxNeighborPoses(from, da);
    }
  }
}

void Team::exchangeAll() {
  for (int i = 0; i < size(); ++i) deliver(i);
}

namespace {
// minimal spinning fork-join pool: one OS thread per agent
class SpinPool {
 public:
  explicit SpinPool(int nthreads) : n_(nthreads) {
    for (int t = 1; t < n_; ++t) workers_.emplace_back([this, t] { loop(t); });
  }
  ~SpinPool() {
    stop_.store(true);
    gen_.fetch_add(1);
    for (auto &w : workers_) w.join();
  }
  template <class F>
  void parallelFor(int count, F &&fn) {
    if (n_ <= 1) {
      for (int i = 0; i < count; ++i) fn(i);
      return;
    }
    fn_ = [&](int i) { fn(i); };
    count_ = count;
    next_.store(0);
    pending_.store(n_ - 1);
    gen_.fetch_add(1);
    work();
    while (pending_.load() != 0) {
    }
  }

 private:
  void work() {
    for (;;) {
      int i = next_.fetch_add(1);
      if (i >= count_) break;
      fn_(i);
    }
  }
  void loop(int) {
    unsigned seen = 0;
    for (;;) {
      unsigned g;
      while ((g = gen_.load()) == seen) {
      }
      seen = g;
      if (stop_.load()) return;
      work();
      pending_.fetch_sub(1);
    }
  }
  int n_;
  std::vector<std::thread> workers_;
  std::function<void(int)> fn_;
  int count_ = 0;
  std::atomic<int> next_{0}, pending_{0};
  std::atomic<unsigned> gen_{0};
  std::atomic<bool> stop_{false};
};
}  // namespace

TeamRunResult Team::run(int maxIters, int numThreads, bool stopOnTerminate) {
  TeamRunResult res;
  const int N = size();
  SpinPool pool(std::max(1, std::min(numThreads, N)));
  auto t0 = std::chrono::high_resolution_clock::now();
  for (int it = 0; it < maxIters; ++it) {
    const int sel = selected_;
    // non-selected robots iterate(false) immediately (src/PGOAgentROS.cpp:1185)
    if (params_.acceleration) {
      pool.parallelFor(N, [&](int a) {
        if (a != sel) agents_[a]->iterate(false);
      });
      for (int a = 0; a < N; ++a)
        if (a != sel) deliver(a);  // mPublishPublicPosesRequested (:109-113)
    } else {
      for (int a = 0; a < N; ++a)
        if (a != sel) agents_[a]->iterate(false);
    }
    // selected robot: gate satisfied (:136-149) -> iterate(true) (:160)
    agents_[sel]->iterate(true);
    deliver(sel);
    // publishStatus (:183, :1186): everyone hears everyone
    for (int a = 0; a < N; ++a) {
      const Status s = agents_[a]->getStatus();
      for (int b = 0; b < N; ++b)
        if (b != a) agents_[b]->setNeighborStatus(s);
    }
    res.iterations++;
    selected_ = (sel + 1) % N;  // RoundRobin (:464-472)
    if (sel == 0) {             // leader decides (:207-217)
      if (agents_[0]->shouldTerminate()) {
        res.terminated = true;
        if (stopOnTerminate) break;
      } else if (agents_[0]->shouldUpdateMeasurementWeights()) {
        // UPDATE_WEIGHT (:1211-1233)
        for (int a = 0; a < N; ++a) agents_[a]->updateMeasurementWeights();
        // publishMeasurementWeights (:721-754) -> measurementWeightsCallback (:1315-1353)
        for (int a = 0; a < N; ++a)
          for (const auto &m : agents_[a]->poseGraph().sharedLoopClosures()) {
            const int other = (m.r1 == a) ? m.r2 : m.r1;
            if (other > a) {
              if (agents_[other]->setMeasurementWeight(m.r1, m.p1, m.r2, m.p2, m.weight, m.fixedWeight))
                agents_[other]->poseGraph().clearDataMatrices();
            }
          }
        exchangeAll();
        res.weightUpdates++;
      }
    }
  }
  res.wallSeconds = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
  return res;
}

TeamRunResult Team::runParallel(int ticks, int numThreads) {
  TeamRunResult res;
  const int N = size();
  if (params_.acceleration) throw std::runtime_error("the asynchronous schedule runs without acceleration");
  SpinPool pool(std::max(1, std::min(numThreads, N)));
  auto t0 = std::chrono::high_resolution_clock::now();
  for (int it = 0; it < ticks; ++it) {
    pool.parallelFor(N, [&](int a) { agents_[a]->iterate(true); });
    for (int a = 0; a < N; ++a) deliver(a);
    for (int a = 0; a < N; ++a) {
      const Status s = agents_[a]->getStatus();
      for (int b = 0; b < N; ++b)
        if (b != a) agents_[b]->setNeighborStatus(s);
    }
    res.iterations++;
  }
  res.wallSeconds = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
  return res;
}

double Team::globalCost() const {
  double cost = 0;
  const int r = params_.r;
  auto edgeCost = [&](const Measurement &m, const double *Xi, const double *Xj) {
    double rot = 0, tr = 0;
    for (int c = 0; c < 3; ++c)
      for (int a = 0; a < r; ++a) {
        double s = -Xj[(size_t)c * r + a];
        for (int k = 0; k < 3; ++k) s += Xi[(size_t)k * r + a] * m.R[c * 3 + k];
        rot += s * s;
      }
    for (int a = 0; a < r; ++a) {
      double s = Xj[(size_t)3 * r + a] - Xi[(size_t)3 * r + a];
      for (int k = 0; k < 3; ++k) s -= Xi[(size_t)k * r + a] * m.t[k];
      tr += s * s;
    }
    return m.weight * (m.kappa * rot + m.tau * tr);
  };
  for (int a = 0; a < size(); ++a) {
    Agent &ag = *agents_[a];
    PoseGraph &pg = ag.poseGraph();
    for (auto *vec : {&pg.odometry(), &pg.privateLoopClosures()})
      for (const auto &m : *vec) cost += edgeCost(m, ag.X().col(4 * m.p1), ag.X().col(4 * m.p2));
    for (const auto &m : pg.sharedLoopClosures())
      if (m.r1 == a) cost += edgeCost(m, ag.X().col(4 * m.p1), agents_[m.r2]->X().col(4 * m.p2));
  }
  return cost;
}

}  // namespace dpgo_oracle
