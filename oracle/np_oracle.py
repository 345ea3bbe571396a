"""Second, independent restatement of the hot path in numpy (dense Q) --
TEST INFRASTRUCTURE ONLY.  It exists to cross-check oracle/dpgo_oracle.cpp on
small cases (tinyGrid3D, smallGrid3D): SURVEY §8c asks for two independent
restatements agreeing to 1e-10 on f / grad / Hess-vec and 1e-8 on iterates.

Independent choices on purpose: dense Q assembled from the cost definition by
explicit Kronecker-style accumulation, LAPACK SVD for the Stiefel projection,
LAPACK QR for the retraction, LAPACK solve for the preconditioner.

Conventions: X is r x 4n; pose i = columns [4i, 4i+4) = [Y_i | p_i].
Reference anchors: src/PGOAgentROS.cpp:160,1185 (iterate), :1276-1278 (neighbour
poses), src/PGOAgentROSNode.cpp:82-100 (solver selection and parameters).
"""
from __future__ import annotations

import numpy as np


def edge_T_Omega(R, t, kappa, tau, w):
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = t
    Om = np.diag([w * kappa, w * kappa, w * kappa, w * tau])
    return T, Om


class NpAgent:
    """One robot's quadratic problem: f(X) = 0.5 <Q, X^T X> + <G, X>."""

    def __init__(self, rid: int, meas, n: int, r: int, lam: float = 0.1):
        self.id, self.n, self.r, self.lam = rid, n, r, lam
        self.meas = meas  # Measurements SoA restricted to this robot
        self.Q = None
        self.G = None

    def build_Q(self):
        n = self.n
        Q = np.zeros((4 * n, 4 * n))
        m = self.meas
        for e in range(len(m)):
            T, Om = edge_T_Omega(m.R[e], m.t[e], m.kappa[e], m.tau[e], m.weight[e])
            i, j = int(m.p1[e]), int(m.p2[e])
            si, sj = slice(4 * i, 4 * i + 4), slice(4 * j, 4 * j + 4)
            if m.r1[e] == self.id and m.r2[e] == self.id:
                Q[si, si] += T @ Om @ T.T
                Q[sj, sj] += Om
                Q[si, sj] += -T @ Om
                Q[sj, si] += -Om @ T.T
            elif m.r1[e] == self.id:
                Q[si, si] += T @ Om @ T.T
            elif m.r2[e] == self.id:
                Q[sj, sj] += Om
        self.Q = Q
        self.P = Q + self.lam * np.eye(4 * n)
        return Q

    def build_G(self, nbr_pose):
        """nbr_pose: dict (robot, frame) -> r x 4 array."""
        G = np.zeros((self.r, 4 * self.n))
        m = self.meas
        for e in range(len(m)):
            if m.r1[e] == m.r2[e]:
                continue
            T, Om = edge_T_Omega(m.R[e], m.t[e], m.kappa[e], m.tau[e], m.weight[e])
            if m.r1[e] == self.id:
                Xj = nbr_pose[(int(m.r2[e]), int(m.p2[e]))]
                i = int(m.p1[e])
                G[:, 4 * i:4 * i + 4] -= Xj @ Om @ T.T
            else:
                Xi = nbr_pose[(int(m.r1[e]), int(m.p1[e]))]
                j = int(m.p2[e])
                G[:, 4 * j:4 * j + 4] -= Xi @ T @ Om
        self.G = G
        return G

    # ---- problem
    def f(self, X):
        return 0.5 * np.sum((X @ self.Q) * X) + np.sum(self.G * X)

    def egrad(self, X):
        return X @ self.Q + self.G

    def rgrad(self, X):
        return tangent_project(X, self.egrad(X))

    def rhess(self, X, V):
        eg = self.egrad(X)
        H = V @ self.Q
        for i in range(self.n):
            s = slice(4 * i, 4 * i + 3)
            S = X[:, s].T @ eg[:, s]
            H[:, s] -= V[:, s] @ (0.5 * (S + S.T))
        return tangent_project(X, H)

    def precond(self, X, V):
        Z = np.linalg.solve(self.P, V.T).T
        return tangent_project(X, Z)

    # ---- local solvers
    def rgd(self, X, stepsize, use_precond):
        g = self.rgrad(X)
        if use_precond:
            g = self.precond(X, g)
        return retract(X, -stepsize * g)

    def rtr(self, X0, max_outer=3, max_inner=50, tol=1e-2, Delta0=100.0):
        """ROPTLIB-style RTRNewton with preconditioned tCG [UPSTREAM-RECALL], written
        from the Absil/Baker/Gallivan formulation, not from the C++ oracle."""
        x = X0.copy()
        fx = self.f(x)
        g = self.rgrad(x)
        Delta, Dmax = Delta0, 5 * Delta0
        it = 0
        stop = False
        inner_total = 0
        while not stop and it < max_outer:
            eta = np.zeros_like(x)
            r = g.copy()
            nr0 = np.linalg.norm(r)
            z = self.precond(x, r)
            zr = np.sum(z * r)
            d = -z
            ePe, ePd, dPd = 0.0, 0.0, zr
            hit = False
            for j in range(max_inner):
                inner_total += 1
                Hd = self.rhess(x, d)
                dHd = np.sum(d * Hd)
                alpha = zr / dHd
                ePe_new = ePe + 2 * alpha * ePd + alpha * alpha * dPd
                if dHd <= 0 or ePe_new >= Delta * Delta:
                    tau = (-ePd + np.sqrt(ePd * ePd + dPd * (Delta * Delta - ePe))) / dPd
                    eta = eta + tau * d
                    hit = True
                    break
                ePe = ePe_new
                eta = eta + alpha * d
                r = r + alpha * Hd
                nr = np.linalg.norm(r)
                if nr <= nr0 * min(nr0, 0.1):
                    break
                z = self.precond(x, r)
                zr_old = zr
                zr = np.sum(z * r)
                beta = zr / zr_old
                d = -z + beta * d
                ePd = beta * (ePd + alpha * dPd)
                dPd = zr + beta * beta * dPd
            x2 = retract(x, eta)
            f2 = self.f(x2)
            Heta = self.rhess(x, eta)
            rho = (fx - f2) / (-np.sum(eta * (g + 0.5 * Heta)))
            if rho > 0.75:
                if hit:
                    Delta = min(2 * Delta, Dmax)
            elif rho < 0.25:
                Delta = 0.25 * Delta
            if rho > 0.1 or (abs(fx - f2) / (abs(fx) + 1) < np.sqrt(np.finfo(float).eps) and f2 < fx):
                x, fx = x2, f2
                g = self.rgrad(x)
            it += 1
            stop = np.linalg.norm(g) < tol
        return x, inner_total


def tangent_project(X, Z):
    out = Z.copy()
    n = X.shape[1] // 4
    for i in range(n):
        s = slice(4 * i, 4 * i + 3)
        Y = X[:, s]
        S = Y.T @ Z[:, s]
        out[:, s] = Z[:, s] - Y @ (0.5 * (S + S.T))
    return out


def manifold_project(M):
    out = M.copy()
    n = M.shape[1] // 4
    for i in range(n):
        s = slice(4 * i, 4 * i + 3)
        U, _, Vt = np.linalg.svd(M[:, s], full_matrices=False)
        out[:, s] = U @ Vt
    return out


def retract(X, xi):
    out = X + xi
    n = X.shape[1] // 4
    for i in range(n):
        s = slice(4 * i, 4 * i + 3)
        Qf, Rf = np.linalg.qr(out[:, s])
        sg = np.sign(np.diag(Rf))
        sg[sg == 0] = 1.0
        out[:, s] = Qf * sg[None, :]
    return out


class NpTeam:
    """Synchronous RoundRobin RBCD / RBCD++ over all robots (numpy restatement)."""

    def __init__(self, problem, ylift, r, method="RGD", stepsize=0.2, use_precond=True, acceleration=False,
                 restart_interval=50, rtr_iterations=3, rtr_tcg=50, gradnorm_tol=1e-2, lam=0.1):
        self.N = problem.num_robots
        self.r = r
        self.method, self.stepsize, self.use_precond = method, stepsize, use_precond
        self.accel, self.restart_interval = acceleration, restart_interval
        self.rtr_iterations, self.rtr_tcg, self.gradnorm_tol = rtr_iterations, rtr_tcg, gradnorm_tol
        self.agents = []
        self.X = []
        for rid in range(self.N):
            m = problem.robot_measurements(rid)
            ag = NpAgent(rid, m, problem.n[rid], r, lam)
            ag.build_Q()
            self.agents.append(ag)
            T = problem.T_init[rid]  # [n,3,4] global frame
            X = np.zeros((r, 4 * problem.n[rid]))
            for i in range(problem.n[rid]):
                X[:, 4 * i:4 * i + 4] = ylift @ T[i]
            self.X.append(X)
        self.V = [x.copy() for x in self.X]
        self.Y = [x.copy() for x in self.X]
        self.gamma = [0.0] * self.N
        self.alpha = [0.0] * self.N
        self.iter = 0
        self.selected = 0
        self.rel_change = [None] * self.N

    def _nbr(self, rid, src):
        ag = self.agents[rid]
        d = {}
        m = ag.meas
        for e in range(len(m)):
            if m.r1[e] == m.r2[e]:
                continue
            if m.r1[e] == rid:
                b, f = int(m.r2[e]), int(m.p2[e])
            else:
                b, f = int(m.r1[e]), int(m.p1[e])
            d[(b, f)] = src[b][:, 4 * f:4 * f + 4]
        return d

    def _solve(self, rid, Xstart, src):
        ag = self.agents[rid]
        ag.build_G(self._nbr(rid, src))
        if self.method == "RGD":
            return ag.rgd(Xstart, self.stepsize, self.use_precond)
        x, _ = ag.rtr(Xstart, self.rtr_iterations, self.rtr_tcg, self.gradnorm_tol)
        return x

    def step(self):
        """One global iteration: non-selected iterate(false) first, then the selected robot."""
        N = self.N
        sel = self.selected
        self.iter += 1
        restart = self.accel and ((self.iter + 1) % self.restart_interval == 0)
        Xprev = [x.copy() for x in self.X]
        order = [a for a in range(N) if a != sel] + [sel]
        for a in order:
            opt = a == sel
            if self.accel:
                g = self.gamma[a]
                g = (1 + np.sqrt(1 + 4 * N * N * g * g)) / (2 * N)
                self.gamma[a] = g
                self.alpha[a] = 1.0 / (g * N)
                al = self.alpha[a]
                self.Y[a] = manifold_project((1 - al) * self.X[a] + al * self.V[a])
                if opt:
                    self.X[a] = self._solve(a, self.Y[a], self.Y)
                else:
                    self.X[a] = self.Y[a].copy()
                self.V[a] = manifold_project(self.V[a] + g * (self.X[a] - self.Y[a]))
                if restart:
                    self.X[a] = Xprev[a].copy()
                    if opt:
                        self.X[a] = self._solve(a, self.X[a], self.X)
                    self.V[a] = self.X[a].copy()
                    self.Y[a] = self.X[a].copy()
                    self.gamma[a] = 0.0
                    self.alpha[a] = 0.0
            elif opt:
                self.X[a] = self._solve(a, self.X[a], self.X)
            if opt:
                self.rel_change[a] = np.sqrt(np.sum((self.X[a] - Xprev[a]) ** 2) / self.agents[a].n)
        self.selected = (sel + 1) % N
