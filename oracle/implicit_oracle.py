"""TEST INFRASTRUCTURE -- CPU restatement of the reference's implicit ARD branch
(PD_ARD_ImplicitSolver, src/pd_ard_implicit.cpp) in numpy / scipy.  Never imported by the product.

PARITY STATUS: pinned on the reference's compiled code except for Eigen's internal arithmetic.  The reference
solves (I - dt M) C = b with Eigen 3.4.0 (GMRES<SparseMatrix, IncompleteLUT>, tolerance 1e-10, restart 50, <= 200
iterations, src/pd_ard_implicit.cpp:384-409); Eigen is fetched at configure time (CMakeLists.txt:27-38) and is absent
from the reference tree and from this image.  This file restates the reference's own code around that call --
operator assembly, boundary right-hand side, adaptive time step, write-back clamp -- line by line, and replaces the
iterative solve by scipy's sparse direct solve, i.e. the exact solution Eigen's GMRES approximates to its 1e-10
relative residual.  tests/test_reference_implicit.py checks it against the UNMODIFIED src/pd_ard_implicit.cpp compiled
against the Eigen work-alike oracle/eigen_min/ (oracle/_ref/libpdrefimp{2,3}d.so): system matrix and right-hand side
to 1e-14, adaptive step 1e-12, step results to the solve tolerance, and the whole implicit run of the reference's
main() row for row.  Unpinned: Eigen's floating-point order inside GMRES / ILUT (acts below 1e-10).

    assemble()            src/pd_ard_implicit.cpp:104-346 (M over the FLUID + SOLID_MG unknowns, BC weights)
    bc_rhs()              :352-362
    step()                :371-429  (A = I - dt M, b = C_old + dt bc_rhs, solve, clamp to [0, C_solid_init])
    adaptive_dt()         :438-487

Uniform grid only (no FICTITIOUS nodes: use_amr = 0).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

FLUID, SOLID, WALL, INLET, OUTLET, OUTSIDE = range(6)


class ImplicitOracle:
    def __init__(self, dim, Nx, Ny, Nz, node_type, off_d, off_dist, off_evec, off_vol, cfg):
        self.dim, self.Nx, self.Ny, self.Nz = dim, Nx, Ny, Nz
        self.nt = np.asarray(node_type, np.uint8)
        self.N = self.nt.size
        self.off_d = np.asarray(off_d, np.int64).reshape(-1, 3)
        self.dist, self.vol = np.asarray(off_dist, float), np.asarray(off_vol, float)
        self.evec = np.asarray(off_evec, float).reshape(-1, dim)
        self.cfg = cfg
        pi = 3.14159265358979323846
        self.alpha_p = float(dim)                                   # :11-20
        if dim == 2:
            self.V_H = pi * cfg.delta * cfg.delta
            self.beta = 4.0 / (pi * cfg.delta * cfg.delta)
        else:
            self.V_H = (4.0 / 3.0) * pi * cfg.delta * cfg.delta * cfg.delta
            self.beta = 12.0 / (pi * cfg.delta * cfg.delta)
        n = np.arange(self.N)
        self.k = n // (Nx * Ny)
        rem = n % (Nx * Ny)
        self.j = rem // Nx
        self.i = rem % Nx
        self.volume_loss = 0.0

    def _nbr(self, o):
        """(valid mask, neighbour index) of offset o for every node; valid = in the box and not OUTSIDE
        (what the reference's CSR holds, src/grid.cpp:194-227)."""
        di, dj, dk = self.off_d[o]
        ni, nj, nk = self.i + di, self.j + dj, self.k + dk
        ok = (ni >= 0) & (ni < self.Nx) & (nj >= 0) & (nj < self.Ny) & (nk >= 0) & (nk < self.Nz)
        nn = np.where(ok, (nk * self.Ny + nj) * self.Nx + ni, 0)
        ok &= self.nt[nn] != OUTSIDE
        return ok, nn

    def assemble(self, C, vel, is_gb, is_precip):
        cfg, nt = self.cfg, self.nt
        C = np.asarray(C, float)
        vel = np.asarray(vel, float).reshape(self.N, self.dim)
        unknown = (nt == FLUID) | (nt == SOLID)                      # :44-62
        self.l2g = np.nonzero(unknown)[0]
        self.g2l = np.full(self.N, -1, np.int64)
        self.g2l[self.l2g] = np.arange(self.l2g.size)
        nu = self.l2g.size
        # salt-layer blocking (:70-89): a solid with ANY FLUID neighbour at C >= C_sat
        salt = np.zeros(self.N, bool)
        for o in range(len(self.off_d)):
            ok, nn = self._nbr(o)
            salt |= (nt == SOLID) & ok & (nt[nn] == FLUID) & (C[nn] >= cfg.C_sat)
        self.salt = salt
        decay = 1.0                                                  # :127-132
        if cfg.corrosion_decay_l > 0.0:
            decay = 10.0 ** (-self.volume_loss / cfg.corrosion_decay_l)
        D_s_node = np.where(np.asarray(is_gb) != 0, cfg.D_gb, np.where(np.asarray(is_precip) != 0, cfg.D_precip, cfg.D_grain)) * decay
        D_if = 2.0 * cfg.D_liquid * D_s_node / (cfg.D_liquid + D_s_node + 1e-30)     # harmonic mean, :229-231
        D_if = np.where(salt, 0.0, D_if)
        div_coeff = self.alpha_p / self.V_H
        i_fl, i_so = nt == FLUID, nt == SOLID
        rows, cols, vals = [], [], []
        diag = np.zeros(self.N)
        bc_k, bc_j, bc_w = [], [], []
        for o in range(len(self.off_d)):
            ok, nn = self._nbr(o)
            ntj = nt[nn]
            act = unknown & ok & (ntj != WALL)                        # :196
            j_fl = (ntj == FLUID) | (ntj == INLET) | (ntj == OUTLET)
            j_so = ntj == SOLID
            act &= ~(i_so & j_so)                                     # :212
            inv_xi = 1.0 / self.dist[o]
            inv_xi2 = inv_xi * inv_xi
            D_avg = np.zeros(self.N)
            D_avg = np.where(i_fl & j_fl, cfg.D_liquid, D_avg)        # :215-217
            D_avg = np.where(i_fl & j_so, D_if[nn], D_avg)            # :218-230 (solid j)
            D_avg = np.where(i_so & j_fl, D_if, D_avg)                # :231-243 (solid i)
            w_diff = self.beta * D_avg * inv_xi2 * self.vol[o]        # :267
            w = w_diff.copy()
            ll = i_fl & j_fl                                          # advection on liquid-liquid bonds, :272-282
            v_dot_e = vel @ self.evec[o]
            w_adv = div_coeff * v_dot_e * inv_xi * self.vol[o]
            w_stab = np.maximum(0.0, w_adv - w_diff)
            w = np.where(ll, (w_diff + w_stab) - w_adv, w)
            diag -= np.where(act, w, 0.0)                             # :284
            offd = act & ((ntj == FLUID) | (ntj == SOLID))            # :287-293
            rows.append(self.g2l[np.nonzero(offd)[0]])
            cols.append(self.g2l[nn[offd]])
            vals.append(w[offd])
            bc = act & ((ntj == INLET) | (ntj == OUTLET))             # :294-297
            bc_k.append(self.g2l[np.nonzero(bc)[0]])
            bc_j.append(nn[bc])
            bc_w.append(w[bc])
        rows.append(np.arange(nu)); cols.append(np.arange(nu)); vals.append(diag[self.l2g])   # :301
        self.M = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(nu, nu))
        self.bc_k, self.bc_j, self.bc_w = np.concatenate(bc_k), np.concatenate(bc_j), np.concatenate(bc_w)
        return self.M

    def bc_rhs(self, C):                                              # :352-362
        out = np.zeros(self.l2g.size)
        np.add.at(out, self.bc_k, self.bc_w * np.asarray(C, float)[self.bc_j])
        return out

    def system(self, C, dt):
        """A = I - dt M, b = C_old + dt bc_rhs over the unknowns (:380-394)"""
        A = sp.identity(self.l2g.size, format="csr") - dt * self.M
        b = np.asarray(C, float)[self.l2g] + dt * self.bc_rhs(C)
        return A, b

    def step(self, C, dt):
        """exact solve + clamp to [0, C_solid_init] (:371-429); returns the new global C array"""
        A, b = self.system(C, dt)
        x = spla.spsolve(A.tocsc(), b)
        out = np.array(C, float, copy=True)
        out[self.l2g] = np.clip(x, 0.0, self.cfg.C_solid_init)
        return out

    def adaptive_dt(self, C, dt_fraction, dt_max):                   # :438-487
        C = np.asarray(C, float)
        MC = self.M @ C[self.l2g] + self.bc_rhs(C)
        dCdt = np.zeros(self.N)
        dCdt[self.l2g] = MC
        m = (self.nt == SOLID) & (C > self.cfg.C_thresh) & (dCdt < 0.0) & (-dCdt >= 1e-30)
        t_phase = (C[m] - self.cfg.C_thresh) / (-dCdt[m])
        t_phase = t_phase[t_phase > 0.0]
        min_t = min(dt_max, float(t_phase.min())) if t_phase.size else dt_max
        dt = dt_fraction * min_t
        dt = min(dt, dt_max)
        return max(dt, dt_max * 0.01)
