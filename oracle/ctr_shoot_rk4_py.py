"""Checker of the RK4 option of the base-angle shooting solve (CTR_FK_FLAG_RK4).  TEST INFRASTRUCTURE ONLY.

The reference has no shooting solve and no Runge-Kutta integrator (SURVEY.md "READ FIRST"): its sweep
(robot_i.h:1101-1122, section_i.h:243-312) is explicit Euler, tip -> base, on the torsional rod equations of
SURVEY.md §9.1.  CTR_FK_FLAG_RK4 integrates the SAME equations with the classical fourth-order scheme; there is no
reference output to compare it with, so its checker is this file, written from the equations and independent of
csrc/ctr_fk_core.cuh (numpy vectors, a generic right-hand side, derivatives by complex step instead of hand-written
variational equations), plus the convergence property that ties it to the reference: as h -> 0 the Euler sweep
(the oracle, pinned against the reference) and the RK4 sweep agree to O(h).

State, with sigma the arc length from the tip and all angles relative to tube 0 (alpha_0 = 0):
    d alpha_t / d sigma = z_t - z_0                              t = 1..T-1
    d z_t     / d sigma = (1 + nu_t) kappa_sec(t) x_t            t = 0..T-2, inside the curved interval of its section
    z_{T-1}             = -(1 + nu_last) / K_last  *  sum_{t<T-1} (K_sec(t)/n_sec(t)) / (1 + nu_t) * z_t
    x_t = (cos alpha_t cx + sin alpha_t cy) / K_active,
    (cx, cy) = sum_t' kappa_sec(t') K_sec(t')/n_sec(t') (-sin alpha_t', cos alpha_t')   over curved sections
with the reference's effective constants: a section's tubes share its stiffness sum equally (section_i.h:189), the
first tube of a two-tube last section uses the second tube's Poisson ratio (section_i.h:273), a section is "active"
while pos <= End and "curved" while Start <= pos <= End (robot_i.h:1139-1160), the last section is always active.
Sample count and step as robot_i.h:1079-1097; N-1 steps from pos = ArcEnd, a step being split wherever a section's
curved interval starts or ends inside it (the right-hand side jumps there).
"""
import numpy as np

from ctr_oracle_py import MODE_STEP, MODE_NSAMPLE  # noqa: F401


def _control(model, jv, mode, step, nsamples):
    S = len(model.sections)
    ends, starts, alpha = [], [], np.zeros(model.ntubes, dtype=complex)
    run = 0.0
    for si, sec in enumerate(model.sections):
        t0 = model.first_tube[si]
        phi = float(np.real(jv[1 + si + t0]))
        if not (0.0 <= phi <= sec.length):
            raise ValueError("phi out of range")
        for k in range(sec.ntube):
            alpha[t0 + k] = jv[2 + si + t0 + k]
        run += phi
        ends.append(run)
        starts.append(max(0.0, run - sec.length))
    if mode == MODE_STEP:
        h = step
        n = max(1, int(run / h))
        if n > model.max_samples:
            n = model.max_samples
            h = run / float(n)
    else:
        n = min(int(nsamples), model.max_samples)
        h = run / float(n)
    assert S == len(ends)
    return alpha, ends, starts, run, n, h


def _constants(model):
    secs = model.sections
    S, T = len(secs), model.ntubes
    sec_of = [si for si, s in enumerate(secs) for _ in range(s.ntube)]
    share = [s.stiffness_sum / s.ntube for s in secs]
    onu = []
    for si, s in enumerate(secs):
        for k in range(s.ntube):
            nu = s.poisson[k]
            if si == S - 1 and s.ntube == 2 and k == 0:
                nu = s.poisson[1]
            onu.append(1.0 + nu)
    last = secs[-1]
    cl = (1.0 + last.poisson[last.ntube - 1]) / last.stiffness_sum
    assert len(onu) == T
    return sec_of, share, onu, cl


def rhs(model, consts, ends, starts, pos, y):
    """y = (alpha_1..alpha_{T-1}, z_0..z_{T-2}) -> dy/dsigma at arc position pos."""
    secs = model.sections
    S, T = len(secs), model.ntubes
    U = T - 1
    sec_of, share, onu, cl = consts
    p = max(pos, 0.0)
    active = [(si == S - 1) or (p <= ends[si]) for si in range(S)]
    curved = [active[si] and (starts[si] <= p) for si in range(S)]
    K = sum(secs[si].stiffness_sum for si in range(S) if active[si])
    alpha = np.concatenate(([0.0 + 0.0j], y[:U]))
    z = list(y[U:])
    zi = -cl * sum(share[sec_of[t]] / onu[t] * z[t] for t in range(U))
    z.append(zi)
    sn, cs = np.sin(alpha), np.cos(alpha)
    cx = sum(-sn[t] * secs[sec_of[t]].curvature[0] * share[sec_of[t]] for t in range(T) if curved[sec_of[t]]) / K
    cy = sum(cs[t] * secs[sec_of[t]].curvature[0] * share[sec_of[t]] for t in range(T) if curved[sec_of[t]]) / K
    dy = np.zeros(2 * U, dtype=complex)
    for t in range(1, T):
        dy[t - 1] = z[t] - z[0]
    for t in range(U):
        if curved[sec_of[t]]:
            dy[U + t] = onu[t] * secs[sec_of[t]].curvature[0] * (cs[t] * cx + sn[t] * cy)
    return dy


def base_angles(model, jv, mode, step=None, nsamples=None, refine=1, split=True):
    """Base angles (relative to tube 0; entry 0 is 0) of the RK4 sweep from the tip angles in jv.  jv may be
    complex: the result then carries the complex-step derivative information.  refine > 1 integrates the SAME arc
    interval [h, ArcEnd] with every step cut into `refine` equal parts (convergence studies); split=False steps
    across the section boundaries without splitting (to show what the splitting buys)."""
    alpha, ends, starts, run, n, h = _control(model, jv, mode, step, nsamples)
    T = model.ntubes
    U = T - 1
    if U == 0:
        return np.zeros(1), n, h
    consts = _constants(model)
    y = np.concatenate((alpha[1:], np.zeros(U, dtype=complex)))
    # The right-hand side jumps at the section boundaries; a step containing one is split there (each part with the
    # interval membership of its own midpoint), otherwise the scheme would drop to first order.
    marks = sorted(set(ends) | set(starts))
    pos = run
    for _ in range(n - 1):
        stop = pos - h
        cuts = [pos] + ([m for m in reversed(marks) if stop < m < pos] if split else []) + [stop]
        if refine > 1:
            cuts = [a - (a - b) * i / refine for a, b in zip(cuts, cuts[1:]) for i in range(refine)] + [stop]
        for hi, lo in zip(cuts, cuts[1:]):
            hs = hi - lo
            mid = hi - 0.5 * hs
            k1 = rhs(model, consts, ends, starts, mid, y)
            k2 = rhs(model, consts, ends, starts, mid, y + 0.5 * hs * k1)
            k3 = rhs(model, consts, ends, starts, mid, y + 0.5 * hs * k2)
            k4 = rhs(model, consts, ends, starts, mid, y + hs * k3)
            y = y + (hs / 6.0) * (k1 + 2.0 * k2 + 2.0 * k3 + k4)
        pos = stop
    return np.concatenate(([0.0 + 0.0j], y[:U])), n, h


def alpha_slots(model):
    return [2 + si + model.first_tube[si] + k for si, s in enumerate(model.sections) for k in range(s.ntube)]


def base_angles_and_jacobian(model, jv, mode, step=None, nsamples=None):
    """Real base angles and d base_{1..} / d tip_{1..} by complex step (exact to rounding: the map is analytic
    away from the interval switches, which depend on the section lengths only)."""
    idx = alpha_slots(model)
    b, n, h = base_angles(model, np.asarray(jv, dtype=complex), mode, step, nsamples)
    T = model.ntubes
    J = np.zeros((T - 1, T - 1))
    eps = 1e-30
    for c in range(T - 1):
        q = np.asarray(jv, dtype=complex).copy()
        q[idx[c + 1]] += 1j * eps
        bc, _, _ = base_angles(model, q, mode, step, nsamples)
        J[:, c] = np.imag(bc[1:]) / eps
    return np.real(b), J


def shoot(model, jv_base, guess=None, mode=MODE_NSAMPLE, step=None, nsamples=None, tol=1e-11, max_iter=40):
    """Plain Newton on base(x) - beta = 0 from `guess` (default x = beta).  Returns (jv_tip, residual, sweeps)."""
    idx = alpha_slots(model)
    q = np.array(jv_base, dtype=float)
    beta = np.array([jv_base[i] for i in idx])
    if guess is not None:
        for t in range(1, len(idx)):
            q[idx[t]] = guess[idx[t]]
    res = np.inf
    for it in range(1, max_iter + 1):
        b, J = base_angles_and_jacobian(model, q, mode, step, nsamples)
        f = b[1:] - beta[1:]
        res = float(np.abs(f).max()) if f.size else 0.0
        if res <= tol:
            return q, res, it
        d = np.clip(np.linalg.solve(J, f), -1.0, 1.0)
        for t in range(1, len(idx)):
            q[idx[t]] -= d[t - 1]
    return q, res, max_iter
