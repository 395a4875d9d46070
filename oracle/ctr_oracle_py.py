"""Second, independent restatement of the reference FK in pure Python.  TEST INFRASTRUCTURE ONLY.

Purpose: a second pin of oracle/ctr_oracle.c next to the reference build (oracle/_ref, see
ctr_oracle.h).  The reference ships no golden vectors; this file was written from the mathematics of the reference rather than from the C code:
per-tube Python lists, 3x3 row-major matrices + translation vectors instead of flat column-major
arrays, an explicit list of samples instead of in-place slots.  Two independently written
restatements agreeing to ~1e-15 is the evidence that both follow robot_i.h / section_i.h.

Reference lines followed: robot_i.h:690-755 (drivers), :1074-1122 (sweep set-up and loop),
:1139-1238 (one sample), :1241-1285 (SE(3) step and Rz), :1286-1307 (radius),
section_i.h:187-326 (per-section arithmetic).

Pure-Python loops: use for small cases only (a few hundred configurations).
"""
import math

MODE_STEP, MODE_NSAMPLE = 0, 1


class Section:
    def __init__(self, stiffness, curvature, length, radius, poisson):
        self.ntube = len(stiffness)
        self.stiffness = list(stiffness)
        self.curvature = list(curvature)
        self.length = float(length)
        self.radius = list(radius)
        self.poisson = list(poisson)

    @property
    def stiffness_sum(self):            # matrix.h:178-188, left to right
        acc = self.stiffness[0]
        for v in self.stiffness[1:]:
            acc += v
        return acc


class RobotModel:
    """sections: list[Section]; init_R: 3x3 row-major nested list; init_p: 3-list."""

    def __init__(self, sections, init_R=None, init_p=None, max_samples=256, zero_tol=1e-12):
        self.sections = sections
        self.init_R = init_R or [[1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]]
        self.init_p = init_p or [0.0, 0.0, 0.0]
        self.max_samples = int(max_samples)
        self.zero_tol = zero_tol
        self.ntubes = sum(s.ntube for s in sections)
        self.first_tube = []
        acc = 0
        for s in sections:
            self.first_tube.append(acc)
            acc += s.ntube

    @property
    def njoints(self):
        return 1 + len(self.sections) + self.ntubes


def _matmul3(A, B):
    return [[A[i][0] * B[0][j] + A[i][1] * B[1][j] + A[i][2] * B[2][j] for j in range(3)]
            for i in range(3)]


def _matvec3(A, v):
    return [A[i][0] * v[0] + A[i][1] * v[1] + A[i][2] * v[2] for i in range(3)]


def _compose(Ra, pa, Rb, pb):
    q = _matvec3(Ra, pb)
    return _matmul3(Ra, Rb), [q[0] + pa[0], q[1] + pa[1], q[2] + pa[2]]


def _arc_step(u, h, zero_tol):
    """robot_i.h:1241-1269"""
    n = math.sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2])
    if abs(n) < zero_tol:
        return [[1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]], [0.0, 0.0, h * n]
    w0, w1, w2 = u[0] / n, u[1] / n, u[2] / n
    s, c = math.sin(h * n), math.cos(h * n)
    v = c - 1.0
    R = [[(w1 * w1 + w2 * w2) * v + 1.0, -s * w2 - w0 * w1 * v, s * w1 - w0 * w2 * v],
         [s * w2 - w0 * w1 * v, (w0 * w0 + w2 * w2) * v + 1.0, -s * w0 - w1 * w2 * v],
         [-s * w1 - w0 * w2 * v, s * w0 - w1 * w2 * v, (w0 * w0 + w1 * w1) * v + 1.0]]
    p = [h * w0 * w2 + (-(w1 * (R[0][0] - 1.0)) + (R[0][1] * w0)) / n,
         h * w1 * w2 + ((w0 * (R[1][1] - 1.0)) - (R[1][0] * w1)) / n,
         h * w2 * w2 + ((R[2][1] * w0) - (R[2][0] * w1)) / n]
    return R, p


def _rotz_left(R, p, sn, cn):
    """robot_i.h:1280-1285"""
    R2 = [[R[0][j] * cn - R[1][j] * sn for j in range(3)],
          [R[1][j] * cn + R[0][j] * sn for j in range(3)],
          [R[2][j] for j in range(3)]]
    p2 = [p[0] * cn - p[1] * sn, p[1] * cn + p[0] * sn, p[2]]
    return R2, p2


def forward_kinematics(model, jv, mode, step=None, nsamples=None):
    """Returns dict(frames=[(R,p)...], radius=[...], n=int, step=float, alpha_base=[...],
    curv=[[ux,uy,uz]...], alpha=[...])."""
    secs = model.sections
    S, T = len(secs), model.ntubes
    theta = jv[0]
    alpha = [0.0] * T
    ends, starts = [], []
    run = 0.0
    for si, sec in enumerate(secs):
        t0 = model.first_tube[si]
        phi = jv[1 + si + t0]
        if not (0.0 <= phi <= sec.length):
            raise ValueError("phi out of range")
        for k in range(sec.ntube):
            alpha[t0 + k] = jv[2 + si + t0 + k]
        run += phi
        ends.append(run)
        starts.append(max(0.0, run - sec.length))
    arc_end = run

    if mode == MODE_STEP:
        h = step
        count = max(1, int(arc_end / h))
        if count > model.max_samples:
            count = model.max_samples
            h = arc_end / float(count)
    else:
        count = min(int(nsamples), model.max_samples)
        h = arc_end / float(count)

    ux = [0.0] * T
    uy = [0.0] * T
    uz = [0.0] * T
    samp_u = [[0.0, 0.0, 0.0] for _ in range(count)]
    samp_a = [0.0] * count

    def xy_at(pos):
        sn = [math.sin(a) for a in alpha]
        cs = [math.cos(a) for a in alpha]
        kap = [float(pos <= ends[i] and pos >= starts[i]) * secs[i].curvature[0] for i in range(S)]
        kst = [float(pos <= ends[i]) * secs[i].stiffness_sum for i in range(S)]
        cx = cy = ktot = 0.0
        for i, sec in enumerate(secs):
            ktot += kst[i]
            for k in range(sec.ntube):
                t = model.first_tube[i] + k
                cx += -sn[t] * kap[i] * (kst[i] / sec.ntube)
                cy += cs[t] * kap[i] * (kst[i] / sec.ntube)
        for t in range(T):
            ux[t] = (cs[t] * cx + sn[t] * cy) / ktot
            uy[t] = (-sn[t] * cx + cs[t] * cy) / ktot
        return kap, kst

    pos = arc_end
    idx = count - 1
    samp_u[idx][2] = 0.0
    while pos >= h and idx >= 1:
        kap, kst = xy_at(pos)
        samp_u[idx][0], samp_u[idx][1] = ux[T - 1], uy[T - 1]
        # angles (old torsion), section_i.h:289-312
        z0 = uz[0]
        new_alpha = list(alpha)
        new_alpha[0] = 0.0
        for t in range(1, T):
            new_alpha[t] = alpha[t] + h * (uz[t] - z0)
        alpha = new_alpha
        samp_a[idx] = alpha[T - 1]
        # torsion, section_i.h:243-286
        cz = 0.0
        last = S - 1
        order = []
        if secs[last].ntube == 2:
            order.append((last, model.first_tube[last], secs[last].poisson[1]))
        for i in range(S - 2, -1, -1):
            for k in range(secs[i].ntube - 1, -1, -1):
                order.append((i, model.first_tube[i] + k, secs[i].poisson[k]))
        for i, t, nu in order:
            cz += (kst[i] / secs[i].ntube) / (1.0 + nu) * uz[t]
            uz[t] += h * (1.0 + nu) * ux[t] * kap[i]
        uz[T - 1] = (-cz) * (1.0 + secs[last].poisson[secs[last].ntube - 1]) / kst[last]
        samp_u[idx - 1][2] = uz[T - 1]
        pos -= h
        idx -= 1
    xy_at(pos)
    samp_u[idx][0], samp_u[idx][1] = ux[T - 1], uy[T - 1]
    samp_a[idx] = alpha[T - 1]

    frames = [None] * model.max_samples
    R = [[1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]]
    p = [0.0, 0.0, 0.0]
    while True:
        if mode == MODE_STEP:
            if not (pos <= arc_end and idx < model.max_samples):
                break
        else:
            if not (idx < nsamples and idx < model.max_samples):
                break
        Re, pe = _arc_step(samp_u[idx], h, model.zero_tol)
        R, p = _compose(R, p, Re, pe)
        a = samp_a[idx] + theta
        Rz, pz = _rotz_left(R, p, math.sin(a), math.cos(a))
        frames[idx] = _compose(model.init_R, model.init_p, Rz, pz)
        if mode == MODE_STEP:
            pos += h
        idx += 1
    n = idx

    radius = [0.0] * n
    begin = 0
    rsec = 0.0
    for i, sec in enumerate(secs):
        if begin >= n:
            continue
        rsec = sec.radius[0]
        if h > 0.0 and math.isfinite(ends[i] / h) and ends[i] / h < n:
            stop = int(ends[i] / h)
        else:
            stop = n
        stop = max(stop, begin)
        for k in range(begin, stop):
            radius[k] = rsec
        begin = stop
    radius[n - 1] = rsec

    return dict(frames=frames[:n], radius=radius, n=n, step=h, alpha_base=list(alpha),
                curv=samp_u, alpha=samp_a)


def frame_to_colmajor(frame):
    R, p = frame
    return [R[0][0], R[1][0], R[2][0], R[0][1], R[1][1], R[2][1], R[0][2], R[1][2], R[2][2],
            p[0], p[1], p[2]]


def model_from_desc(desc):
    """desc: the ctypes ctr_robot_desc mirror from ctrfk.py (duck-typed)."""
    secs = []
    for i in range(desc.nsections):
        sd = desc.sec[i]
        n = sd.ntube
        secs.append(Section([sd.stiffness[k] for k in range(n)], [sd.curvature[k] for k in range(n)],
                            sd.length, [sd.radius[k] for k in range(n)],
                            [sd.poisson[k] for k in range(n)]))
    t = list(desc.init_trafo)
    R = [[t[0], t[3], t[6]], [t[1], t[4], t[7]], [t[2], t[5], t[8]]]
    p = [t[9], t[10], t[11]]
    ztol = desc.zero_tol if desc.zero_tol > 0 else 1e-12
    return RobotModel(secs, R, p, desc.max_samples, ztol)
