"""Second, independent restatement of the reference's per-ray path -- pure Python floats.

TEST INFRASTRUCTURE.  Written from the Fortran sources separately from oracle/ort_oracle.cpp
(different language, different structure: plain tuples and scalar code, no shared helpers) so
that a transcription slip in either shows up as a disagreement (tests/test_pyref.py).  It takes
its uniforms from a caller-supplied function, so it carries no generator of its own.

Python floats are IEEE doubles and `math` calls the same libm as the C++ oracle, so agreement
is expected to the last few ulps.  Cited lines are in the reference tree.
"""
import math

PI = 4.0 * math.atan(1.0)          # src/constants.f90:5
TWOPI = 2.0 * 4.0 * math.atan(1.0)


def _sub(a, b):
    return (a[0] - b[0], a[1] - b[1], a[2] - b[2])


def _add(a, b):
    return (a[0] + b[0], a[1] + b[1], a[2] + b[2])


def _scale(a, k):
    return (a[0] * k, a[1] * k, a[2] * k)


def _dot(a, b):
    return (a[0] * b[0]) + (a[1] * b[1]) + (a[2] * b[2])


def _unit(a):                       # vector%magnitude(), src/vector_class.f90:175-186
    t = math.sqrt(a[0] ** 2 + a[1] ** 2 + a[2] ** 2)
    return (a[0] / t, a[1] / t, a[2] / t)


def _sqrt(x):
    return math.sqrt(x) if x >= 0.0 else float("nan")


def solve_quadratic(a, b, c):       # src/surfaces.f90:227-260
    disc = b ** 2 - 4.0 * a * c
    if disc < 0.0:
        return None
    if disc == 0.0:
        x0 = -0.5 * b / a
        return x0, x0
    if b > 0.0:
        q = -0.5 * (b + math.sqrt(disc))
    else:
        q = -0.5 * (b - math.sqrt(disc))
    return q / a, c / q


def _first_hit(roots):              # tail of every intersect_*, e.g. src/surfaces.f90:74-87
    if roots is None:
        return None
    t0, t1 = roots
    if t0 > t1:
        t0, t1 = t1, t0
    if t0 < 0.0:
        t0 = t1
        if t0 < 0.0:
            return None
    return t0


def hit_sphere(o, d, c, R):         # src/surfaces.f90:52-89
    L = _sub(o, c)
    return _first_hit(solve_quadratic(_dot(d, d), 2.0 * _dot(d, L), _dot(L, L) - R ** 2))


def hit_cylinder(o, d, c, R):       # src/surfaces.f90:91-130
    L = _sub(o, c)
    a = d[2] ** 2 + d[1] ** 2
    b = 2 * (d[2] * L[2] + d[1] * L[1])
    cc = L[2] ** 2 + L[1] ** 2 - R ** 2
    return _first_hit(solve_quadratic(a, b, cc))


def hit_ellipse(o, d, c, sa, sb):   # src/surfaces.f90:133-176
    ia = 1. / sa ** 2
    ib = 1. / sb ** 2
    L = _sub(o, c)
    a = ia * d[2] ** 2 + ib * d[1] ** 2
    b = 2 * (ia * d[2] * L[2] + ib * d[1] * L[1])
    cc = ia * L[2] ** 2 + ib * L[1] ** 2 - 1
    return _first_hit(solve_quadratic(a, b, cc))


def fresnel(I, N, n1, n2):          # src/surfaces.f90:336-372
    costt = abs(_dot(I, N))
    sintt = _sqrt(1. - costt * costt)
    sint2 = n1 / n2 * sintt
    if sint2 > 1.:
        return 1.0
    if costt == 1.:
        return 0.
    sint2 = (n1 / n2) * sintt
    cost2 = _sqrt(1. - sint2 * sint2)
    try:
        f1 = abs((n1 * costt - n2 * cost2) / (n1 * costt + n2 * cost2)) ** 2
        f2 = abs((n1 * cost2 - n2 * costt) / (n1 * cost2 + n2 * costt)) ** 2
    except ZeroDivisionError:
        return 1.
    tir = 0.5 * (f1 + f2)
    if math.isnan(tir) or tir > 1. or tir < 0.:
        tir = 1.
    return tir


def reflect_refract(I, N, n1, n2, u):   # src/surfaces.f90:262-333 -> (new I, reflected?)
    if u <= fresnel(I, N, n1, n2):
        k = 2. * _dot(N, I)
        return _sub(I, _scale(N, k)), True
    eta = n1 / n2
    Nt = N
    c1 = _dot(Nt, I)
    if c1 < 0.:
        c1 = -c1
    else:
        Nt = _scale(N, -1.)
    c2 = _sqrt(1.0 - eta ** 2 * (1.0 - c1 ** 2))
    return _add(_scale(I, eta), _scale(Nt, eta * c1 - c2)), False


def source_point(cos_theta_max, offset, u):     # src/sourceMod.f90:12-47
    phi = TWOPI * u(1)   # slot 1 (narrow draw); cos theta takes the wide slot 0
    cosp, sinp = math.cos(phi), math.sin(phi)
    ran = u(0)
    cost = (1.0 - ran) + ran * cos_theta_max
    sint = math.sqrt(1.0 - cost ** 2)
    return (0.0, 0.0, 0.0 + offset), (sint * cosp, sint * sinp, cost)


def source_ring(S, u):                          # src/sourceMod.f90:250-300
    Ra, Rb = S.bottle.radiusa, S.bottle.radiusb
    r = S.r1 + u(0) * (S.r2 - S.r1)
    theta = u(1) * TWOPI
    px = math.sqrt(r) * math.cos(theta)
    py = math.sqrt(r) * math.sin(theta)
    if S.bottle.ellipse:
        pz = S.bottle.centre[2] + math.sqrt(Ra ** 2 - (py * Ra / Rb) ** 2)
    else:
        pz = S.bottle.centre[2] + math.sqrt(Ra ** 2 - py ** 2)
    r = 0. + u(2) * ((S.L2.radius + 10e-3) ** 2 - 0.)
    theta = u(3) * TWOPI
    lx = math.sqrt(r) * math.cos(theta)
    ly = math.sqrt(r) * math.sin(theta)
    lz = S.L2.fb
    dist = math.sqrt((lx - px) ** 2 + (ly - py) ** 2 + (lz - pz) ** 2)
    d = ((lx - px) / dist, (ly - py) / dist, (lz - pz) / dist)
    return (px, py, pz), _unit(d)


def bottle_forward(S, pos, d, u, fix_outer=False):   # src/lens.f90:230-350, clear bottles only
    B = S.bottle
    c = tuple(B.centre)
    if B.ellipse:
        t = hit_ellipse(pos, d, c, B.radiusa - B.thickness, B.radiusb - B.thickness)
    else:
        t = hit_cylinder(pos, d, c, B.radiusa - B.thickness)
    if t is None:
        return pos, d, 1
    pos = _add(pos, _scale(d, t))
    normal = _unit(_sub(c, (c[0], pos[1], pos[2])))
    d, refl = reflect_refract(d, normal, B.ncontents, B.nbottle, u(2))
    if refl:
        return pos, d, 4
    if B.ellipse:
        if fix_outer:
            t = hit_ellipse(pos, d, c, B.radiusa, B.radiusb)
        else:
            t = hit_ellipse(pos, d, c, B.radiusa / 2., B.radiusb / 2.)
    else:
        t = hit_cylinder(pos, d, c, B.radiusa)
    if t is None:
        return pos, d, 5
    pos = _add(pos, _scale(d, t))
    normal = _unit(_sub(c, (c[0], pos[1], pos[2])))
    d, refl = reflect_refract(d, normal, B.nbottle, 1.0, u(3))
    if refl:
        return pos, d, 8
    return pos, d, 0


def plano_forward(L, pos, d, u):    # src/lens.f90:425-481
    c = tuple(L.centre)
    a = c[2] + L.curve_radius - L.thickness
    k = (a - pos[2]) / d[2] if d[2] != 0.0 else math.copysign(math.inf, a - pos[2])
    pos = _add(pos, _scale(d, k))
    if math.sqrt(pos[0] ** 2 + pos[1] ** 2) > L.radius:
        return pos, d, 9
    d, _ = reflect_refract(d, tuple(L.flat_normal), L.n1, L.n2, u(4))   # flag ignored :458-459
    t = hit_sphere(pos, d, c, L.curve_radius)
    if t is None:
        return pos, d, 10
    pos = _add(pos, _scale(d, t))
    d, refl = reflect_refract(d, _unit(_sub(c, pos)), L.n2, L.n1, u(5))
    return pos, d, (11 if refl else 0)


def doublet_forward(L, pos, d, iris1, iris2, iris_radius, u):   # src/lens.f90:531-645
    c1, c2, c3 = tuple(L.centre1), tuple(L.centre2), tuple(L.centre3)
    if iris1:
        t = ((c1[2] - L.R1) - pos[2]) / d[2]
        p = _add(pos, _scale(d, t))
        if math.sqrt(p[0] ** 2 + p[1] ** 2) > L.radius * iris_radius:
            return p, d, 12
    t = hit_sphere(pos, d, c1, L.R1)
    if t is None:
        return pos, d, 13
    pos = _add(pos, _scale(d, t))
    if math.sqrt(pos[0] ** 2 + pos[1] ** 2) > (L.radius * 1.0):
        return pos, d, 14
    d, refl = reflect_refract(d, _unit(_sub(pos, c1)), L.n1, L.n2, u(6))
    if refl:
        return pos, d, 15
    t = hit_sphere(pos, d, c2, L.R2)
    if t is None:
        return pos, d, 16
    pos = _add(pos, _scale(d, t))
    d, refl = reflect_refract(d, _unit(_sub(c2, pos)), L.n2, L.n3, u(7))
    if refl:
        return pos, d, 17
    t = hit_sphere(pos, d, c3, L.R3)
    if t is None:
        return pos, d, 18
    pos = _add(pos, _scale(d, t))
    d, refl = reflect_refract(d, _unit(_sub(c3, pos)), L.n3, L.n1, u(8))
    if refl:
        return pos, d, 19
    if iris2:
        t = ((c3[2] + L.R3) - pos[2]) / d[2]
        p = _add(pos, _scale(d, t))
        if math.sqrt(p[0] ** 2 + p[1] ** 2) > L.radius * iris_radius:
            return p, d, 20
    return pos, d, 0


def make_image(d, pos, diameter):   # src/imageMod.f90:19-58 -> (status, xp, yp)
    n = _unit((0., 0., -1.))
    dd = _scale(_unit(d), -1.)
    top = _dot(n, dd)
    bottom = math.sqrt(_dot(dd, dd)) * math.sqrt(_dot(n, n))
    q = top / bottom
    angle = math.acos(q) if -1.0 <= q <= 1.0 else float("nan")
    if angle > math.asin(0.22):
        return 21, None, None
    binwid = diameter / 401.
    if pos[0] > 1000 or pos[1] > 1000:
        return 22, None, None
    xp = math.floor(pos[0] / binwid)
    yp = math.floor(pos[1] / binwid)
    if abs(xp) > 200 or abs(yp) > 200:
        return 23, None, None
    return 0, xp, yp


def trace_one(job, S, u):
    """One loop iteration of src/main.f90:90-109 (phase 1) or :127-162 (phase 2) for a clear
    bottle.  `u(slot)` supplies the uniforms.  -> (pos, dir, status, xp, yp)"""
    if job.phase == 1:
        pos, d = source_ring(S, u)
    else:
        pos, d = source_point(S.cos_theta_max, S.point_offset, u)
        if job.use_bottle:
            pos, d, st = bottle_forward(S, pos, d, u, bool(job.flags & 1))
            if st:
                return pos, d, st, None, None
    pos, d, st = plano_forward(S.L2, pos, d, u)          # telescope, src/optics_system.f90:28
    if st:
        return pos, d, st, None, None
    pos, d, st = doublet_forward(S.L3, pos, d, bool(job.iris_before), bool(job.iris_after),
                                 job.iris_radius, u)
    if st:
        return pos, d, st, None, None
    k = ((S.img_plane + job.fibre_offset) - pos[2]) / d[2]   # src/optics_system.f90:48-49
    pos = _add(pos, _scale(d, k))
    st, xp, yp = make_image(d, pos, job.image_diameter)
    return pos, d, st, xp, yp
