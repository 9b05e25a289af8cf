"""Builds the constraint polynomial of the declared synthetic circuit as DATA: a tap table and a PolyStep list in
the shape of upstream's `PolyExtStepDef` (TEST INFRASTRUCTURE: lives with the oracle; the product only ever sees
the resulting arrays through `hfb200_init_ir`).

Ops: 0 CONST(a = canonical value) | 1 GET(a = tap index) | 2 GET_GLOBAL(a = 0 globals / 1 mix, b = offset) |
     3 ADD(a, b) | 4 SUB(a, b) | 5 MUL(a, b) | 6 TRUE | 7 AND_EQZ(a = mix var, b = fp var) | 8 AND_COND(a, b = cond, c = inner)
fp vars and mix vars are two separate SSA index spaces; every step pushes one value."""
import numpy as np

P = 2013265921
ACCUM, CODE, DATA = 0, 1, 2
CODE_FIXED = 4


def build(widths, variant=0, nest=False):
    wc, wd, wa = widths
    F, npv, nch = wd // 2, wd // 4, wa // 4
    taps = []
    for c in range(wa):
        taps += [(ACCUM, c, 0), (ACCUM, c, 1)]
    for c in range(wc):
        taps.append((CODE, c, 0))
    for c in range(wd):
        taps.append((DATA, c, 0))
        if c < npv:
            taps.append((DATA, c, 1))
            if variant:
                taps.append((DATA, c, 2))
    tap_index = {t: i for i, t in enumerate(taps)}
    steps = []
    nfp = [0]
    nmix = [0]
    cache = {}

    def fp(op, a=0, b=0, c=0):
        steps.append((op, a, b, c))
        nfp[0] += 1
        return nfp[0] - 1

    def mixv(op, a=0, b=0, c=0):
        steps.append((op, a, b, c))
        nmix[0] += 1
        return nmix[0] - 1

    def get(g, off, back=0):
        k = ("t", g, off, back)
        if k not in cache:
            cache[k] = fp(1, tap_index[(g, off, back)])
        return cache[k]

    def const(v):
        k = ("c", v % P)
        if k not in cache:
            cache[k] = fp(0, v % P)
        return cache[k]

    def glob(arr, i):
        k = ("g", arr, i)
        if k not in cache:
            cache[k] = fp(2, arr, i)
        return cache[k]

    add = lambda a, b: fp(3, a, b)
    sub = lambda a, b: fp(4, a, b)
    mul = lambda a, b: fp(5, a, b)

    active, first = get(CODE, 0), get(CODE, 1)
    m = mixv(6)
    for k in range(F):
        A, B, Cc, D = get(DATA, k % F), get(DATA, (5 * k + 1) % F), get(DATA, (11 * k + 2) % F), get(DATA, (17 * k + 3) % F)
        pcol = (7 * k + 1) % npv
        X = get(CODE, CODE_FIXED + k % (wc - CODE_FIXED))
        f = k & 3
        if f == 0:
            e = add(mul(A, B), Cc)
        elif f == 1:
            e = add(mul(mul(A, B), Cc), get(DATA, pcol, 2 if variant else 1))
        elif f == 2:
            e = mul(mul(mul(add(A, X), B), Cc), D)
        else:
            e = add(add(mul(get(DATA, pcol, 1), B), mul(Cc, D)), X)
        diff = sub(get(DATA, F + k), e)
        if nest and f == 0:
            # the same constraint written with AND_COND: cond = active, inner = (True AND_EQZ diff)
            inner = mixv(7, mixv(6), diff)
            m = mixv(8, m, active, inner)
        else:
            m = mixv(7, m, mul(active, diff))
    one, nbeta = const(1), const(P - 11)
    nf = sub(one, first)
    for r in range(nch):
        acc = [get(ACCUM, 4 * r + k) for k in range(4)]
        s = [mul(nf, get(ACCUM, 4 * r + k, 1)) for k in range(4)]
        t = [glob(1, 4 * r + k) for k in range(4)]
        s[0] = add(s[0], first)
        t[0] = add(t[0], get(DATA, (13 * r + 5) % wd))
        # Fp4 product s * t with x^4 = -11, same association as oracle/fp.h and csrc/field.cuh
        pr0 = add(mul(s[0], t[0]), mul(nbeta, add(add(mul(s[1], t[3]), mul(s[2], t[2])), mul(s[3], t[1]))))
        pr1 = add(add(mul(s[0], t[1]), mul(s[1], t[0])), mul(nbeta, add(mul(s[2], t[3]), mul(s[3], t[2]))))
        pr2 = add(add(add(mul(s[0], t[2]), mul(s[1], t[1])), mul(s[2], t[0])), mul(nbeta, mul(s[3], t[3])))
        pr3 = add(add(mul(s[0], t[3]), mul(s[1], t[2])), add(mul(s[2], t[1]), mul(s[3], t[0])))
        for k, pr in enumerate((pr0, pr1, pr2, pr3)):
            m = mixv(7, m, mul(active, sub(acc[k], pr)))
    m = mixv(7, m, mul(first, sub(get(DATA, 0), glob(0, 0))))
    return {"taps": np.array(taps, np.uint32), "steps": np.array(steps, np.uint32), "ret": m, "n_globals": 32, "n_mix": 4 * nch,
            "widths": tuple(widths), "info": b"SYNTH_RV32IM:v1_"}  # the stand-in circuit itself, as data: same CIRCUIT_INFO as the built-in form


def build_scaled(widths=(24, 320, 56), n_groups=420, per_group=16, seed=7):
    """A synthetic circuit at the SCALE of rv32im-v2 (thousands of constraints, W ~ 400, taps at several
    `back` distances): ~n_groups * per_group constraints of degree <= 5 over random taps, grouped under AND_COND selectors the way
    zirgen's major/minor muxes are, written in the same PolyExtStep shape.  It exercises the data-defined path (hfb200_init_ir:
    bytecode compiler, NVRTC specialisation, generic DEEP kernels) at the limits the library declares: 4 distinct back values,
    8 distinct tap sets, up to 4 taps per register.  NOT satisfiable by construction -- the prover does not need a satisfying
    witness to be timed or compared with the oracle (the seal then fails verification at the constraint check, as it must)."""
    import random
    rnd = random.Random(seed)
    wc, wd, wa = widths
    tap_sets = [(0,), (0, 1), (0, 1, 2), (0, 1, 2, 3), (0, 2), (0, 3), (0, 1, 3), (0, 2, 3)]   # 8 sets, backs {0, 1, 2, 3}
    taps = []
    for c in range(wa):
        taps += [(ACCUM, c, 0), (ACCUM, c, 1)]
    for c in range(wc):
        taps.append((CODE, c, 0))
    data_sets = {}
    for c in range(wd):
        ts = tap_sets[c % len(tap_sets)] if c % 3 else (0, 1, 2, 3)
        data_sets[c] = ts
        for b in ts:
            taps.append((DATA, c, b))
    tap_index = {t: i for i, t in enumerate(taps)}
    steps, nfp, nmix, cache = [], [0], [0], {}

    def fp(op, a=0, b=0, c=0):
        steps.append((op, a, b, c)); nfp[0] += 1
        return nfp[0] - 1

    def mixv(op, a=0, b=0, c=0):
        steps.append((op, a, b, c)); nmix[0] += 1
        return nmix[0] - 1

    def get(g, off, back=0):
        k = (g, off, back)
        if k not in cache:
            cache[k] = fp(1, tap_index[k])
        return cache[k]

    def rnd_data():
        c = rnd.randrange(wd)
        return get(DATA, c, rnd.choice(data_sets[c]))

    add = lambda a, b: fp(3, a, b)
    sub = lambda a, b: fp(4, a, b)
    mul = lambda a, b: fp(5, a, b)
    m = mixv(6)
    n_cons = 0
    for g in range(n_groups):
        sel = get(CODE, g % wc)
        inner = mixv(6)
        for k in range(per_group):
            A, B, Cc, D, E = rnd_data(), rnd_data(), rnd_data(), rnd_data(), rnd_data()
            f = (g + k) & 3
            if f == 0:
                e = sub(add(mul(A, B), Cc), D)
            elif f == 1:
                e = sub(mul(mul(A, B), Cc), mul(D, E))
            elif f == 2:
                e = mul(mul(mul(add(A, fp(2, 1, (g + k) % wa)), B), Cc), sub(D, E))
            else:
                e = add(add(mul(A, get(ACCUM, (g + k) % wa, 1)), mul(Cc, D)), fp(2, 0, (g + k) % 32))
            inner = mixv(7, inner, e)
            n_cons += 1
        m = mixv(8, m, sel, inner)
    return {"taps": np.array(taps, np.uint32), "steps": np.array(steps, np.uint32), "ret": m, "n_globals": 32, "n_mix": wa,
            "widths": tuple(widths), "n_constraints": n_cons, "info": None}  # None: the default of a data-defined circuit (upstream's rv32im-v2 string)
