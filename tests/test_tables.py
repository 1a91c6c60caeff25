"""Structure tables: the generated header (tools/gen_tables.py, index arithmetic) against the oracle's
independently built tables (breadth-first enumeration) and the reference outputs.  CPU only."""
import importlib.util
import os
import re

import numpy as np

from oracle import misti_oracle as mo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = open(os.path.join(ROOT, "misti_b200", "csrc", "misti_tables.h")).read()


def _macro(name):
    m = re.search(r"#define %s (.*)" % name, HEADER)
    body = m.group(1).replace("{", "[").replace("}", "]")
    return eval(body)


def test_header_is_up_to_date():
    spec = importlib.util.spec_from_file_location("gen_tables", os.path.join(ROOT, "tools", "gen_tables.py"))
    gt = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gt)
    ent, diag = gt.generator_entries()
    assert [list(e) for e in ent] == _macro("MISTI_GEN_ENTRIES_INIT")
    assert [list(d) for d in diag] == _macro("MISTI_GEN_DIAG_INIT")
    # the two state enumerations agree
    for i, st in enumerate(gt.STATES):
        assert sorted(st) == sorted(mo.STATES2[i])


def test_generator_tables_match_oracle():
    ent, diag = _macro("MISTI_GEN_ENTRIES_INIT"), _macro("MISTI_GEN_DIAG_INIT")
    rng = np.random.default_rng(3)
    for _ in range(5):
        rate = rng.uniform(0.1, 3.0, 4)
        M = np.zeros((44, 44))
        for r, c, k, n in ent:
            M[r, c] += n * rate[k]
        for c in range(44):
            M[c, c] -= sum(diag[c][k] * rate[k] for k in range(4))
        assert np.max(np.abs(M - mo.generator_two_pop(*rate))) < 1e-13
    ell = _macro("MISTI_ELL_INIT")
    from_ell = sorted([r, c, k, n] for r in range(44) for (c, k, n) in ell[r] if n)
    assert from_ell == sorted(ent)
    assert all(len(row) == 4 for row in ell)


def test_branch_counts_collapse_ancient():
    assert np.array_equal(np.array(_macro("MISTI_W44_INIT")), mo.W44.astype(int))
    assert np.array_equal(np.array(_macro("MISTI_W8_INIT")), mo.W8.astype(int))
    col = _macro("MISTI_COLLAPSE_INIT")
    b = mo.COLLAPSE_BOUNDS
    assert col == [max(k for k in range(8) if b[k] <= i) for i in range(44)]
    assert _macro("MISTI_STATIONARY_INIT") == mo.STATIONARY
    P0 = np.random.default_rng(5).uniform(0, 1, 44)
    out = np.zeros(44)
    out[2] = sum(P0[i] for i in range(44) if _macro("MISTI_ANC2_INIT")[i])
    out[11] = sum(P0[i] for i in range(44) if _macro("MISTI_ANC11_INIT")[i])
    assert np.allclose(out, mo.ancient_sample_reset(P0), atol=1e-15)


def test_pulse_tables():
    for src in (0, 1):
        ent = _macro("MISTI_PULSE%d_INIT" % src)
        rp = _macro("MISTI_PULSE%d_ROWPTR_INIT" % src)
        assert rp[0] == 0 and rp[44] == len(ent)
        for r in (0.0, 0.05, 0.6, 1.0):
            Pm = np.zeros((44, 44))
            for row in range(44):
                for e in ent[rp[row]:rp[row + 1]]:
                    assert e[0] == row
                    Pm[row, e[1]] += e[4] * (1 - r) ** e[2] * r ** e[3]
            assert np.max(np.abs(Pm - mo.pulse_matrix(r, src))) < 1e-14
            assert np.allclose(Pm.sum(axis=0), 1.0)


def test_one_population_spectral_tables():
    L = np.array(_macro("MISTI_L8_INIT"))
    assert np.array_equal(L, mo.L8)
    WG = [np.array(_macro("MISTI_WG%d_INIT" % a)) for a in (6, 3, 1)]
    # W8 exp(x L8) P = sum_k exp(-a_k x) WG_k P
    from scipy.linalg import expm
    rng = np.random.default_rng(11)
    for x in (0.0, 0.3, 2.5):
        P = rng.uniform(0, 1, 8)
        lhs = mo.W8 @ expm(x * L) @ P
        rhs = sum(np.exp(-a * x) * (G @ P) for a, G in zip((6, 3, 1), WG))
        assert np.max(np.abs(lhs - rhs)) < 1e-13


def test_mirror_state_maps():
    """misti_b200.populations.MapIndToState (host bookkeeping, no device call) enumerates the same states."""
    from misti_b200.populations import OnePopulation, TwoPopulations
    tp = TwoPopulations.__new__(TwoPopulations)
    tp.Msize = 44
    for i in range(44):
        st = tp.MapIndToState(i)
        assert sorted((l.d0, l.d1, l.pop) for l in st) == sorted(mo.STATES2[i])
        assert tp.MapStateToInd(st) == i
    op = OnePopulation.__new__(OnePopulation)
    op.Msize = 8
    for i in range(8):
        assert sorted((l.d0, l.d1) for l in op.MapIndToState(i)) == sorted(mo.STATES1[i])
        assert op.MapStateToInd(op.MapIndToState(i)) == i


def test_uniformisation_rate_table():
    """max_c |M_cc| is attained on one of the Pareto-maximal diagonal rows (MISTI_QDIAG)."""
    diag, qd = np.array(_macro("MISTI_GEN_DIAG_INIT")), np.array(_macro("MISTI_QDIAG_INIT"))
    assert len(qd) == _macro("MISTI_QDIAG_N")
    rng = np.random.default_rng(17)
    for _ in range(200):
        rate = rng.uniform(0.0, 1.0, 4) ** rng.integers(1, 6) * 10 ** rng.uniform(-3, 3)
        if rng.random() < 0.3:
            rate[rng.integers(0, 4)] = 0.0
        assert (qd @ rate).max() == (diag @ rate).max()


def test_zero_migration_projector_table():
    """sum_ab coef_ab G0_a G1_b reproduces expm over a run of zero-migration intervals and its time integral."""
    from scipy.linalg import expm
    rp, col, ab, val = (_macro("MISTI_NM_%s_INIT" % n) for n in ("ROWPTR", "COL", "AB", "VAL"))
    assert rp[0] == 0 and rp[44] == _macro("MISTI_NM_NNZ") == len(col) == len(ab) == len(val)
    G = np.zeros((8, 44, 44))
    for r in range(44):
        for e in range(rp[r], rp[r + 1]):
            G[ab[e], r, col[e]] = val[e]
    assert np.allclose(G.sum(axis=0), np.eye(44), atol=1e-15)
    AB = [(0, 0), (1, 0), (3, 0), (6, 0), (0, 1), (0, 3), (0, 6), (1, 1)]
    rng = np.random.default_rng(23)
    for trial in range(4):
        n = 5
        la0, la1, T = rng.uniform(0.2, 4.0, n), rng.uniform(0.2, 4.0, n), rng.uniform(1e-3, 0.6, n)
        P0 = rng.uniform(0, 1, 44)
        P0 /= P0.sum()
        P, I = P0.copy(), np.zeros(44)
        for i in range(n):  # reference: Van Loan augmented exponential per interval
            aug = np.zeros((45, 45))
            aug[:44, :44] = mo.generator_two_pop(la0[i], la1[i], 0.0, 0.0) * T[i]
            aug[:44, 44] = P * T[i]
            E = expm(aug)
            I += E[:44, 44]
            P = E[:44, :44] @ P
        X0, X1, c = 0.0, 0.0, np.zeros(8)
        for i in range(n):
            for k, (a, b) in enumerate(AB):
                z = a * la0[i] + b * la1[i]
                c[k] += np.exp(-a * X0 - b * X1) * (T[i] if z == 0 else -np.expm1(-z * T[i]) / z)
            X0 += la0[i] * T[i]
            X1 += la1[i] * T[i]
        e = [np.exp(-a * X0 - b * X1) for a, b in AB]
        assert np.max(np.abs(sum(e[k] * (G[k] @ P0) for k in range(8)) - P)) < 1e-14
        assert np.max(np.abs(sum(c[k] * (G[k] @ P0) for k in range(8)) - I)) < 1e-14


def test_sixteen_lane_layout():
    """every state is owned by exactly one (lane, slot); local + remote entries of a row are exactly its generator
    entries; a slot's 16 words sit in 16 different bank pairs."""
    row, rem, loc, pos = (_macro("MISTI_L16_%s_INIT" % n) for n in ("ROW", "REM", "LOC", "POS"))
    ell = _macro("MISTI_ELL_INIT")
    assert sorted(r for ln in row for r in ln) == list(range(48))
    assert sorted(pos) == list(range(48))
    rw = (3, 2, 3)
    for ln in range(16):
        for sl in range(3):
            r = row[ln][sl]
            assert pos[r] == 16 * sl + ln
            got = []
            for e in range(rw[sl]):
                c, code = rem[ln][sl][e]
                if code != 12:
                    got.append([c, code & 3, 1 << (code >> 2)])
            assert all(code == 12 for (_c, code) in rem[ln][sl][rw[sl]:])
            for j in range(2):
                if loc[ln][sl][j] != 12:
                    got.append([row[ln][(sl + 1 + j) % 3], loc[ln][sl][j] & 3, 1 << (loc[ln][sl][j] >> 2)])
            want = [list(e) for e in ell[r] if e[2]] if r < 44 else []
            assert sorted(got) == sorted(want), (ln, sl)
    codes = {c for ln in rem for sl in ln for (_x, c) in sl} | {c for ln in loc for sl in ln for c in sl}
    assert codes <= set(range(10)) | {12}  # the record's table slots 10, 11 are free for 1/q and lam


def test_sixteen_lane_run_table():
    """the balanced run table of the 16-lane layout holds exactly the projector entries, row by row"""
    rp, col, ab, val = (_macro("MISTI_NM_%s_INIT" % n) for n in ("ROWPTR", "COL", "AB", "VAL"))
    pos = _macro("MISTI_L16_POS_INIT")
    n = _macro("MISTI_R16_LEN")
    rv = _macro("MISTI_R16_VAL_INIT")
    rm = [int(x) for x in re.search(r"#define MISTI_R16_META_INIT (.*)", HEADER).group(1).replace("u", "").strip("{} ").split(",")]
    assert len(rv) == len(rm) == 16 * n
    got = []
    for ln in range(16):
        open_row = None
        for k in range(n):
            m, v = rm[16 * k + ln], rv[16 * k + ln]
            if v == 0.0 and m == 0:
                assert open_row is None  # padding only after a completed row
                continue
            ypos, c, e, last, rpos = m & 255, (m >> 8) & 255, (m >> 16) & 255, (m >> 24) & 1, m >> 25
            assert e == (16 if c == 0 else 7 + c)
            assert open_row in (None, rpos)
            open_row = None if last else rpos
            got.append((pos.index(rpos), pos.index(ypos), c, v))
        assert open_row is None
    want = [(r, col[i], ab[i], val[i]) for r in range(44) for i in range(rp[r], rp[r + 1])]
    assert sorted(got) == sorted(want)


def test_pair_kernel_code_is_up_to_date_and_equals_the_generator(tmp_path, monkeypatch):
    """csrc/misti_pair_code.h (the straight-line code of the pair-of-lanes JSFS kernel) is what tools/gen_pair_tables.py
    generates today -- the generator asserts the deme-exchange symmetry of the generator entries, of the projector products
    of the zero-migration runs, of StateToJAF and of CollapsePops on the way -- and one term of its sweep, interpreted here
    for both lanes of a pair, is the mat-vec with A = I + M / q of the oracle's generator."""
    spec = importlib.util.spec_from_file_location("gen_pair_tables", os.path.join(ROOT, "tools", "gen_pair_tables.py"))
    gp = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gp)
    out = tmp_path / "misti_pair_code.h"
    monkeypatch.setattr(gp, "OUT", str(out))
    gp.main()
    code = out.read_text()
    assert code == open(os.path.join(ROOT, "misti_b200", "csrc", "misti_pair_code.h")).read()
    # interpret MISTI_PAIR_TERM for the two lanes of a pair
    rows = eval(re.search(r"#define MISTI_PAIR_ROW_INIT (.*)", code).group(1).replace("{", "[").replace("}", "]"))
    dcls = eval(re.search(r"#define MISTI_PAIR_DIAG_INIT (.*)", code).group(1).replace("{", "[").replace("}", "]"))
    body = code[code.index("#define MISTI_PAIR_TERM"):]
    body = body[body.index("\n") + 1:body.index("} while (0)")]
    stmts = [ln.strip().rstrip("\\").strip() for ln in body.splitlines() if ln.strip()]
    rng = np.random.default_rng(8)
    rate = rng.uniform(0.1, 2.0, 4)  # coalescence in deme 0 / 1, migration out of deme 0 / 1
    ent, diag = _macro("MISTI_GEN_ENTRIES_INIT"), _macro("MISTI_GEN_DIAG_INIT")
    q = max(sum(d[k] * rate[k] for k in range(4)) for d in diag)
    A = np.zeros((44, 44))
    for r, c, k, n in ent:
        A[r, c] += n * rate[k] / q
    for c in range(44):
        A[c, c] = 1.0 - sum(diag[c][k] * rate[k] for k in range(4)) / q
    y44 = rng.uniform(0.0, 1.0, 44)
    want = A @ y44
    lanes = []
    for role in (0, 1):
        rq = [rate[k ^ role] / q for k in range(4)]  # lane 1 reads the rate table with the kinds swapped
        cf = [(1, 2, 4)[c // 4] * rq[c % 4] for c in range(10)]
        dg = [max(0.0, 1.0 - sum(t[k] * rq[k] for k in range(4))) for t in dcls]
        lanes.append({"y": [y44[s] for s in rows[role]], "cf": cf, "dg": dg})
    new = []
    for role in (0, 1):
        me, other = lanes[role], lanes[1 - role]
        env = {"y": list(me["y"]), "Ia": [0.0] * 23, "P1": [0.0] * 23, "cf": me["cf"], "dg": me["dg"], "tq": 0.0, "p": 1.0,
               "fma": lambda a, b, c: a * b + c}
        for st in stmts:
            st = st.replace("const double ", "").rstrip(";")
            for part in [x for x in st.split(";") if x.strip()]:
                part = part.strip()
                m = re.match(r"(z\d+) = __shfl_xor_sync\(msk, y\[(\d+)\], 1\)", part)
                if m:
                    env[m.group(1)] = other["y"][int(m.group(2))]  # the partner's OLD value
                    continue
                exec(part, {}, env)
        new.append(env["y"])
    for role in (0, 1):
        for i, s in enumerate(rows[role]):
            assert abs(new[role][i] - want[s]) < 1e-14, (role, i, s)


def test_pair_kernel_run_code_equals_the_projector_table():
    """MISTI_PAIR_RUN (a run of intervals without migration in the pair kernel), interpreted for both lanes: P <- sum_ab e_ab
    G_ab P and integral += sum_ab c_ab G_ab P with the projector products of misti_tables.h; lane 1 reads the coefficients
    with a and b exchanged."""
    code = open(os.path.join(ROOT, "misti_b200", "csrc", "misti_pair_code.h")).read()
    rows = eval(re.search(r"#define MISTI_PAIR_ROW_INIT (.*)", code).group(1).replace("{", "[").replace("}", "]"))
    swap = eval(re.search(r"#define MISTI_PAIR_ABSWAP_INIT (.*)", code).group(1).replace("{", "[").replace("}", "]"))
    body = code[code.index("#define MISTI_PAIR_RUN"):]
    body = body[body.index("\n") + 1:body.index("} while (0)")]
    stmts = [ln.strip().rstrip("\\").strip() for ln in body.splitlines() if ln.strip()]
    rp, col, ab, val = (_macro("MISTI_NM_ROWPTR_INIT"), _macro("MISTI_NM_COL_INIT"), _macro("MISTI_NM_AB_INIT"), _macro("MISTI_NM_VAL_INIT"))
    rng = np.random.default_rng(12)
    e = np.concatenate([[1.0], rng.uniform(0.1, 1.0, 7)])
    c = rng.uniform(0.1, 2.0, 8)
    y44 = rng.uniform(0.0, 1.0, 44)
    want_p, want_i = np.zeros(44), np.zeros(44)
    for r in range(44):
        for k in range(rp[r], rp[r + 1]):
            want_p[r] += e[ab[k]] * val[k] * y44[col[k]]
            want_i[r] += c[ab[k]] * val[k] * y44[col[k]]
    lanes = [[y44[s] for s in rows[role]] for role in (0, 1)]
    for role in (0, 1):
        env = {"y": list(lanes[role]), "Ia": [0.0] * 23, "fma": lambda a, b, cc: a * b + cc,
               "E": [e[swap[a] if role else a] for a in range(8)], "C": [c[swap[a] if role else a] for a in range(8)]}
        for st in stmts:
            st = st.replace("const double ", "").replace("double ", "").replace("{", "").replace("}", "")
            for part in [x.strip() for x in st.split(";") if x.strip()]:
                m = re.match(r"(z\d+) = __shfl_xor_sync\(msk, y\[(\d+)\], 1\)", part)
                if m:
                    env[m.group(1)] = lanes[1 - role][int(m.group(2))]
                    continue
                for piece in part.split(", ") if re.match(r"pe\d+ = 0.0, ir\d+ = 0.0", part) else [part]:
                    exec(piece, {}, env)
        for i, s in enumerate(rows[role]):
            assert abs(env["y"][i] - want_p[s]) < 1e-13 and abs(env["Ia"][i] - want_i[s]) < 1e-13, (role, i, s)
