#!/usr/bin/env python3
"""Generate misti_b200/csrc/misti_pair_code.h: the straight-line code of the TWO-lanes-per-item JSFS kernel.

The lineage chain of two demes has a symmetry the 16-lane kernel does not use: exchanging the two demes (sigma) maps the
44 states onto themselves, generator entries onto generator entries with the rate kinds swapped (coalescence in deme 0 <->
deme 1, migration out of deme 0 <-> out of deme 1), and leaves StateToJAF and CollapsePops alone.  sigma has 21 two-cycles
and 2 fixed states.  So an item is run by a PAIR of lanes: lane 0 owns one state of every two-cycle plus the fixed states,
lane 1 the images -- and both lanes execute THE SAME instruction stream (row i of lane 1 is sigma(row i of lane 0); its
entries have the same local column indices, own / partner relation included), only with the rate table read with swapped
kinds.  Every generator entry is therefore a compile-time fact: the mat-vec of the uniformisation sweep is 23 rows of
fused multiply-adds on registers, the partner's values come by shuffle, no shared memory, no index arithmetic.

Which state of a two-cycle goes to lane 0 is chosen (seeded local search) so that few of the partner's values are needed.
Rows are ordered by lineage count (2, 3, 4): coalescence only lowers the count, so a level can be overwritten as soon as
the level below has been computed, which keeps the temporaries of a mat-vec at one level's rows.

Input: the tables of misti_tables.h (tools/gen_tables.py) -- parsed from the header so that the two stay consistent.
"""
import os
import random
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gen_tables as gt  # noqa: E402

HDR = os.path.join(HERE, "..", "misti_b200", "csrc", "misti_tables.h")
OUT = os.path.join(HERE, "..", "misti_b200", "csrc", "misti_pair_code.h")


def table(name):
    txt = open(HDR).read()
    m = re.search(r"#define %s (\{.*\})\s*$" % name, txt, re.M)
    assert m, name
    return eval(m.group(1).replace("{", "[").replace("}", "]").replace("u", ""))


def main():
    ent, diag = gt.generator_entries()
    STATES, IDX = gt.STATES, gt.IDX
    sigma = [IDX[gt.key([(d0, d1, 1 - p) for (d0, d1, p) in st])] for st in STATES]
    assert all(sigma[sigma[i]] == i for i in range(44))
    entset = {(r, c): (k, n) for (r, c, k, n) in ent}
    for (r, c), (k, n) in entset.items():
        assert entset[(sigma[r], sigma[c])] == (k ^ 1, n)
    for c in range(44):
        assert [diag[sigma[c]][k ^ 1] for k in range(4)] == diag[c]
    level = [len(st) for st in STATES]
    fixed = [i for i in range(44) if sigma[i] == i]
    cycles = sorted((i, sigma[i]) for i in range(44) if i < sigma[i])
    assert len(fixed) == 2 and len(cycles) == 21
    rows_of = [[(c, k, n) for (r, c, k, n) in ent if r == rr] for rr in range(44)]

    def need(orient):
        """partner values needed by lane 0's mat-vec: orient[j] = 0/1 picks which state of cycle j lane 0 owns"""
        mine = set(cy[o] for cy, o in zip(cycles, orient)) | set(fixed)
        nd = set()
        for r in mine:
            for (c, _k, _n) in rows_of[r]:
                if c not in mine:
                    nd.add(sigma[c])  # the partner's local row of that column = lane 0's row sigma(c)
        return nd

    rnd = random.Random(20261019)
    best_o, best_c = None, 99
    for _restart in range(40):
        o = [rnd.randrange(2) for _ in cycles]
        c = len(need(o))
        for _ in range(4000):
            j = rnd.randrange(len(cycles))
            o[j] ^= 1
            c2 = len(need(o))
            if c2 <= c:
                c = c2
            else:
                o[j] ^= 1
        if c < best_c:
            best_c, best_o = c, list(o)
    orient = best_o
    # local rows: by level (2, 3, 4), cycles then fixed states
    mine = [cy[o] for cy, o in zip(cycles, orient)] + fixed
    mine.sort(key=lambda s: (level[s], s in fixed, s))
    N = len(mine)
    assert N == 23
    row0 = mine
    row1 = [sigma[s] for s in mine]
    loc0 = {s: i for i, s in enumerate(row0)}  # lane 0's local index of its own states
    loc1 = {s: i for i, s in enumerate(row1)}  # local index (in the partner) of the partner's states
    isfixed = [int(s in fixed) for s in row0]

    def ref(c):
        """how lane 0 reads state c: ('y', i) own or ('z', i) partner's"""
        if c in loc0:
            return ("y", loc0[c])
        return ("z", loc1[c])

    def code_of(kind, cnt):
        return kind + {1: 0, 2: 4, 4: 8}[cnt]
    # diagonal classes
    dcls = sorted(set(tuple(diag[s]) for s in row0))
    dci = [dcls.index(tuple(diag[s])) for s in row0]
    # ---- the sweep term, level by level
    lev_rows = {L: [i for i in range(N) if level[row0[i]] == L] for L in (2, 3, 4)}
    shuffled = set()
    term = []
    needed = set()
    for L in (2, 3, 4):
        want = []
        for i in lev_rows[L]:
            for (c, _k, _n) in rows_of[row0[i]]:
                kind, j = ref(c)
                if kind == "z" and j not in shuffled:
                    shuffled.add(j)
                    want.append(j)
                    needed.add(j)
        for j in sorted(want):
            term.append("const double z%d = __shfl_xor_sync(msk, y[%d], 1);" % (j, j))
        for i in lev_rows[L]:
            term.append("Ia[%d] = fma(tq, y[%d], Ia[%d]);" % (i, i, i))
            expr = "dg[%d] * y[%d]" % (dci[i], i)
            for (c, k, n) in rows_of[row0[i]]:
                kind, j = ref(c)
                src = "y[%d]" % j if kind == "y" else "z%d" % j
                expr = "fma(cf[%d], %s, %s)" % (code_of(k, n), src, expr)
            term.append("const double n%d = %s;" % (i, expr))
        for i in lev_rows[L]:
            term.append("y[%d] = n%d; P1[%d] = fma(p, n%d, P1[%d]);" % (i, i, i, i, i))
    nfma = sum(len(rows_of[s]) for s in row0)
    # ---- zero-migration runs: P <- sum_ab E[ab] G_ab P, Ia += sum_ab C[ab] G_ab P
    rp, col, ab, val = table("MISTI_NM_ROWPTR_INIT"), table("MISTI_NM_COL_INIT"), table("MISTI_NM_AB_INIT"), table("MISTI_NM_VAL_INIT")
    nm = {}
    for r in range(44):
        for e in range(rp[r], rp[r + 1]):
            nm[(r, col[e], ab[e])] = val[e]
    SWAP = [0, 4, 5, 6, 1, 2, 3, 7]  # ab index of (b, a): see NM_AB in gen_tables.py
    for (r, c, a), v in nm.items():
        assert nm[(sigma[r], sigma[c], SWAP[a])] == v
    run = []
    zrun = sorted(set(loc1[c] for (r, c, _a) in nm if r in loc0 and c not in loc0))
    for j in zrun:
        run.append("const double z%d = __shfl_xor_sync(msk, y[%d], 1);" % (j, j))
    nrun = 0
    for i, s in enumerate(row0):
        es = sorted((c, a, v) for (r, c, a), v in nm.items() if r == s)
        run.append("double pe%d = 0.0, ir%d = 0.0;" % (i, i))
        for (c, a, v) in es:
            kind, j = ref(c)
            src = "y[%d]" % j if kind == "y" else "z%d" % j
            run.append("{ const double t = %s * %s; pe%d = fma(E[%d], t, pe%d); ir%d = fma(C[%d], t, ir%d); }" % (repr(float(v)), src, i, a, i, i, a, i))
            nrun += 1
    for i in range(N):
        run.append("y[%d] = pe%d; Ia[%d] += ir%d;" % (i, i, i, i))
    # ---- tail: jl[c] += W[c][row] X[row] (X = the occupancy integrals) and V[c][block(row)] P[row]
    W44, collapse = table("MISTI_W44_INIT"), table("MISTI_COLLAPSE_INIT")
    for s in range(44):
        assert collapse[sigma[s]] == collapse[s] and all(W44[c][sigma[s]] == W44[c][s] for c in range(7))
    tail_w = []
    for c in range(7):
        terms = ["%d.0 * X[%d]" % (W44[c][s], i) for i, s in enumerate(row0) if W44[c][s]]
        # pairwise-free plain sum, fixed order
        expr = "0.0"
        for i, s in enumerate(row0):
            if W44[c][s]:
                expr = "fma(%d.0, X[%d], %s)" % (W44[c][s], i, expr)
        tail_w.append("jw[%d] = %s;" % (c, expr))
    tail_v = []
    for c in range(7):
        expr = "jv[%d]" % c
        for i, s in enumerate(row0):
            expr = "fma(V[%d][%d], y[%d], %s)" % (c, collapse[s], i, expr)
        tail_v.append("jv[%d] = %s;" % (c, expr))
    o = []
    o.append("// GENERATED by tools/gen_pair_tables.py -- do not edit.  See that file for the derivation.")
    o.append("#pragma once")
    o.append("#define MISTI_PAIR_N %d" % N)
    o.append("// global state index of local row i: lane 0, lane 1 (= the deme-swapped image)")
    o.append("#define MISTI_PAIR_ROW_INIT { {%s}, {%s} }" % (",".join(map(str, row0)), ",".join(map(str, row1))))
    o.append("// rows that are their own image (held by both lanes; lane 0's copy counts)")
    o.append("#define MISTI_PAIR_FIXED_INIT { %s }" % ",".join(map(str, isfixed)))
    o.append("#define MISTI_PAIR_NDIAG %d" % len(dcls))
    o.append("// diagonal multiplicities {coal 0, coal 1, mig 0, mig 1} of the diagonal classes (lane 0's kinds)")
    o.append("#define MISTI_PAIR_DIAG_INIT { %s }" % ", ".join("{%d,%d,%d,%d}" % t for t in dcls))
    o.append("// ab index of the projector product with the two demes exchanged")
    o.append("#define MISTI_PAIR_ABSWAP_INIT { %s }" % ",".join(map(str, SWAP)))
    o.append("// one term of a uniformisation sweep: Ia += tq y;  y <- A y;  P1 += p y   (%d rows, %d off-diagonal entries, %d values"
             % (N, nfma, len(needed)))
    o.append("// of the partner by shuffle).  y, Ia, P1: double[%d]; cf: coefficient by code; dg: diagonal by class" % N)
    o.append("#define MISTI_PAIR_TERM(y, Ia, P1, cf, dg, tq, p, msk) do { \\")
    for ln in term:
        o.append("    " + ln + " \\")
    o.append("} while (0)")
    o.append("// a run of intervals without migration (%d projector entries, %d values of the partner)" % (nrun, len(zrun)))
    o.append("#define MISTI_PAIR_RUN(y, Ia, C, E, msk) do { \\")
    for ln in run:
        o.append("    " + ln + " \\")
    o.append("} while (0)")
    o.append("// StateToJAF . X")
    o.append("#define MISTI_PAIR_TAIL_W(X, jw) do { \\")
    for ln in tail_w:
        o.append("    " + ln + " \\")
    o.append("} while (0)")
    o.append("// + V[c][CollapsePops(row)] y[row]")
    o.append("#define MISTI_PAIR_TAIL_V(V, y, jv) do { \\")
    for ln in tail_v:
        o.append("    " + ln + " \\")
    o.append("} while (0)")
    # diagonal of the uniformised generator by class: 1 - sum_k multiplicity_k rate_k / q, floored at 0
    o.append("// dg[class] from rq[k] = rate_k / q (the lane's own kinds)")
    o.append("#define MISTI_PAIR_DIAG(dg, rq) do { \\")
    for j, t in enumerate(dcls):
        a = " + ".join("%d.0 * rq[%d]" % (t[k], k) for k in (0, 1) if t[k]) or "0.0"
        b = " + ".join("%d.0 * rq[%d]" % (t[k], k) for k in (2, 3) if t[k]) or "0.0"
        o.append("    { const double d = (%s) + (%s); dg[%d] = d < 1.0 ? 1.0 - d : 0.0; } \\" % (a, b, j))
    o.append("} while (0)")
    wg = [table("MISTI_WG6_INIT"), table("MISTI_WG3_INIT"), table("MISTI_WG1_INIT")]
    o.append("// V[c][b] = (c6 WG6 + c3 WG3) + c1 WG1: the one-population tail folded into weights of the collapsed state")
    o.append("#define MISTI_PAIR_V(V, c6, c3, c1) do { \\")
    for c in range(7):
        for b in range(8):
            o.append("    V[%d][%d] = (c6 * %s + c3 * %s) + c1 * %s; \\" % (c, b, repr(float(wg[0][c][b])), repr(float(wg[1][c][b])), repr(float(wg[2][c][b]))))
    o.append("} while (0)")
    o.append("// the rows held by both lanes count once: lane 1 clears its copy before the tail")
    o.append("#define MISTI_PAIR_ZERO_FIXED(a, keep) do { %s } while (0)" % " ".join("a[%d] = keep ? a[%d] : 0.0;" % (i, i) for i in range(N) if isfixed[i]))
    # the sampling configuration: state 2 (both genome-1 lineages in deme 0, genome-2 in deme 1)
    ir, ii = (0, row0.index(2)) if 2 in row0 else (1, row1.index(2))
    o.append("// where the chain starts (state 2): local row and owning lane")
    o.append("#define MISTI_PAIR_START_ROW %d" % ii)
    o.append("#define MISTI_PAIR_START_ROLE %d" % ir)
    with open(OUT, "w") as f:
        f.write("\n".join(o) + "\n")
    print("wrote", OUT, "rows", N, "fixed", fixed, "off-diagonal entries per lane", nfma, "partner values per term", len(needed),
          "run entries per lane", nrun, "partner values per run", len(zrun), "diag classes", len(dcls))


if __name__ == "__main__":
    main()
