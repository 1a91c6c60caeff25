#!/usr/bin/env python3
"""Generate misti_b200/csrc/misti_tables.h: the constant structure tables of the lineage chains.

Derivation is index-arithmetic on the state layout documented in SURVEY.md section 8(a) (which
restates TwoPopulations.MapStateToInd/MapIndToState, TwoPopulations.py:99-186, and
OnePopulation.py:64-107) -- deliberately a different construction from oracle/misti_oracle.py
(which closes the state space by breadth-first search), so that tests/test_tables.py can
cross-check the two and both against tests/golden/tables.json (reference output).

A state is a multiset of lineages; a lineage is (d0, d1, deme) = (#genome-1 samples below it,
#genome-2 samples below it, deme it sits in).

Emitted tables
  MISTI_GEN_*     off-diagonal generator entries (row, col, kind, count); kind 0/1 = coalescence
                  in deme 0/1 (rate la[kind]), 2/3 = migration out of deme 0/1 (rate mu[kind-2]);
                  MISTI_GEN_DIAG[col][kind] = multiplicity on the diagonal (entered with minus).
  MISTI_ELL       the same off-diagonal entries grouped by row, 4 slots {col, kind, count} per row.
  MISTI_W44/W8    branch-type counts per state (StateToJAF).
  MISTI_PULSE_*   pulse map entries (row, col, a, b): value (1-r)^a r^b, per source deme.
  MISTI_ANC_*     ancient-sample reset masks (AncientSampleP0).
  MISTI_COLLAPSE  44 -> 8 block index (CollapsePops).
  MISTI_WG*       W8 * G_k for the three spectral projectors of the one-population generator
                  L8 (eigenvalues -6, -3, -1), and L8 itself.
"""
import itertools
import os
from fractions import Fraction

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "misti_b200", "csrc", "misti_tables.h")

# ---- lineage "types" by descendant counts
TYPES_BY_COUNT = {
    4: [[(1, 0), (1, 0), (0, 1), (0, 1)]],
    3: [[(2, 0), (0, 1), (0, 1)], [(1, 1), (1, 0), (0, 1)], [(0, 2), (1, 0), (1, 0)]],
    2: [[(2, 1), (0, 1)], [(1, 2), (1, 0)], [(2, 0), (0, 2)], [(1, 1), (1, 1)]],
}
CONFIGS = TYPES_BY_COUNT[4] + TYPES_BY_COUNT[3] + TYPES_BY_COUNT[2]  # == one-population state order
BASE = [0, 9, 15, 23, 29, 33, 37, 41, 44]


def key(state):
    return tuple(sorted(state))


def build_states():
    """index -> canonical multiset of (d0, d1, deme), following the layout formulas."""
    idx = {}
    for cfg_i, cfg in enumerate(CONFIGS):
        for demes in itertools.product((0, 1), repeat=len(cfg)):
            st = [(d0, d1, p) for (d0, d1), p in zip(cfg, demes)]
            if cfg_i == 0:
                j = demes[0] + demes[1]
                i = demes[2] + demes[3]
                ind = i + 3 * j
            elif cfg_i in (1, 3):
                ind = BASE[cfg_i] + 3 * demes[0] + demes[1] + demes[2]
            elif cfg_i == 2:
                ind = BASE[cfg_i] + 4 * demes[0] + 2 * demes[1] + demes[2]
            elif cfg_i in (4, 5, 6):
                ind = BASE[cfg_i] + 2 * demes[0] + demes[1]
            else:
                ind = BASE[cfg_i] + demes[0] + demes[1]
            k = key(st)
            if k in idx:
                assert idx[k] == ind
            idx[k] = ind
    assert sorted(idx.values()) == list(range(44))
    inv = {v: list(k) for k, v in idx.items()}
    return idx, [inv[i] for i in range(44)]


IDX, STATES = build_states()
SLOT = {(1, 0): 0, (2, 0): 1, (0, 1): 2, (1, 1): 3, (2, 1): 4, (0, 2): 5, (1, 2): 6}


def generator_entries():
    off = {}
    diag = [[0, 0, 0, 0] for _ in range(44)]
    for col, st in enumerate(STATES):
        n = len(st)
        for i in range(n):
            d0, d1, p = st[i]
            moved = list(st)
            moved[i] = (d0, d1, 1 - p)
            row = IDX[key(moved)]
            off[(row, col, 2 + p)] = off.get((row, col, 2 + p), 0) + 1
            diag[col][2 + p] += 1
            for j in range(i + 1, n):
                if st[j][2] != p:
                    continue
                rest = [st[k] for k in range(n) if k not in (i, j)]
                rest.append((d0 + st[j][0], d1 + st[j][1], p))
                diag[col][p] += 1
                if len(rest) >= 2:
                    row = IDX[key(rest)]
                    off[(row, col, p)] = off.get((row, col, p), 0) + 1
    pos = {}
    for (r, c, k) in off:
        assert (r, c) not in pos, "two rate kinds on one off-diagonal entry"
        pos[(r, c)] = k
    ent = sorted((r, c, k, cnt) for (r, c, k), cnt in off.items())
    return ent, diag


def pulse_entries(src):
    ent = {}
    for col, st in enumerate(STATES):
        movers = [k for k, l in enumerate(st) if l[2] == src]
        for mask in range(1 << len(movers)):
            new = list(st)
            a = b = 0
            for bit, k in enumerate(movers):
                if mask >> bit & 1:
                    new[k] = (st[k][0], st[k][1], 1 - src)
                    b += 1
                else:
                    a += 1
            row = IDX[key(new)]
            ent.setdefault((row, col, a, b), 0)
            ent[(row, col, a, b)] += 1
    return sorted((r, c, a, b, m) for (r, c, a, b), m in ent.items())


def onepop():
    cfg_index = {tuple(sorted(c)): i for i, c in enumerate(CONFIGS)}
    L = [[Fraction(0)] * 8 for _ in range(8)]
    for col, cfg in enumerate(CONFIGS):
        n = len(cfg)
        for i in range(n):
            for j in range(i + 1, n):
                rest = [cfg[k] for k in range(n) if k not in (i, j)]
                rest.append((cfg[i][0] + cfg[j][0], cfg[i][1] + cfg[j][1]))
                L[col][col] -= 1
                if len(rest) >= 2:
                    L[cfg_index[tuple(sorted(rest))]][col] += 1
    return L


def matmul(A, B):
    n, m, p = len(A), len(B), len(B[0])
    return [[sum(A[i][k] * B[k][j] for k in range(m)) for j in range(p)] for i in range(n)]


def shifted(L, s):
    return [[L[i][j] + (s if i == j else 0) for j in range(8)] for i in range(8)]


def fmt(v):
    return repr(float(v))


def main():
    ent, diag = generator_entries()
    W44 = [[0] * 44 for _ in range(7)]
    for i, st in enumerate(STATES):
        for (d0, d1, _p) in st:
            W44[SLOT[(d0, d1)]][i] += 1
    W8 = [[0] * 8 for _ in range(7)]
    for i, cfg in enumerate(CONFIGS):
        for d in cfg:
            W8[SLOT[d]][i] += 1
    collapse = [max(b for b in range(8) if BASE[b] <= i) for i in range(44)]
    anc2 = [int(sum(1 for l in st if l == (1, 0, 0)) == 2) for st in STATES]
    anc11 = [int(sum(1 for l in st if l == (2, 0, 0)) == 1) for st in STATES]
    stationary = [i for i, st in enumerate(STATES) if len(st) == 2 and st[0][2] != st[1][2]]
    # ---- state -> (lane, slot) map of the 16-lane device kernel.  Each lane owns three states (a, b, c); an
    # off-diagonal entry whose column is owned by the same lane is served from registers ("local"), the others
    # ("remote") are read from shared memory: at most 3 / 2 / 3 remote entries for slot a / b / c.  Lanes are
    # paths a - b - c of the migration graph (b in the middle) or adjacent pairs (a, c) with a two-entry state
    # (or nothing) in slot b; slots left empty write to the pad rows 44..47.
    ell = [[e[1:] for e in ent if e[0] == r] for r in range(44)]
    lanes16 = [[3, 4, 5], [1, None, 7], [9, 10, 11], [12, 13, 14], [23, 24, 25], [26, 27, 28], [15, 16, 18], [17, 21, 22],
               [41, 42, 43], [29, 0, 30], [31, 2, 32], [33, 6, 34], [35, 8, 36], [37, None, 38], [39, None, 40], [19, None, 20]]
    assert sorted(v for ln in lanes16 for v in ln if v is not None) == list(range(44))
    RW = (3, 2, 3)

    def code_of(kind, cnt):
        return kind + {1: 0, 2: 4, 4: 8}[cnt]
    l16_row, l16_loc, rem_sets, pad = [], [], [], 44
    for ln in lanes16:
        rows, rems, locs = [], [], []
        for sl, r in enumerate(ln):
            if r is None:
                rows.append(pad)
                pad += 1
                rems.append([])
                locs.append([12, 12])
                continue
            rows.append(r)
            others = [ln[(sl + 1) % 3], ln[(sl + 2) % 3]]
            loc, rem = [12, 12], []
            for (c, k, n) in ell[r]:
                if c in others:
                    loc[others.index(c)] = code_of(k, n)
                else:
                    rem.append((c, code_of(k, n)))
            assert len(rem) <= RW[sl], (r, rem)
            rems.append(rem)
            locs.append(loc)
        l16_row.append(rows)
        rem_sets.append(rems)
        l16_loc.append(locs)
    assert pad <= 48
    # shared-memory position of a state: 16 * slot + lane, so that the 16 lanes of a group store one slot without a bank
    # conflict (8-byte words: bank pair = position mod 16 = owning lane).  The remote loads conflict when two lanes of
    # the group read different words owned by the same lane in the same load instruction; the ORDER of a lane's remote
    # entries and the address of its padding entries (coefficient 0: any word will do) are chosen by a seeded local
    # search to minimise the wavefronts per mat-vec.
    l16_pos = [0] * 48
    for ln, rows in enumerate(l16_row):
        for sl, r in enumerate(rows):
            l16_pos[r] = 16 * sl + ln

    def wavefronts(addrs):
        banks = {}
        for a in addrs:
            banks.setdefault(a % 16, set()).add(a)
        return max(len(v) for v in banks.values())

    def layout_cost(lay):
        tot = 0
        for sl in range(3):
            for e in range(RW[sl]):
                tot += wavefronts([l16_pos[lay[ln][sl][e][0]] for ln in range(16)])
        return tot
    import random
    rnd = random.Random(20261018)
    lay = [[rem_sets[ln][sl] + [(l16_row[ln][sl], 12)] * (3 - len(rem_sets[ln][sl])) for sl in range(3)] for ln in range(16)]
    best = layout_cost(lay)
    for _ in range(60000):
        ln, sl = rnd.randrange(16), rnd.randrange(3)
        cand = list(lay[ln][sl])
        if rnd.random() < 0.5:
            head = cand[:RW[sl]]
            rnd.shuffle(head)
            cand = head + cand[RW[sl]:]
        else:
            pads = [i for i in range(RW[sl]) if cand[i][1] == 12]
            if not pads:
                continue
            cand[rnd.choice(pads)] = (rnd.randrange(48), 12)
        old = lay[ln][sl]
        lay[ln][sl] = cand
        c = layout_cost(lay)
        if c <= best:
            best = c
        else:
            lay[ln][sl] = old
    l16_rem = lay
    l16_wavefronts = best + 3
    # ---- q = max |M_cc| needs only the Pareto-maximal diagonal multiplicity tuples (all rates are >= 0)
    tuples = sorted(set(tuple(d) for d in diag))
    qdiag = [t for t in tuples if not any(o != t and all(o[k] >= t[k] for k in range(4)) for o in tuples)]
    # ---- zero-migration runs in closed form.  Without migration M = la0 C0 + la1 C1 with C0, C1 the coalescence
    # generators of deme 0 / deme 1; they commute (lineages never change deme, so the two demes evolve independently) and
    # are diagonalisable with eigenvalues 0, -1, -3, -6 (minus the number of pairs among 0/1, 2, 3, 4 lineages; blocks of
    # equal diagonal are scalar).  Hence exp(sum_i M_i T_i) = sum_ab exp(-a X0 - b X1) G0_a G1_b, X = sum la T, and only
    # the 8 products with a "lineage budget" of at most four are non-zero.
    def cmat(kind):
        M = [[Fraction(0)] * 44 for _ in range(44)]
        for (r, c, k, n) in ent:
            if k == kind:
                M[r][c] += n
        for c in range(44):
            M[c][c] -= diag[c][kind]
        return M

    def mm44(A, B):
        return [[sum(A[i][k] * B[k][j] for k in range(44) if A[i][k] != 0) for j in range(44)] for i in range(44)]
    C0, C1 = cmat(0), cmat(1)
    assert mm44(C0, C1) == mm44(C1, C0)
    I44 = [[Fraction(int(i == j)) for j in range(44)] for i in range(44)]
    EV = (0, 1, 3, 6)

    def projector(C, ev):
        Pm = I44
        for o in EV:
            if o != ev:
                Pm = mm44(Pm, [[(C[i][j] + (o if i == j else 0)) / Fraction(o - ev) for j in range(44)] for i in range(44)])
        return Pm
    G0 = {a: projector(C0, a) for a in EV}
    G1 = {a: projector(C1, a) for a in EV}
    for C, G in ((C0, G0), (C1, G1)):
        assert [[sum(G[a][i][j] for a in EV) for j in range(44)] for i in range(44)] == I44
        for a in EV:
            assert mm44(C, G[a]) == [[-a * x for x in row] for row in G[a]]
    NM_AB = [(0, 0), (1, 0), (3, 0), (6, 0), (0, 1), (0, 3), (0, 6), (1, 1)]
    nm = []  # (row, col, ab index, value)
    for a in EV:
        for b in EV:
            Gab = mm44(G0[a], G1[b])
            nz = [(r, c, Gab[r][c]) for r in range(44) for c in range(44) if Gab[r][c] != 0]
            if (a, b) in NM_AB:
                nm += [(r, c, NM_AB.index((a, b)), v) for (r, c, v) in nz]
            else:
                assert not nz
    nm.sort()
    nm_rowptr = [sum(1 for e in nm if e[0] < r) for r in range(45)]
    L = onepop()
    # spectral projectors of L8: eigenvalues -6, -3, -1 (diagonal blocks are scalar => diagonalisable)
    G6 = [[v / 15 for v in row] for row in matmul(shifted(L, 3), shifted(L, 1))]
    G3 = [[v / -6 for v in row] for row in matmul(shifted(L, 6), shifted(L, 1))]
    G1 = [[v / 10 for v in row] for row in matmul(shifted(L, 6), shifted(L, 3))]
    for i in range(8):
        for j in range(8):
            assert G6[i][j] + G3[i][j] + G1[i][j] == (1 if i == j else 0)
            assert -6 * G6[i][j] - 3 * G3[i][j] - G1[i][j] == L[i][j]
    W8f = [[Fraction(v) for v in row] for row in W8]
    WG = [matmul(W8f, G) for G in (G6, G3, G1)]
    pulses = [pulse_entries(0), pulse_entries(1)]

    o = []
    o.append("// GENERATED by tools/gen_tables.py -- do not edit.  See that file for the derivation.")
    o.append("#pragma once")
    o.append("#define MISTI_NSTATE2 44")
    o.append("#define MISTI_NSTATE1 8")
    o.append("#define MISTI_GEN_NNZ %d" % len(ent))
    o.append("// {row, col, kind, count}")
    o.append("#define MISTI_GEN_ENTRIES_INIT { %s }" % ", ".join("{%d,%d,%d,%d}" % e for e in ent))
    o.append("#define MISTI_GEN_DIAG_INIT { %s }" % ", ".join("{%d,%d,%d,%d}" % tuple(d) for d in diag))
    # row-oriented (ELL) copy of the off-diagonal entries: 4 slots per row, {col, kind, count}, padded with count 0
    ell = [[e[1:] for e in ent if e[0] == r] for r in range(44)]
    assert max(len(r) for r in ell) == 4
    o.append("#define MISTI_ELL_WIDTH 4")
    o.append("#define MISTI_ELL_INIT { %s }" % ", ".join(
        "{" + ",".join("{%d,%d,%d}" % e for e in (row + [(r, 0, 0)] * (4 - len(row)))) + "}" for r, row in enumerate(ell)))
    o.append("// 16-lane layout: state (or pad row 44..47) per lane and slot; remote entries {col, code} per slot (3 each,")
    o.append("// slot b uses 2); local coefficient codes {from slot (s+1)%3, from slot (s+2)%3}; code = kind + 4 log2(count), 12 = none")
    o.append("// shared-memory position of each state in the 16-lane layout (16 * slot + lane); %d wavefronts per mat-vec and group" % l16_wavefronts)
    o.append("#define MISTI_L16_POS_INIT { %s }" % ",".join(str(v) for v in l16_pos))
    o.append("#define MISTI_L16_ROW_INIT { %s }" % ", ".join("{" + ",".join(str(v) for v in r) + "}" for r in l16_row))
    o.append("#define MISTI_L16_REM_INIT { %s }" % ", ".join(
        "{" + ",".join("{" + ",".join("{%d,%d}" % e for e in slot) + "}" for slot in lane) + "}" for lane in l16_rem))
    o.append("#define MISTI_L16_LOC_INIT { %s }" % ", ".join(
        "{" + ",".join("{%d,%d}" % tuple(slot) for slot in lane) + "}" for lane in l16_loc))
    o.append("// Pareto-maximal rows of MISTI_GEN_DIAG: max_c |M_cc| is attained on one of them")
    o.append("#define MISTI_QDIAG_N %d" % len(qdiag))
    o.append("#define MISTI_QDIAG_INIT { %s }" % ", ".join("{%d,%d,%d,%d}" % t for t in qdiag))
    o.append("// zero-migration runs: non-zero entries of the spectral projector products G0_a G1_b, row-major;")
    o.append("// ab index -> (a, b): 0 (0,0), 1 (1,0), 2 (3,0), 3 (6,0), 4 (0,1), 5 (0,3), 6 (0,6), 7 (1,1)")
    # ---- the run table as the 16-lane groups use it.  Any lane may compute any row (inputs and outputs go through shared
    # memory), so rows are packed into 16 lists of equal length (first fit decreasing); a list is walked in lock step by
    # the 16 lanes, entry k of lane l at index 16 k + l.  meta = shared-memory word of the column | slot of c_ab << 8 |
    # slot of e_ab << 16 (e_0 = 1 sits behind the record, slot 16) | "last entry of its row" << 24 | word of the row << 25.
    # The order of the rows in a list and of the entries in a row is chosen for few bank conflicts of the y[col] gather.
    by_row = [nm[nm_rowptr[r]:nm_rowptr[r + 1]] for r in range(44)]
    runlen = -(-len(nm) // 16)
    while True:
        bins = [[] for _ in range(16)]
        ok = True
        for r in sorted(range(44), key=lambda r: -len(by_row[r])):
            fits = [bn for bn in bins if sum(len(by_row[x]) for x in bn) + len(by_row[r]) <= runlen]
            if not fits:
                ok = False
                break
            fits[0].append(r)
        if ok:
            break
        runlen += 1

    def lane_list(bn):
        return [(e, i == len(by_row[r]) - 1) for r in bn for i, e in enumerate(by_row[r])]

    def gather_cost():
        lists = [lane_list(bn) for bn in bins]
        return sum(wavefronts([l16_pos[x[k][0][1]] for x in lists if len(x) > k]) for k in range(runlen))
    cost0 = gather_cost()
    bestc = cost0
    for _ in range(30000):
        if rnd.random() < 0.5:
            r = rnd.randrange(44)
            if len(by_row[r]) < 2:
                continue
            i, j = rnd.sample(range(len(by_row[r])), 2)
            by_row[r][i], by_row[r][j] = by_row[r][j], by_row[r][i]
            c = gather_cost()
            if c <= bestc:
                bestc = c
            else:
                by_row[r][i], by_row[r][j] = by_row[r][j], by_row[r][i]
        else:
            bn = bins[rnd.randrange(16)]
            if len(bn) < 2:
                continue
            i, j = rnd.sample(range(len(bn)), 2)
            bn[i], bn[j] = bn[j], bn[i]
            c = gather_cost()
            if c <= bestc:
                bestc = c
            else:
                bn[i], bn[j] = bn[j], bn[i]
    print("run table: %d entries per lane, gather wavefronts per group %d -> %d (minimum %d)" % (runlen, cost0, bestc, runlen))
    r16_val, r16_meta = [0.0] * (16 * runlen), [0] * (16 * runlen)
    for ln, bn in enumerate(bins):
        for k, (e, last) in enumerate(lane_list(bn)):
            ab = e[2]
            r16_val[16 * k + ln] = e[3]
            r16_meta[16 * k + ln] = l16_pos[e[1]] | (ab << 8) | ((16 if ab == 0 else 7 + ab) << 16) | (int(last) << 24) | (l16_pos[e[0]] << 25)
    o.append("// run table of the 16-lane layout: entry k of lane l at index 16 k + l (see tools/gen_tables.py)")
    o.append("#define MISTI_R16_LEN %d" % runlen)
    o.append("#define MISTI_R16_VAL_INIT { %s }" % ",".join(fmt(v) for v in r16_val))
    o.append("#define MISTI_R16_META_INIT { %s }" % ",".join("%du" % v for v in r16_meta))
    o.append("#define MISTI_NM_NNZ %d" % len(nm))
    o.append("#define MISTI_NM_ROWPTR_INIT { %s }" % ",".join(str(v) for v in nm_rowptr))
    o.append("#define MISTI_NM_COL_INIT { %s }" % ",".join(str(e[1]) for e in nm))
    o.append("#define MISTI_NM_AB_INIT { %s }" % ",".join(str(e[2]) for e in nm))
    o.append("#define MISTI_NM_VAL_INIT { %s }" % ",".join(fmt(e[3]) for e in nm))
    o.append("#define MISTI_W44_INIT { %s }" % ", ".join("{" + ",".join(str(v) for v in row) + "}" for row in W44))
    o.append("#define MISTI_W8_INIT { %s }" % ", ".join("{" + ",".join(str(v) for v in row) + "}" for row in W8))
    o.append("#define MISTI_COLLAPSE_INIT { %s }" % ",".join(str(v) for v in collapse))
    o.append("#define MISTI_ANC2_INIT { %s }" % ",".join(str(v) for v in anc2))
    o.append("#define MISTI_ANC11_INIT { %s }" % ",".join(str(v) for v in anc11))
    o.append("#define MISTI_NSTATIONARY %d" % len(stationary))
    o.append("#define MISTI_STATIONARY_INIT { %s }" % ",".join(str(v) for v in stationary))
    for s in (0, 1):
        o.append("#define MISTI_PULSE%d_NNZ %d" % (s, len(pulses[s])))
        o.append("// {row, col, a, b, multiplicity}: value = multiplicity * (1-r)^a * r^b")
        o.append("#define MISTI_PULSE%d_INIT { %s }" % (s, ", ".join("{%d,%d,%d,%d,%d}" % e for e in pulses[s])))
        rowptr = [sum(1 for e in pulses[s] if e[0] < r) for r in range(45)]
        o.append("#define MISTI_PULSE%d_ROWPTR_INIT { %s }" % (s, ",".join(str(v) for v in rowptr)))
    o.append("#define MISTI_L8_INIT { %s }" % ", ".join("{" + ",".join(fmt(v) for v in row) + "}" for row in L))
    for name, M in zip(("WG6", "WG3", "WG1"), WG):
        o.append("#define MISTI_%s_INIT { %s }" % (name, ", ".join("{" + ",".join(fmt(v) for v in row) + "}" for row in M)))
    with open(OUT, "w") as f:
        f.write("\n".join(o) + "\n")
    print("wrote", OUT, "gen nnz", len(ent), "pulse nnz", len(pulses[0]), len(pulses[1]), "stationary", stationary)


if __name__ == "__main__":
    main()
