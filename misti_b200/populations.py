"""TwoPopulations / OnePopulation: drop-ins for the reference's lineage-chain classes.

Mirror of TwoPopulations(l1, l2, m1, m2) (reference TwoPopulations.py:56-377) and OnePopulation(l1)
(OnePopulation.py:37-178) for code that builds the chains by hand (TestModel-style scripts,
MigrationInference.JAFSpectrum in the reference).  The generator, the state -> JSFS branch counts,
the pulse map and the ancient-sample reset are produced BY THE DEVICE from the same constant tables
the evaluation kernels use (misti_generator / misti_state_to_jaf / misti_pulse / misti_ancient_reset
in include/misti_b200.h); only index bookkeeping (state <-> index, deletion / re-insertion of the
stationary states in the zero-migration case) is done here.  The batched evaluation path does not
use these classes: the kernels never form or invert M (see csrc/misti_jsfs.cuh).
"""
import sys

import numpy as np

from .engine import default_engine

_BASE = [0, 9, 15, 23, 29, 33, 37, 41, 44]
# lineage configurations by block, as (d0, d1) pairs in the reference's canonical order
_CONFIGS = [[(1, 0), (1, 0), (0, 1), (0, 1)], [(2, 0), (0, 1), (0, 1)], [(1, 1), (1, 0), (0, 1)], [(0, 2), (1, 0), (1, 0)],
            [(2, 1), (0, 1)], [(1, 2), (1, 0)], [(2, 0), (0, 2)], [(1, 1), (1, 1)]]


class lineage:
    """(d0, d1, pop): descendants in genome 1 / genome 2, current deme (TwoPopulations.py:50-54)."""

    def __init__(self, d0, d1, pop=0):
        self.d0, self.d1, self.pop = d0, d1, pop


def _index_to_demes(ind):
    """block number and deme of each lineage (in _CONFIGS order) for a two-population state index."""
    blk = max(b for b in range(8) if _BASE[b] <= ind)
    k = ind - _BASE[blk]
    if blk == 0:  # index = i + 3 j: j genome-1 singletons and i genome-2 singletons in deme 1
        j, i = divmod(k, 3)
        return blk, [int(j >= 2), int(j >= 1), int(i >= 2), int(i >= 1)]
    if blk in (1, 3):
        a, n = divmod(k, 3)
        return blk, [a, int(n >= 2), int(n >= 1)]
    if blk == 2:
        return blk, [k >> 2 & 1, k >> 1 & 1, k & 1]
    if blk in (4, 5, 6):
        return blk, [k >> 1 & 1, k & 1]
    return blk, [int(k >= 2), int(k >= 1)]


class TwoPopulations:
    def __init__(self, l1, l2, m1, m2, engine=None):
        if m1 < 0 or m2 < 0:
            self.PrintError("TwoPopulations", "migration rates cannot be negative")
        if l1 < 0 or l2 < 0:
            self.PrintError("TwoPopulations", "coalescent rates cannot be negative")
        self.mu = [m1, m2]
        self.la = [l1, l2]
        self.P0 = []
        self.Msize = 44
        self._engine = engine
        self.stationary = []
        if m1 + m2 == 0:
            for i in range(self.Msize):
                st = self.MapIndToState(i)
                if len(st) == 2 and st[0].pop != st[1].pop:
                    self.stationary.append(i)

    def _eng(self):
        if self._engine is None:
            self._engine = default_engine()
        return self._engine

    def MSize(self):
        return self.Msize - len(self.stationary)

    def StateNum(self):
        return self.Msize

    def PrintError(self, func, text):
        print("TwoPopulations class error in function", func + "():", text)
        sys.exit(0)

    def MapIndToState(self, ind):
        blk, demes = _index_to_demes(ind)
        return [lineage(d[0], d[1], p) for d, p in zip(_CONFIGS[blk], demes)]

    def MapStateToInd(self, state):
        key = sorted((l.d0, l.d1, l.pop) for l in state)
        if sum(k[0] for k in key) != 2 or sum(k[1] for k in key) != 2:
            self.PrintError("CheckState", "CheckState() not passed: expected number of lineages is 2 and 2")
        if len(key) < 2:
            return self.Msize
        for i in range(self.Msize):
            if sorted((l.d0, l.d1, l.pop) for l in self.MapIndToState(i)) == key:
                return i
        self.PrintError("MapStateToInd", "unknown state")

    def StateToJAF(self, sti):
        return [int(v) for v in self._eng().state_to_jaf()[sti]]

    def SetMatrix(self):
        M = self._eng().generator(self.la[0], self.la[1], self.mu[0], self.mu[1])
        M = np.delete(M, self.stationary, 0)
        M = np.delete(M, self.stationary, 1)
        return np.asmatrix(M)

    def SetInitialConditions(self, P0):
        self.P0 = P0
        if self.mu[0] + self.mu[1] == 0:
            P0 = np.delete(P0, self.stationary)
        return P0

    def AncientSampleP0(self, P0):
        return [float(v) for v in self._eng().ancient_reset(P0)]

    def PulseMigration(self, P0, migRate, pop1):
        return [float(v) for v in self._eng().pulse(P0, migRate, pop1)]

    def _class(self, ind):
        st = self.MapIndToState(ind)
        return (sum(l.d0 * l.pop for l in st), sum(l.d1 * l.pop for l in st))

    def _restore(self, vec, scale):
        # re-insert the stationary states by conservation of the mass of their class (TwoPopulations.py:264-309)
        if len(vec) == self.Msize:
            return vec
        if len(vec) != self.Msize - len(self.stationary):
            self.PrintError("UpdateInitialConditions", "unexpected length of vector " + str(len(vec)))
        full = np.asarray(vec, dtype=float).ravel()
        for ind in self.stationary:
            full = np.insert(full, ind, 0)
        out = [float(v) for v in full]
        classes = [self._class(i) for i in range(self.Msize)]
        for ind in self.stationary:
            for i in range(self.Msize):
                if classes[i] == classes[ind]:
                    out[ind] += scale * self.P0[i] - full[i]
        return out

    def UpdateInitialConditions(self, P0):
        return self._restore(P0, 1.0)

    def UpdateIntegral(self, integralP, T):
        return self._restore(integralP, T)


class OnePopulation:
    def __init__(self, l1, engine=None):
        if l1 < 0:
            self.PrintError("OnePopulation", "coalescent rates cannot be negative")
        self.l1 = l1
        self.Msize = 8
        self._engine = engine

    def _eng(self):
        if self._engine is None:
            self._engine = default_engine()
        return self._engine

    def MSize(self):
        return self.Msize

    def StateNum(self):
        return self.Msize

    def PrintError(self, func, text):
        print("OnePopulation class error in function", func + "():", text)
        sys.exit(0)

    def MapIndToState(self, ind):
        return [lineage(d[0], d[1]) for d in _CONFIGS[ind]]

    def MapStateToInd(self, state):
        key = sorted((l.d0, l.d1) for l in state)
        for i, cfg in enumerate(_CONFIGS):
            if sorted(cfg) == key:
                return i
        return self.Msize

    def StateToJAF(self, sti):
        return [int(v) for v in self._eng().state_to_jaf(one_pop=True)[sti]]

    def SetMatrix(self):
        return np.asmatrix(self._eng().generator(self.l1, one_pop=True))

    def SetInitialConditions(self, P0):
        return P0

    def UpdateInitialConditions(self, P0):
        return P0

    def UpdateIntegral(self, integralP, T):
        return integralP
