"""Input / output either side of the hot path: PSMC pairs -> merged grid, joint SFS files, bootstrap rows,
the `.mi` result file.  Host-side text handling only (runs once per process); nothing here is accelerated.

Functional restatement of migrationIO.Units / ReadPSMCFile / ReadPSMC (migrationIO.py:100-176, 183-295),
ReadPSMC1 (:297-344, with the parts of psmc.py it uses), ReadJAFS (:557-656), BootstrapJAFS (:506-524),
utils/generateJSFS_bs.py:39-48, OutputMigration (:346-375) and ReadMigration (:377-505, minus plotting), without the reference's process-global state: units are a value object (the reference mutates
class statics), readers return fresh objects (the reference's JAFS() default list accumulates rows across
calls, migrationIO.py:39), bootstrap resampling takes an explicit seed (the reference uses the unseeded
global `random`).  Reference-named aliases (ReadPSMC, ReadJAFS, BootstrapJAFS, OutputMigration) are kept.
"""
import math
import random
import sys


class Units:
    """mutation rate per bp per generation, PSMC bin size, reference N0, generation time, heterozygosity loss."""

    def __init__(self, mutRate=1.25e-8, binsize=100, N0=10000, genTime=1, hetloss=(0.0, 0.0)):
        self.mutRate, self.binsize, self.N0, self.genTime = mutRate, binsize, N0, genTime
        self.hetloss1, self.hetloss2 = 0.0, 0.0
        self.SetHetLoss(hetloss)

    @classmethod
    def from_file(cls, fn, hetloss=(0.0, 0.0)):
        """`name=value` lines (setunits.txt); unknown / unreadable entries keep their defaults."""
        u = cls(hetloss=hetloss)
        try:
            with open(fn) as f:
                for line in f:
                    kv = line.split("=")
                    if len(kv) == 2 and kv[0] in ("mutRate", "binsize", "N0", "genTime"):
                        try:
                            setattr(u, kv[0], float(kv[1]))
                        except ValueError:
                            print("Cannot read %s entry from file, using default or previous values" % kv[0])
        except OSError:
            print("Units input file not found, using default values.")
        return u

    def SetHetLoss(self, hl):
        for k, v in enumerate(hl or ()):
            if v is None:
                continue
            if not (0.0 <= v < 1.0):
                sys.stderr.write("Hetloss should be between 0 and 1.\n")
                sys.exit(0)
            setattr(self, "hetloss%d" % (k + 1), v)

    def PrintUnits(self):
        print("Units: mutation rate =", self.mutRate, "\tbinsize =", self.binsize, "\tN0 =", self.N0, "\tgeneration time =", self.genTime)


class InputData:
    """What MiSTI.py hands to MigrationInference (migrationIO.InputData, migrationIO.py:46-63)."""

    def __init__(self, times, lambdas, scaleTime, theta, divTime=-1, scaleEPS=1.0, rho=None, sampleDateDiscr=0, Tpsmc=None,
                 mi=None, pu=None):
        self.times, self.lambdas = times, lambdas
        self.divergenceTime = divTime
        self.scaleTime, self.theta, self.scaleEPS, self.rho = scaleTime, theta, scaleEPS, rho
        self.sampleDateDiscr, self.Tpsmc = sampleDateDiscr, Tpsmc
        self.mi, self.pu = mi, pu


class JAFS:
    def __init__(self, jafs=None, pop1=None, pop2=None):
        self.jafs = [] if jafs is None else jafs
        self.pop1, self.pop2 = pop1, pop2


def read_psmc_file(fn, RD=-1):
    """One PSMC output: interval start times t_k and relative sizes lambda_k of round RD (-1 = last),
    theta and rho of that round (migrationIO.ReadPSMCFile, migrationIO.py:183-222)."""
    with open(fn) as f:
        lines = [ln.split() for ln in f if ln.split()]
    rounds = [int(ln[1]) for ln in lines if ln[0] == "RD"]
    if not rounds:
        print("Corrupted or empty input file")
        sys.exit(0)
    if RD == -1 or RD > max(rounds):
        RD = max(rounds)
    Tk, Lk, th, rh = [], [], 0.0, 0.0
    inside = False
    for ln in lines:
        if ln[0] == "RD":
            if inside:
                break
            inside = int(ln[1]) == RD
        elif inside:
            if ln[0] == "TR":
                th, rh = float(ln[1]), float(ln[2])
            elif ln[0] == "RS":
                Tk.append(float(ln[2]))
                Lk.append(float(ln[3]))
            elif ln[0] == "PA":
                break
    return [Tk, Lk, RD, th, rh]


def read_psmc(fn1, fn2, sampleDate=0.0, RD=-1, units=None):
    """Merge two PSMC trajectories onto one time grid (migrationIO.ReadPSMC, migrationIO.py:224-295):
    each genome's times and sizes are rescaled by theta_g / theta (theta_g inflated by its heterozygosity loss),
    genome 2 is shifted by the sampling date (with a dummy unit-rate interval before it), the union of the
    breakpoints is the grid and each genome's rate 1/lambda is held piecewise constant on it."""
    u = units if units is not None else Units()
    d1, d2 = read_psmc_file(fn1, RD), read_psmc_file(fn2, RD)
    d1[3] = d1[3] / (1.0 - u.hetloss1)
    d2[3] = d2[3] / (1.0 - u.hetloss2)
    theta = 4.0 * u.binsize * u.mutRate * u.N0
    scaleTime = 2 * u.genTime * u.N0
    for d in (d1, d2):
        d[0] = [v * d[3] / theta for v in d[0]]
        d[1] = [v * d[3] / theta for v in d[1]]
    sdResc = sampleDate / 2 / u.N0 / u.genTime
    if sdResc > 0:
        d2[0] = [0.0] + [v + sdResc for v in d2[0]]
        d2[1] = [1.0] + d2[1]
    Tk = sorted(d1[0] + d2[0][1:])
    if sdResc not in Tk:
        sys.stderr.write("Unexpected error in ReadPSMC(). Get in touch with the author.\n")
        sys.exit(0)
    sampleDateDiscr = Tk.index(sdResc)
    rates, Tpsmc = [], []
    for d in (d1, d2):
        L, marks, j = [], [0], 0
        for i in range(len(d[0]) - 1):
            while Tk[j] < d[0][i + 1]:
                L.append(1.0 / d[1][i])
                j += 1
            marks.append(j)
        L += [1.0 / d[1][-1]] * (len(Tk) - len(L))
        marks.append(len(Tk))
        rates.append(L)
        Tpsmc.append(marks)
    lambdas = [[a, b] for a, b in zip(rates[0], rates[1])]
    times = [b - a for a, b in zip(Tk[:-1], Tk[1:])]
    return InputData(times, lambdas, scaleTime, theta, scaleEPS=1, rho=d1[4] * theta / d1[3], sampleDateDiscr=sampleDateDiscr,
                     Tpsmc=Tpsmc)


def psmc_pattern(fn):
    """Free-parameter pattern of a PSMC run as the list of its segment lengths, e.g. `MM pattern:4+25*2+4+6,` ->
    [4, 2 x 25, 4, 6] (psmc.PSMC.ReadPSMCFile, psmc.py:51-61)."""
    segs = []
    with open(fn) as f:
        for ln in f:
            w = ln.split()
            if len(w) > 1 and w[0] == "MM" and w[1].startswith("pattern"):
                for part in w[1][:-1].split(":")[1].split("+"):
                    v = [int(x) for x in part.split("*")]
                    segs += [v[0]] if len(v) == 1 else [v[1]] * v[0]
    return segs


def _pieces(t, eps, t1, t2):
    """The pieces [tl, tu) x size of the piecewise-constant trajectory (t, eps) inside [t1, t2); t2 may be inf."""
    inf = float("inf")
    k = 0
    while k < len(t) and t[k] <= t1:
        k += 1
    k -= 1
    while k < len(t) and t[k] < t2:
        nxt = t[k + 1] if k + 1 < len(t) else inf
        yield max(t1, t[k]), min(t2, nxt), eps[k]
        k += 1


def psmc_mean_size(t, eps, t1, t2):
    """Size of a constant population with the same coalescence hazard over [t1, t2): the time-weighted harmonic mean
    (psmc.PSMC.AverageCoalescentRate, psmc.py:97-119; same order of operations)."""
    if t1 > t2:
        sys.exit(1)
    hazard, span = 0.0, 0.0
    for tl, tu, e in _pieces(t, eps, t1, t2):
        hazard += tu / e - tl / e
        span += tu - tl
    return span / hazard


def psmc_tail_size(t, eps, t1):
    """Size of a constant population with the same expected coalescence time beyond t1 (psmc.PSMC.FitCoalescentTime with
    t2 = inf, psmc.py:121-146, incl. its bounded least-squares solve of `size = expected time - t1`)."""
    from scipy.optimize import least_squares
    inf = float("inf")
    et, lognc = 0.0, 0.0
    for tl, tu, e in _pieces(t, eps, t1, inf):
        ru, rl = tu / e, tl / e
        vu = 0.0 if ru == inf else (ru + 1.0) * math.exp(rl - ru)
        et += math.exp(lognc) * ((rl + 1.0) - vu) * e
        lognc -= ru - rl
    et = et / (1.0 - math.exp(lognc))
    return least_squares(lambda size: (et - t1) - size, 1.0, bounds=(0.0, inf), ftol=4e-16, xtol=4e-16, gtol=4e-16).x[0]


def read_psmc1(fn1, fn2, RD=-1, divergenceTime=-1, units=None):
    """The reference's second input mode (MiSTI.py -pm 1; migrationIO.ReadPSMC1, migrationIO.py:297-344): both trajectories
    are rescaled to the common theta, the grid is the average of their collapsed (one point per pattern segment) grids with
    the split time -- given in YEARS -- inserted, and each trajectory is re-estimated on it (mean size per interval, fitted
    size beyond the last point).  Returns the intervals, the size pairs, and the index of the split time in the grid
    (-1 when none was given).  Like the reference, this mode ignores heterozygosity loss and the sampling date."""
    u = units if units is not None else Units()
    if u.hetloss1 != 0.0 or u.hetloss2 != 0.0:
        print("Hetloss id not implemented in this version.")
    theta = 4.0 * u.binsize * u.mutRate * u.N0
    scaleTime = 2 * u.genTime * u.N0
    traj, collapsed = [], []
    for fn in (fn1, fn2):
        t, eps, _, th, _ = read_psmc_file(fn, RD)
        t = [v * th / theta for v in t]
        eps = [v * th / theta for v in eps]
        traj.append((t, eps))
        starts, k = [], 0
        for n in psmc_pattern(fn):
            starts.append(t[k])
            k += n
        collapsed.append(starts)
    if len(collapsed[0]) != len(collapsed[1]):
        sys.exit(1)
    split = None if divergenceTime == -1 else divergenceTime / scaleTime
    Tk = sorted(set(([] if split is None else [split]) + [(a + b) / 2.0 for a, b in zip(collapsed[0], collapsed[1])]))
    sizes = []
    for t, eps in traj:
        sizes.append([psmc_mean_size(t, eps, a, b) for a, b in zip(Tk[:-1], Tk[1:])] + [psmc_tail_size(t, eps, Tk[-1])])
    return InputData([b - a for a, b in zip(Tk[:-1], Tk[1:])], [[a, b] for a, b in zip(sizes[0], sizes[1])], scaleTime, theta,
                     divTime=-1 if split is None else Tk.index(split))


def read_jafs(fn, silent_mode=True):
    """Joint SFS file (migrationIO.ReadJAFS, migrationIO.py:557-608): header lines starting with '#', an optional
    column line starting with 'total', then rows of 8 TAB-separated numbers
    [total sites, 0100, 1100, 0001, 0101, 1101, 0011, 0111].  Files of a format version < 1 go to the old reader."""
    out = JAFS()
    with open(fn) as f:
        lines = [ln.rstrip("\n") for ln in f]
    if not lines or not lines[0].startswith(("#MiSTI_JSFS", "#MiSTI_JAF", "#Migration_JAF")):
        sys.stderr.write("Corrupted JSFS file header.\n")
        sys.exit(0)
    if float(lines[0].split(" ")[2]) < 1:
        return _read_jafs_v0(lines, silent_mode)
    if not lines[0].startswith("#MiSTI_JSFS"):  # the two older names only exist with format versions < 1
        sys.stderr.write("Corrupted JSFS file header.\n")
        sys.exit(0)
    for ln in lines[1:]:
        if ln.startswith("#"):
            if ln[1:5] in ("pop1", "pop2"):
                pars = ln.split("\t")
                if len(pars) != 2:
                    sys.stderr.write("Corrupted JSFS file header.\n")
                    sys.exit(0)
                setattr(out, ln[1:5], pars[1])
                if not silent_mode:
                    print(ln[1:5] + "\t", pars[1])
            continue
        if ln.startswith("total") or ln == "":
            continue
        cols = ln.split("\t")
        if len(cols) != 8:
            sys.stderr.write("Unexpected line. Expected an entry for JSFS with eight TAB-separated columns.\n")
            sys.exit(0)
        out.jafs.append([float(v) for v in cols])
    return out


def _read_jafs_v0(lines, silent_mode):
    """Format versions < 1 (migrationIO.ReadJAFS_old, migrationIO.py:610-656): header fields separated by a blank, then
    exactly eight lines `label<TAB>integer count` -- one spectrum, no chunk rows."""
    out, counts = JAFS(), []
    for ln in (x.rstrip() for x in lines):
        if ln.startswith("#") and not counts:
            w = ln.split(" ")
            if ln[1:10] == "MiSTI_JAF" or ln[1:14] == "Migration_JAF":
                if len(w) < 3:
                    sys.stderr.write("Corrupted JAF file header.\n")
                    sys.exit(0)
                if not silent_mode:
                    print("JAFS format version:", w[2])
            elif ln[1:5] in ("pop1", "pop2"):
                if len(w) != 2:
                    sys.stderr.write("Corrupted JAF file header.\n")
                    sys.exit(0)
                setattr(out, ln[1:5], w[1])
                if not silent_mode:
                    print(ln[1:5] + "\t", w[1])
            continue
        w = ln.split("\t")
        if len(w) != 2:
            sys.stderr.write("Unexpected line. Expected an entry for JAFS with two TAB-separated columns.\n")
            sys.exit(0)
        counts.append(int(w[1]))
    if len(counts) != 8:
        print("Unexpected number of lines in the JAFS file.")
        sys.exit(0)
    out.jafs.append(counts)
    return out


def column_sums(rows):
    """The data spectrum MiSTI.py uses with -bs -1: the sum of all chunk rows (MiSTI.py:172-176)."""
    s = [0 for _ in range(8)]
    for r in rows:
        s = [v + w for v, w in zip(s, r)]
    return s


def bootstrap_jafs(rows, rng, normalize=False):
    """One bootstrap replicate: chunk rows drawn with replacement until the genome length is reached
    (migrationIO.BootstrapJAFS, migrationIO.py:506-524).  rng: random.Random."""
    for el in rows:
        if len(el) != 8:
            sys.stderr.write("Cannot use provided SFS for bootstrap.\n")
            sys.exit(0)
    genomeLen = sum(el[0] for el in rows)
    seg = sum(sum(el[1:]) for el in rows)
    sfs = [0 for _ in range(8)]
    while sfs[0] < genomeLen:
        pick = rows[rng.randint(0, len(rows) - 1)]
        sfs = [a + b for a, b in zip(sfs, pick)]
    if normalize:
        k = seg / sum(sfs[1:])
        sfs = [v * k for v in sfs]
    return sfs


def generate_bootstrap(rows, n, seed=0):
    """Row 0 = the data total, rows 1..n = bootstrap replicates (utils/generateJSFS_bs.py:39-48), seeded."""
    rng = random.Random(seed)
    return [column_sums(rows)] + [bootstrap_jafs(rows, rng) for _ in range(n)]


def write_jafs(rows, pop1=None, pop2=None, file=None):
    """A joint SFS file (migrationIO.PrintJAFSFile, migrationIO.py:526-555): header, optional population names, the column
    line, then one row per spectrum -- 8 numbers, or 7 with the total prepended; a single flat spectrum is one row.
    `file`: an open text file (default: stdout, as the reference prints).  read_jafs reads it back."""
    out = sys.stdout if file is None else file
    print("#MiSTI_JSFS version 1.0", file=out)
    for tag, name in (("#pop1", pop1), ("#pop2", pop2)):
        if name:
            print(tag, str(name).strip("\n\r"), sep="\t", file=out)
    print("\t".join(["total", "0100", "1100", "0001", "0101", "1101", "0011", "0111"]), file=out)
    if not isinstance(rows, list):
        sys.stderr.write("Unexpected SFS value: should be a list of a list of lists\n")
        sys.exit(0)
    if not isinstance(rows[0], list):
        rows = [rows]
    for sfs in rows:
        if len(sfs) == 7:
            sfs = [sum(sfs)] + list(sfs)
        elif len(sfs) != 8:
            print("Unexpected SFS entry.")
            sys.exit(0)
        print("\t".join(str(v) for v in sfs), file=out)


def output_migration(fout, mu, Migration, scaleTime=1, scaleEPS=1):
    """The `.mi` result file, format "#MiSTI2 ver 0.4" (migrationIO.OutputMigration, migrationIO.py:346-375)."""
    llh = Migration.llh if len(mu) == 0 else Migration.JAFSLikelihood(mu)
    times = [sum(Migration.times[0:i]) for i in range(len(Migration.times) + 1)]
    tot = sum(Migration.dataJAFS)
    out = ["#MiSTI2 ver 0.4", "LK\t" + str(llh), "ST\t" + str(Migration.splitT), "SD\t" + str(Migration.sampleDate),
           "TR\t" + str(Migration.thrh[0]) + "\t" + str(Migration.thrh[1]), "SFS\t" + "\t".join(map(str, Migration.JAFS)),
           "DSF\t" + "\t".join(str(v / tot) for v in Migration.dataJAFS), "SCT\t" + str(scaleTime), "SCE\t" + str(scaleEPS)]
    for i, t in enumerate(times):
        row = ["RS", str(t), str(1.0 / Migration.lc[i][0]), str(1.0 / Migration.lc[i][1]), str(1.0 / Migration.lh[i][0]),
               str(1.0 / Migration.lh[i][1]), str(Migration.mi[i][0]), str(Migration.mi[i][1])]
        if i < Migration.splitT:
            for val in Migration.Pr[i]:
                row += [str(val[0]), str(val[1])]
        out.append("\t".join(row))
    text = "\n".join(out) + "\n"
    if fout == "":
        print(text)
    else:
        with open(fout, "w") as f:
            f.write(text)


class MigData:
    """What read_migration returns (migrationIO.MigData, migrationIO.py:65-98) plus the columns the reference parses and
    then only plots: migration rates per interval (mu1, mu2) and the lineage-state probabilities (pr11, pr22, pr12)."""

    def __init__(self):
        self.llh = self.splitT = self.migStart = self.migEnd = self.thrh = self.mi = self.sampleDate = self.jaf = None
        self.times, self.lambda1, self.lambda2, self.lambdah1, self.lambdah2 = [], [], [], [], []
        self.mu1, self.mu2 = [], []
        self.pr11, self.pr22, self.pr12 = [[], []], [[], []], [[], []]


def read_migration(fmigr, scaleTime=1, scaleEPS=1):
    """A `.mi` result file back into numbers (migrationIO.ReadMigration, migrationIO.py:377-505, without its plotting
    branch): times in units of scaleTime, corrected / PSMC-apparent RATES lambda = 1 / size / scaleEPS.  Files of the
    current family ("#MiSTI2 ver >= 0.3") carry their own SCT / SCE lines, which replace the arguments from the line on
    where they stand; the older family has neither, nor the PSMC columns, but a single band (MS / ME / MU)."""
    d = MigData()
    with open(fmigr) as f:
        head = next(f).rstrip().split(" ")
        version = float(head[2])
        print("Format version: ", version)
        if version < 0.3:
            sys.stderr.write("File version is not supported anymore.\n")
            sys.exit(0)
        current = head[0] == "#MiSTI2"
        for ln in f:
            w = ln.split("\t")
            key = w[0]
            if key == "LK":
                d.llh = float(w[1])
            elif key == "ST":
                d.splitT = int(w[1])
            elif key == "SD":
                d.sampleDate = int(w[1])
            elif key == "TR":
                d.thrh = [float(w[1]), float(w[2])]
            elif key == "SFS":
                d.jaf = [float(v) for v in w[1:]]
            elif key == "SCT" and current:
                scaleTime = float(w[1])
            elif key == "SCE" and current:
                scaleEPS = float(w[1])
            elif key == "MS" and not current:
                d.migStart = int(w[1])
            elif key == "ME" and not current:
                d.migEnd = int(w[1])
            elif key == "MU" and not current:
                d.mi = [float(w[1]), float(w[2])]
            elif key == "RS":
                d.times.append(float(w[1]) * scaleTime)
                d.lambda1.append(1.0 / float(w[2]) / scaleEPS)
                d.lambda2.append(1.0 / float(w[3]) / scaleEPS)
                if not current:
                    continue
                c = 4
                if version >= 0.4:
                    d.lambdah1.append(1.0 / float(w[4]) / scaleEPS)
                    d.lambdah2.append(1.0 / float(w[5]) / scaleEPS)
                    c = 6
                d.mu1.append(float(w[c]))
                d.mu2.append(float(w[c + 1]))
                pr = [float(v) for v in w[c + 2:c + 8]] if len(w) > c + 2 else [0] * 6
                for k, dst in enumerate((d.pr11, d.pr22, d.pr12)):
                    dst[0].append(pr[2 * k])
                    dst[1].append(pr[2 * k + 1])
    return d


# reference-named aliases
ReadPSMCFile, ReadPSMC, ReadPSMC1, ReadJAFS, OutputMigration, ReadMigration = (read_psmc_file, read_psmc, read_psmc1, read_jafs,
                                                                                output_migration, read_migration)


def BootstrapJAFS(Jafs, normalize=False, rng=None):
    return bootstrap_jafs(Jafs.jafs, rng if rng is not None else random.Random(), normalize)


def read_ms(argument_string):
    """An ms command line -> the model it describes, as TestModel.py consumes it (migrationIO.ReadMS, migrationIO.py:659-766):
    InputData with times (interval lengths in units of 2 N0 generations), lambdas (1 / population size per deme),
    divergenceTime = index of the split interval, mi = [-mi pop start end rate 0] and pu = [-pu pop time rate 0].

    Understood, with the reference's conventions (it warns that it makes many assumptions about the command line):
    `-n i x`, `-en t i x`, `-eN t x` (sizes), `-em t i j r` (from time t on, migration rate r for deme i -- the target j is
    not looked at -- until the next -em of the same deme or the split; MiSTI's rate is 2 r), `-es t i p` (pulse: a
    fraction 1 - p of deme i's lineages moves), `-ej t i j` with i <= 2 (the split: deme i joins, and afterwards carries
    the other deme's sizes).  Everything else is skipped token by token.  A size of exactly 0 means "not set here"."""
    args = argument_string.split(" ")
    sizes = [{0.0: 1.0}, {0.0: 1.0}]   # per deme: time -> size
    bands = [{}, {}]                   # per deme: time -> rate
    pulses = {}                        # time -> (rate, deme)
    split_time, joining = 0, None
    i = 0
    while i < len(args):
        a = args[i]
        if a in ("-n", "-en"):
            off = 0 if a == "-n" else 1
            t = 0.0 if a == "-n" else float(args[i + 1])
            deme, x = int(args[i + 1 + off]), float(args[i + 2 + off])
            if deme not in (1, 2):
                print("Population id should be 1 or 2.")
                print(*args[i:i + 3 + off])
                sys.exit(0)
            sizes[deme - 1][t] = x
            i += 3 + off
        elif a == "-eN":
            t, x = float(args[i + 1]), float(args[i + 2])
            sizes[0][t] = x
            sizes[1][t] = x
            i += 3
        elif a == "-em":
            t, deme, r = float(args[i + 1]), int(args[i + 2]), float(args[i + 4])
            bands[deme - 1][t] = r
            i += 5
        elif a == "-es":
            t, deme, p = float(args[i + 1]), int(args[i + 2]), float(args[i + 3])
            pulses[t] = (1 - p, deme)
            i += 4
        elif a == "-ej":
            if int(args[i + 2]) <= 2:
                split_time, joining = float(args[i + 1]), int(args[i + 2]) - 1
            i += 4
        else:
            i += 1
    if joining is None:
        print("Populations should be merged. (-ej [time] 2 1)")
        sys.exit(0)
    events = set([split_time]) | set(pulses)
    for k in (0, 1):
        events |= set(sizes[k]) | set(bands[k])
    grid = sorted(events)
    index = {t: n for n, t in enumerate(grid)}
    split_index = index[split_time]
    # sizes held constant between their change points; after the split the joining deme carries the other one's
    table = [[0, 0] for _ in grid]
    for k in (0, 1):
        for t, x in sizes[k].items():
            table[index[t]][k] = x
        current = 0
        for row in table:
            if row[k] == 0:
                row[k] = current
            else:
                current = row[k]
    for row in table[split_index:]:
        row[joining] = row[1 - joining]
    mi = []
    for k in (0, 1):
        for t, r in bands[k].items():
            mi.append([k + 1, index[t], split_index, 2 * r, 0])
    mi.sort(key=lambda el: (el[0], el[1]))
    for cur, nxt in zip(mi[:-1], mi[1:]):
        if cur[0] == nxt[0]:
            cur[2] = nxt[1]
    pu = [[deme, index[t], rate, 0] for t, (rate, deme) in pulses.items()]
    times = [2 * (b - a) for a, b in zip(grid[:-1], grid[1:])]
    lambdas = [[1.0 / row[0], 1.0 / row[1]] for row in table]
    return InputData(times, lambdas, 1.0, 1.0, divTime=split_index, mi=mi, pu=pu)


ReadMS, PrintJAFSFile = read_ms, write_jafs
