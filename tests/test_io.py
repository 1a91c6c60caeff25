"""Loaders either side of the path (misti_b200/io.py) against what the reference's migrationIO produced from the
same files (tests/golden/datasets.json).  CPU only."""
import os
import random

import numpy as np
import pytest

from misti_b200 import io as mio

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DATA = os.path.join(ROOT, "data", "synthetic")


def test_read_psmc_plain_and_ancient(golden_datasets):
    units = mio.Units.from_file(os.path.join(DATA, "setunits.txt"))
    inp = mio.read_psmc(os.path.join(DATA, "m1.psmc"), os.path.join(DATA, "m2.psmc"), 0, -1, units)
    g = golden_datasets["synthetic"]
    assert inp.times == g["times"] and inp.lambdas == g["lambdas"]  # bit-identical: same arithmetic in the same order
    assert inp.sampleDateDiscr == g["sampleDate"] and inp.theta == g["theta"] and inp.rho == g["rho"]
    assert inp.scaleTime == g["scaleTime"] and len(inp.times) == len(inp.lambdas) - 1 == 126
    g = golden_datasets["synthetic_ancient"]
    units = mio.Units.from_file(os.path.join(DATA, "setunits.txt"), hetloss=g["hetloss"])
    inp = mio.read_psmc(os.path.join(DATA, "m1.psmc"), os.path.join(DATA, "m2.psmc"), g["sdate"], -1, units)
    assert inp.times == g["times"] and inp.lambdas == g["lambdas"] and inp.sampleDateDiscr == g["sampleDate"]
    assert inp.theta == g["theta"] and inp.rho == g["rho"]


def test_read_jafs_is_side_effect_free(golden_datasets):
    a = mio.read_jafs(os.path.join(DATA, "m.sfs"))
    b = mio.read_jafs(os.path.join(DATA, "m.sfs"))  # the reference would now hold 400 rows (mutable default list)
    assert len(a.jafs) == len(b.jafs) == 200 and a.pop1 == "A" and a.pop2 == "B"
    assert mio.column_sums(a.jafs) == golden_datasets["synthetic"]["sfs"]
    bs = mio.read_jafs(os.path.join(DATA, "bs.sfs"))
    assert bs.jafs[:6] == golden_datasets["synthetic"]["bs_rows"]


def test_bootstrap_semantics():
    rows = mio.read_jafs(os.path.join(DATA, "m.sfs")).jafs
    total = mio.column_sums(rows)
    out = mio.generate_bootstrap(rows, 5, seed=12345)
    assert out[0] == total and len(out) == 6
    for r in out[1:]:
        assert total[0] <= r[0] < total[0] + max(x[0] for x in rows)  # drawn until the genome length is reached
    assert out == mio.generate_bootstrap(rows, 5, seed=12345) and out != mio.generate_bootstrap(rows, 5, seed=1)
    # identical to the reference's procedure driven by the same generator
    rng = random.Random(12345)
    sfs = [0] * 8
    while sfs[0] < total[0]:
        pick = rows[rng.randint(0, len(rows) - 1)]
        sfs = [a + b for a, b in zip(sfs, pick)]
    assert sfs == out[1]


def test_units_file_and_hetloss(tmp_path):
    p = tmp_path / "u.txt"
    p.write_text("mutRate=2.5e-8\nbinsize=100\nN0=5000\ngenTime=25\njunk\n")
    u = mio.Units.from_file(str(p), hetloss=(0.05, None))
    assert (u.mutRate, u.N0, u.genTime, u.hetloss1, u.hetloss2) == (2.5e-8, 5000.0, 25.0, 0.05, 0.0)
    assert mio.Units.from_file(str(tmp_path / "missing.txt")).N0 == 10000


def test_mi_writer(tmp_path):
    class M:
        llh, splitT, sampleDate, thrh = -12.5, 1, 0, [0.05, 0.01]
        times, JAFS, dataJAFS = [0.1, 0.2], [0.1] * 7, [1.0] * 7
        lc, lh, mi = [[1, 2], [4, 4], [5, 5]], [[1, 1], [2, 2], [4, 4]], [[0.0, 0.5], [0, 0], [0, 0]]
        Pr = [[[1.0, 0.0], [0.0, 1.0], [0.0, 0.0]], [[0.9, 0.0], [0.0, 0.8], [0.0, 0.1]]]
    f = tmp_path / "o.mi"
    mio.output_migration(str(f), [], M, 20000, 1)
    lines = f.read_text().splitlines()
    assert lines[0] == "#MiSTI2 ver 0.4" and lines[1] == "LK\t-12.5" and lines[2] == "ST\t1"
    rs = [ln.split("\t") for ln in lines if ln.startswith("RS")]
    assert len(rs) == 3 and rs[0][1:8] == ["0", "1.0", "0.5", "1.0", "1.0", "0.0", "0.5"] and len(rs[0]) == 14 and len(rs[1]) == 8
    assert np.isclose(float(rs[2][1]), 0.3)


def test_read_ms_against_reference():
    """ms command line -> model (migrationIO.ReadMS, migrationIO.py:659-766), against outputs of the unmodified reference
    (tests/golden/ms.json, gen_ms_golden.py): README example, the function's own example, pulses, several bands per deme,
    deme 1 joining deme 2."""
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ms.json")) as f:
        cases = json.load(f)["cases"]
    assert len(cases) >= 5
    for c in cases:
        d = mio.read_ms(c["ms"])
        assert d.times == c["times"] and d.lambdas == c["lambdas"], c["ms"]
        assert d.divergenceTime == c["splitT"]
        assert [list(m) for m in d.mi] == c["mi"] and [list(p) for p in d.pu] == c["pu"], c["ms"]
    with pytest.raises(SystemExit):
        mio.read_ms("-n 1 2.0 -em 0.0 2 1 0.5")  # no -ej: the reference prints a notice and exits


def test_write_jafs_round_trip(tmp_path):
    """PrintJAFSFile's format (migrationIO.py:526-555) as written by write_jafs reads back unchanged; the bootstrap file of
    utils/generateJSFS_bs.py (row 0 = the data total, then the replicates) is generate_bootstrap + write_jafs."""
    rows = [[1000.0, 3.0, 1.0, 4.0, 1.0, 5.0, 9.0, 2.0], [2000.0, 6.0, 5.0, 3.0, 5.0, 8.0, 9.0, 7.0], [1500.0, 9.0, 3.0, 2.0, 3.0, 8.0, 4.0, 6.0]]
    bs = mio.generate_bootstrap(rows, 5, seed=3)
    fn = tmp_path / "bs.sfs"
    with open(fn, "w") as f:
        mio.write_jafs(bs, "A", "B\n", file=f)
    back = mio.read_jafs(str(fn))
    assert back.jafs == bs and back.pop1 == "A" and back.pop2 == "B"
    assert back.jafs[0] == mio.column_sums(rows)
    text = open(fn).read().splitlines()
    assert text[0] == "#MiSTI_JSFS version 1.0" and text[3].split("\t") == ["total", "0100", "1100", "0001", "0101", "1101", "0011", "0111"]
    with open(fn, "w") as f:
        mio.write_jafs([3.0, 1.0, 4.0, 1.0, 5.0, 9.0, 2.0], file=f)  # a single spectrum without its total
    assert mio.read_jafs(str(fn)).jafs == [[25.0, 3.0, 1.0, 4.0, 1.0, 5.0, 9.0, 2.0]]


def _with_pattern(tmp_path, pattern):
    names = []
    for n in ("m1.psmc", "m2.psmc"):
        with open(os.path.join(DATA, n)) as f:
            body = f.read()
        dst = tmp_path / n
        dst.write_text("MM\tpattern:%s, n:63, n_free_lambdas:%d\n" % (pattern, len(pattern.split("+"))) + body)
        names.append(str(dst))
    return names


def test_read_psmc1_equals_the_reference(tmp_path):
    """MiSTI.py -pm 1 (migrationIO.ReadPSMC1 over psmc.PSMC): grid, re-estimated sizes and split index as the reference
    produced them from the same files (tests/golden/psmc1.json, generator gen_psmc1_golden.py) -- bit for bit."""
    import json
    with open(os.path.join(ROOT, "tests", "golden", "psmc1.json")) as f:
        cases = json.load(f)["cases"]
    units = mio.Units.from_file(os.path.join(DATA, "setunits.txt"))
    for c in cases:
        f1, f2 = _with_pattern(tmp_path, c["pattern"])
        assert sum(mio.psmc_pattern(f1)) == 64
        inp = mio.read_psmc1(f1, f2, c["RD"], divergenceTime=c["st_years"], units=units)
        assert inp.times == c["times"] and inp.lambdas == c["lambdas"], c["name"]
        assert inp.divergenceTime == c["divTime"] and inp.scaleTime == c["scaleTime"] and inp.theta == c["theta"]
        assert inp.sampleDateDiscr == 0 and inp.rho is None and inp.Tpsmc is None


def test_read_jafs_format_version_0(tmp_path, capsys):
    """files of a format version < 1 (one spectrum as eight `label<TAB>count` lines, header fields separated by a blank) go
    through the old reader (migrationIO.ReadJAFS_old, migrationIO.py:610-656); the expectation below -- rows, population
    names and what is printed -- is what the reference's ReadJAFS returned for this very file in the build container."""
    counts = [2500000000, 910, 320, 905, 411, 380, 333, 290]
    labels = ["total", "0100", "1100", "0001", "0101", "1101", "0011", "0111"]
    fn = tmp_path / "old.jafs"
    fn.write_text("#MiSTI_JAF version 0.3\n#pop1 YRI\n#pop2 CEU\n" + "".join("%s\t%d\n" % lv for lv in zip(labels, counts)))
    j = mio.read_jafs(str(fn), silent_mode=False)
    assert j.jafs == [counts] and all(isinstance(v, int) for v in j.jafs[0]) and (j.pop1, j.pop2) == ("YRI", "CEU")
    assert capsys.readouterr().out == "JAFS format version: 0.3\npop1\t YRI\npop2\t CEU\n"
    fn.write_text("#MiSTI_JAF version 0.3\n" + "".join("%s\t%d\n" % lv for lv in zip(labels[:7], counts[:7])))
    with pytest.raises(SystemExit):
        mio.read_jafs(str(fn))
    fn.write_text("#MiSTI_JAF version 1.0\n")  # the older names only exist with versions < 1
    with pytest.raises(SystemExit):
        mio.read_jafs(str(fn))


def test_read_migration_equals_the_reference(tmp_path, capsys):
    """`.mi` files written by the reference (tests/golden/mi.json, generator gen_mi_golden.py) parse to exactly what the
    reference's ReadMigration made of them: scaled times, corrected and PSMC-apparent rates, likelihood, split, spectrum."""
    import json
    with open(os.path.join(ROOT, "tests", "golden", "mi.json")) as f:
        cases = json.load(f)["cases"]
    for c in cases:
        fn = tmp_path / (c["name"] + ".mi")
        fn.write_text(c["text"])
        d = mio.read_migration(str(fn))
        assert capsys.readouterr().out == "Format version:  0.4\n"
        for k in ("llh", "splitT", "sampleDate", "thrh", "jaf", "times", "lambda1", "lambda2", "lambdah1", "lambdah2"):
            assert getattr(d, k) == c[k], (c["name"], k)
        n = len(c["times"])
        assert len(d.mu1) == len(d.mu2) == n and all(len(p[0]) == len(p[1]) == n for p in (d.pr11, d.pr22, d.pr12))
        assert d.migStart is None and d.mi is None
        if c["name"] == "config2_cpfit":  # the band -mi 2 5 12 0.8: rates into population 2 on intervals 5..11
            assert [k for k, v in enumerate(d.mu2) if v != 0] == list(range(5, 12)) and not any(d.mu1)
            assert abs(d.pr11[0][0] - 1.0) < 1e-12 and d.pr11[0][c["splitT"]] == 0  # no Pr columns from the split on
