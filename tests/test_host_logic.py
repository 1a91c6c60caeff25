"""Host-side logic of the MigrationInference mirror that needs no device: constructor semantics
(fractional split surgery on the caller's lists), SetModel validation, MapParameters, SetJAFS constants."""
import numpy as np
import pytest

from misti_b200 import MigrationInference, llh_constants
from oracle.misti_oracle import OracleModel


def _grid(n=6):
    return [0.1 * (i + 1) for i in range(n - 1)], [[1.0 + i, 2.0 + i] for i in range(n)]


def test_fractional_split_mutates_caller_lists_like_the_reference():
    times, lam = _grid()
    M = MigrationInference(times, lam, [7] + [1] * 7, 2.25)
    assert M.splitT == 3 and M.numT == 7 and len(times) == 6 and len(lam) == 7  # the caller's lists grew
    assert times[2] == pytest.approx(0.25 * 0.3) and times[3] == pytest.approx(0.3 - 0.25 * 0.3)
    assert lam[3] == lam[2]
    assert M.times is times
    om = OracleModel(*_grid(), [7] + [1] * 7, 2.25)
    assert om.times == times and om.splitT == 3


def test_set_model_and_map_parameters():
    times, lam = _grid()
    M = MigrationInference(times, lam, [7] + [1] * 7, 4, [["1", "0", "2", "0.3", "1"], ["2", "1", "4", "0.8", "0"]],
                           [["2", "1", "0.05", "1"]], cpfit=True, unfolded=True)
    assert M.optMisSize == 1 and M.optPusSize == 1
    assert M.mi[0] == [0.3, 0.0] and M.mi[1] == [0.3, 0.8] and M.mi[3] == [0.0, 0.8] and M.mi[4] == [0.0, 0.0]
    assert M.pu[1] == [0.0, 0.05]
    M.MapParameters([1.5, 0.2])
    assert M.mi[0][0] == 1.5 and M.mi[1] == [1.5, 0.8] and M.pu[1][1] == 0.2
    assert M._current_params() == [1.5, 0.2]
    assert M._flags() == 1 | 2 | 8


@pytest.mark.parametrize("mi,pu", [
    ([["3", "0", "2", "0.3", "1"]], []),                                  # population index
    ([["1", "2", "2", "0.3", "1"]], []),                                  # empty band
    ([["1", "0", "3", "0.3", "1"], ["1", "2", "4", "0.1", "0"]], []),     # overlap
    ([], [["1", "1", "1.5", "1"]]),                                       # pulse rate > 1
    ([], [["1", "1", "0.1", "1"], ["2", "1", "0.1", "1"]]),               # two pulses in one interval
])
def test_invalid_models_exit_like_the_reference(mi, pu, capsys):
    times, lam = _grid()
    with pytest.raises(SystemExit) as e:
        MigrationInference(times, lam, [7] + [1] * 7, 4, mi, pu)
    assert e.value.code == 0  # the reference exits with status 0 (PrintError)
    assert "MigrationInference class error" in capsys.readouterr().out


def test_sample_date_rules(capsys):
    times, lam = _grid()
    with pytest.raises(SystemExit):
        MigrationInference(times, lam, [7] + [1] * 7, 1, sampleDate=2)
    with pytest.raises(SystemExit):
        MigrationInference(times, lam, [7] + [1] * 7, 4, [["1", "1", "3", "0.3", "1"]], sampleDate=2)


def test_likelihood_constants_match_the_oracle(golden_datasets):
    d = golden_datasets["synthetic"]
    for uf in (False, True):
        for row in [d["sfs"]] + d["bs_rows"]:
            om = OracleModel(d["times"], d["lambdas"], row, 40, unfolded=uf)
            assert llh_constants([row], uf)[0] == om.llh_const
    times, lam = _grid()
    M = MigrationInference(times, lam, d["sfs"], 3, unfolded=True)
    assert M.snps == sum(d["sfs"][1:]) and M.dataJAFS == d["sfs"][1:]
    with pytest.raises(SystemExit):
        M.SetJAFS([1] * 7)


def test_negative_parameter_short_circuits_without_a_device(capsys):
    times, lam = _grid()
    M = MigrationInference(times, lam, [7] + [1] * 7, 4, [["1", "0", "2", "0.3", "1"]])
    before = MigrationInference.COUNT_LLH
    assert M.JAFSLikelihood([-0.1]) == -np.inf  # MigrationInference.py:569-572: returns before any evaluation
    assert MigrationInference.COUNT_LLH == before + 1
    assert "Hit negative value of migration rate" in capsys.readouterr().out


def test_st_token_of_the_sweep_command_line(capsys):
    """`-mi 1 4 st 3 1` (test.bs/din_sar.bs.sh:29-38): the token stands for the split INDEX.  argparse delivers -st as a float;
    the substitution must write "40", which SetModel's int() accepts, and refuse a fractional split time."""
    import pytest
    from misti_b200.cli import _subst
    assert _subst([["1", "4", "st", "3", "1"]], 40.0) == [["1", "4", "40", "3", "1"]]
    assert _subst([["1", "4", "st", "3", "1"]], 37) == [["1", "4", "37", "3", "1"]]
    assert int(_subst([["1", "4", "st", "3", "1"]], 40.0)[0][2]) == 40
    assert _subst([["2", "5", "12", "0.8", "1"]], 40.5) == [["2", "5", "12", "0.8", "1"]]  # no token: nothing to substitute
    with pytest.raises(SystemExit):
        _subst([["1", "4", "st", "3", "1"]], 40.5)
    assert "whole split time" in capsys.readouterr().out


def test_split_time_interval_is_the_notebooks_reduction():
    """misti_b200.sweep.split_time_interval against the formulas of test.bs/bs_conf_int.ipynb (conf_int_bs): per replicate the
    split time of the highest likelihood, Student-t interval of their mean, histogram; row 0 (the data) kept apart."""
    from collections import Counter
    import scipy.stats as st
    from misti_b200.sweep import split_time_interval
    rng = np.random.default_rng(8)
    sts = np.arange(36, 45)
    llh = -1e4 - (sts[None, :] - 40 - rng.normal(0, 1.2, (101, 1))) ** 2 + rng.normal(0, 0.1, (101, 9))
    best = sts[np.argmax(llh, axis=1)]
    got = split_time_interval(best)
    a = [float(v) for v in best[1:]]  # the notebook: bs_mas[1:], "zero is without bootstrap"
    want = st.t.interval(0.975, len(a) - 1, loc=np.mean(a), scale=st.sem(a))
    assert got["interval"] == (float(want[0]), float(want[1])) and got["from_data"] == float(best[0])
    assert got["histogram"] == dict(Counter(a))
