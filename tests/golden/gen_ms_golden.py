#!/usr/bin/env python3
"""Golden vectors for migrationIO.ReadMS (migrationIO.py:659-766): ms command line -> times, rates, split index, -mi / -pu
layouts.  Run in the build container only (needs /root/reference through ref_shim); writes ms.json next to this script."""
import contextlib
import io
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

STRINGS = [
    # README.md:102
    "4 100 -t 15000 -r 1920 30000000 -l -I 2 2 2 -n 1 10 -n 2 4.5 -eN 0.025 0.2 -ej 0.045 2 1 -eN 0.175 3 -eN 0.625 1.8 -eN 3 3.2 -eN 8 5.5",
    # the example inside migrationIO.ReadMS (:661)
    "-n 2 3.0 -em 0.0 1 2 2.0 -em 0.05 2 1 3.0 -en 0.01 1 0.5 -en 0.02 2 0.05 -en 0.0375 1 0.5 -en 0.0375 2 0.5 -ej 1.25 2 1 -eM 1.25 0.0 -eN 1.25 1.0 -eN 2.0 5.0",
    # pulses (-es), several bands into the same deme, population 1 joining population 2
    "4 1 -I 2 2 2 -n 1 1.5 -n 2 0.7 -es 0.01 1 0.9 -em 0.02 1 2 4.0 -em 0.06 1 2 1.0 -em 0.03 2 1 2.5 -en 0.04 2 0.3 -ej 0.2 1 2 -eN 0.5 2.0",
    "4 1 -I 2 2 2 -es 0.005 2 0.95 -es 0.03 1 0.8 -en 0.01 1 2.0 -ej 0.1 2 1 -eN 0.1 1.2 -eN 1.0 0.4",
    "-n 1 2.0 -em 0.0 2 1 0.5 -ej 0.3 2 1",
]


def main():
    R = ref_shim.load()
    mio = R["migrationIO"]
    cases = []
    for s in STRINGS:
        with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
            d = mio.ReadMS(s)
        cases.append({"ms": s, "times": [float(v) for v in d.times], "lambdas": [[float(v[0]), float(v[1])] for v in d.lambdas],
                      "splitT": int(d.divergenceTime), "mi": [[int(m[0]), int(m[1]), int(m[2]), float(m[3]), int(m[4])] for m in d.mi],
                      "pu": [[int(p[0]), int(p[1]), float(p[2]), int(p[3])] for p in d.pu]})
    with open(os.path.join(HERE, "ms.json"), "w") as f:
        json.dump({"meta": {"generator": "tests/golden/gen_ms_golden.py"}, "cases": cases}, f, indent=1)
    print("wrote", len(cases))


if __name__ == "__main__":
    main()
