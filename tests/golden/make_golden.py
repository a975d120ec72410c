#!/usr/bin/env python3
"""Generate the golden vectors tests/golden/*.npz from the CPU oracle.

There is no reference-produced fixture for this path (no Julia here; the reference's tests pin no numbers),
so these vectors pin the ORACLE (regression) -- they are not Oceananigans outputs.  Stored as float64
signatures (three horizontal slices per field + sum, sum of squares, max-abs of the full field) after
`steps` time steps of the closed-form cases in tests/cases.py, to keep the fixtures small.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import __graft_entry__ as ge  # noqa: E402
import cases  # noqa: E402

STEPS = {"ref_test": 10, "dd_small": 20, "bumpy": 20}   # ref_test: 10 iterations as test/runtests.jl:15

if __name__ == "__main__":
    pkg, orc = ge.package(), ge.oracle_module()
    orc.build()
    for name, steps in STEPS.items():
        c = cases.build_case(name)
        m = c.make_model(pkg, orc.OracleBackend)
        for _ in range(steps):
            pkg.time_step(m, c.dt)
        out = cases.signature({n: m.get(n) for n in ("U", "V", "W", "ETA", "T", "S")})
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), steps=steps, dt=c.dt, **out)
        print(name, {k: float(v[2]) for k, v in out.items() if k.endswith("_stats")})
