#!/usr/bin/env python3
"""Consumer of julia/compare_with_oceananigans.jl: diff a dump of the REAL reference (Oceananigans v0.90.1, CPU) against
the oracle's frozen reference-arithmetic mode, its product-arithmetic mode and -- when a GPU is present -- libocean_b200.so,
on the case `ref_test` (tests/cases.py), after the number of steps stored in the dump.

    python tests/compare_dump.py ref_test_oceananigans.npz [--tol 1e-10]

Prints the relative L-infinity distance per prognostic field and exits 1 if any exceeds `--tol` for the reference-
arithmetic oracle (the pin DESIGN.md section 2 is waiting for).  tests/test_oracle_cpu.py runs it on a self-made dump
(oracle written through the same .npz layout) so that the tool itself is tested without Julia.
"""
import argparse
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

PROG = ("U", "V", "W", "ETA", "T", "S")


def run_case(pkg, factory, steps, dt, reference_arithmetic=False):
    import cases
    c = cases.build_case("ref_test")
    m = c.make_model(pkg, factory, initialize=False)
    if reference_arithmetic:
        m.backend.set_option("reference_arithmetic", 1)
    c.initialize(pkg, m)
    for _ in range(steps):
        pkg.time_step(m, dt)
    return {n: m.get(n) for n in PROG}


def distances(ours, dump):
    out = {}
    for n in PROG:
        ref = np.asarray(dump[n], dtype=np.float64).reshape(ours[n].shape)
        out[n] = float(np.max(np.abs(ours[n] - ref))) / max(float(np.max(np.abs(ref))), 1e-300)
    return out


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("dump")
    ap.add_argument("--tol", type=float, default=1e-10)
    args = ap.parse_args(argv)
    import __graft_entry__ as ge
    pkg, orc = ge.package(), ge.oracle_module()
    orc.build()
    dump = np.load(args.dump)
    steps, dt = int(dump["steps"]), float(dump["dt"])
    rows = [("oracle, reference arithmetic", run_case(pkg, orc.OracleBackend, steps, dt, True)),
            ("oracle, product arithmetic", run_case(pkg, orc.OracleBackend, steps, dt))]
    try:
        import torch
        if torch.cuda.is_available():
            rows.append(("libocean_b200.so (GPU)", run_case(pkg, None, steps, dt)))
    except Exception:
        pass
    worst = 0.0
    print(f"{args.dump}: {steps} steps of ref_test (72 x 30 x 10, dt = {dt:g} s); relative L-infinity vs the dump")
    for label, fields in rows:
        d = distances(fields, dump)
        print(f"  {label:32s} " + "  ".join(f"{n} {d[n]:.2e}" for n in PROG))
        if label.startswith("oracle, reference"):
            worst = max(d.values())
    print(f"  -> {'WITHIN' if worst <= args.tol else 'ABOVE'} {args.tol:g} (reference-arithmetic oracle vs dump: {worst:.2e})")
    return 0 if worst <= args.tol else 1


if __name__ == "__main__":
    sys.exit(main())
