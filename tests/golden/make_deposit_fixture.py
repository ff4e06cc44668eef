#!/usr/bin/env python
"""Writes tests/golden/deposit_xyz.npz: the 736 sphere centres of the reference's own deposit deck
(/root/reference/example/spheres/deposit/input_deposit_xyz.xyz, read by example/spheres/deposit/add_xyz.msp) as a float64 array.
The GPU box has no /root/reference, so the coordinates travel as this small fixture; `workloads.deposit_xyz` tiles them into
S1 of SURVEY 8(d) (BASELINE configs[0]).

    python tests/golden/make_deposit_fixture.py
"""
from pathlib import Path

import numpy as np

SRC = Path("/root/reference/example/spheres/deposit/input_deposit_xyz.xyz")

if __name__ == "__main__":
    lines = SRC.read_text().split("\n")
    n = int(lines[0])
    box = np.array([float(x) for x in lines[1].split()])
    xyz = np.array([[float(x) for x in ln.split()[1:4]] for ln in lines[2:2 + n]])
    assert xyz.shape == (n, 3) and (xyz >= 0).all() and (xyz <= box).all()
    np.savez_compressed(Path(__file__).resolve().parent / "deposit_xyz.npz", xyz=xyz, box=box)
    print("wrote", n, "centres, box", box)
