"""Regenerates tests/golden/*.npz from the CPU oracle (run from the repo root).
The reference ships no golden vectors and cannot be executed here, so these fixtures
pin the ORACLE against accidental change; they are not outputs of the reference."""
import sys
sys.path.insert(0, ".")
import numpy as np
from oracle import cpu_oracle as O
from spectrobot_b200 import synthetic as S

g = S.spectral_grid(2999.0, 3001.0)
lines = S.line_table(80, 2996.0, 3004.0, n_levels=4, seed=12)
cell = O.gcoeff_cell(lines, g, 150.0, 0.1, S.CH4_MM, 4)
np.savez_compressed("tests/golden/cell_small.npz", cell=cell[:, :, ::40])
x = 3000.0 + (np.arange(13010) - 6505) * 5e-4
np.savez_compressed("tests/golden/humliv_small.npz",
                    y=O.humliv_bb(x, 1, 13010, 3000.00013, 2e-3, 4e-3)[::10])
