"""Run HERE (container with /root/reference): converts the reference's shipped example system
examples/matrix.mtx (12x12, 61 nnz, values 1..61) into a small fixture so that the GPU box, which has
no /root/reference, can run config 1.  Output: tests/golden/example_matrix_12x12.npz"""
from pathlib import Path

import numpy as np

src = Path("/root/reference/examples/matrix.mtx")
rows, cols, vals = [], [], []
n = None
for line in src.read_text().splitlines():
    if line.startswith("%") or not line.strip():
        continue
    t = line.split()
    if n is None:
        n = int(t[0])
        continue
    rows.append(int(t[0]) - 1)
    cols.append(int(t[1]) - 1)
    vals.append(float(t[2]))
rows, cols, vals = np.array(rows), np.array(cols), np.array(vals)
order = np.lexsort((cols, rows))
rows, cols, vals = rows[order], cols[order], vals[order]
rp = np.zeros(n + 1, np.int32)
np.add.at(rp, rows + 1, 1)
rp = np.cumsum(rp).astype(np.int32)
out = Path(__file__).with_name("example_matrix_12x12.npz")
np.savez_compressed(out, row_ptr=rp, col_idx=cols.astype(np.int32), values=vals)
print(out, n, len(vals))
