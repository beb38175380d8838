"""The lane → (column group, row) mapping of the kernels that pack several rows into a warp
(pointwise_kernel, spmm_fused_kernel, spmm_f32_kernel's ragged last tile — csrc/spmm.cu),
restated in Python and checked exhaustively: every (row, column group) of a CTA is visited
exactly once, and the cp.async ring of pointwise_kernel always finds the row it waits for at its
head.  Runs on the CPU; the bitwise GPU checks are tests/test_gpu_pointwise_shapes.py and
tests/test_gpu_spmm_ragged_tiles.py.
"""

from __future__ import annotations

K_WARPS, K_WARP, K_PW_ROWS, K_AHEAD = 8, 32, 4, 4


def vbits_of(n_vec: int, start: int) -> int:
    vb = start
    while vb > 0 and (1 << (vb - 1)) >= n_vec:
        vb -= 1
    return vb


def test_pointwise_kernel_visits_every_row_and_column_group_once():
    for t_here in range(1, 9):  # tile-table entries of the CTA: spare warps share the rows
        groups = K_WARPS // t_here
        for n_vec in range(1, 33):
            vb = vbits_of(n_vec, 5)
            rpw = K_WARP >> vb
            for nrows in range(1, 33):
                want = {(r, v) for r in range(nrows) for v in range(n_vec)}
                for path in ("copy", "ring"):
                    seen: dict[tuple[int, int], int] = {}
                    for group in range(groups):
                        for lane in range(K_WARP):
                            v, sub = lane & ((1 << vb) - 1), lane >> vb
                            if v >= n_vec:
                                continue
                            rfirst, rstep = group * rpw + sub, groups * rpw
                            if path == "copy":
                                for lr in range(rfirst, nrows, rstep * K_PW_ROWS):
                                    for j in range(K_PW_ROWS):
                                        if lr + j * rstep < nrows:
                                            seen[(lr + j * rstep, v)] = seen.get((lr + j * rstep, v), 0) + 1
                            else:
                                ring = [rfirst + j * rstep for j in range(K_AHEAD) if rfirst + j * rstep < nrows]
                                for lr in range(rfirst, nrows, rstep):
                                    assert ring and ring.pop(0) == lr
                                    seen[(lr, v)] = seen.get((lr, v), 0) + 1
                                    if lr + K_AHEAD * rstep < nrows:
                                        ring.append(lr + K_AHEAD * rstep)
                                assert not ring
                    assert set(seen) == want and set(seen.values()) == {1}, (t_here, n_vec, nrows, path)


def test_spmm_packed_tiles_visit_every_row_and_column_once():
    for start, widest in ((5, 32), (4, 16)):  # spmm_fused_kernel (any tile), spmm_f32_kernel (ragged last tile <= 16)
        for n_vec in range(1, widest + 1):
            vb = vbits_of(n_vec, start)
            rpw = K_WARP >> vb
            for nrows in range(1, 65):
                seen: dict[tuple[int, int], int] = {}
                for warp in range(K_WARPS):
                    for lane in range(K_WARP):
                        v = lane & ((1 << vb) - 1)
                        if v >= n_vec:
                            continue
                        for lr in range(warp * rpw + (lane >> vb), nrows, K_WARPS * rpw):
                            seen[(lr, v)] = seen.get((lr, v), 0) + 1
                assert set(seen) == {(r, v) for r in range(nrows) for v in range(n_vec)} and set(seen.values()) == {1}
