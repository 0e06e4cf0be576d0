"""CPU spec of the index algebra behind the float64 tensor-core sweeps (csrc/gpmpc_uniform_impl.cuh, uni_*_mma8):
an emulation of `mma.sync.aligned.m8n8k4.row.col.f64` on per-lane fragments, driven exactly as the kernels drive it --
which lane holds which row / column / state dimension, that a lane's own accumulator values can serve as the A
fragment of the next product when the B fragments are picked by column parity, and where the shuffle reductions leave
their results.  (The kernels themselves are checked against the oracle on the GPU; this test pins the layout reasoning.)"""
import numpy as np


def dmma(d, a, b):
    """Warp-wide D (8x8) += A (8x4) B (4x8).  Fragments per lane (g = lane >> 2, q = lane & 3):
    a[lane] = A[g][q], b[lane] = B[q][g], d[lane] = (D[g][2q], D[g][2q+1])."""
    A = a.reshape(8, 4)                      # [g][q]
    B = b.reshape(8, 4).T                    # b[lane = 4 n + k] = B[k][n]  ->  B[k][n]
    D = d.reshape(8, 8) + A @ B              # d[lane = 4 g + q] = D[g][2q : 2q + 2]
    return D.reshape(32, 2)


def lanes():
    lane = np.arange(32)
    return lane, lane >> 2, lane & 3


def test_exponent_tile_is_one_dmma_per_k_step():
    rng = np.random.default_rng(0)
    E = 8
    u = rng.standard_normal((32, E))         # rows ib .. ib + 31
    nu = rng.standard_normal((8, E))         # columns j0 .. j0 + 7
    kap = rng.standard_normal(8)
    lane, g, q = lanes()
    for m in range(4):                       # row group m: rows 8 m + g
        d = np.stack([kap[2 * q], kap[2 * q + 1]], axis=1)             # accumulator starts at kap of the lane's columns
        for ks in range(2):
            a = u[8 * m + g, 4 * ks + q]                               # A fragment: the lane's row, state dimension 4 ks + q
            b = nu[g, 4 * ks + q]                                      # B fragment: column j0 + g, same state dimension
            d = dmma(d, a, b)
        want = kap[None, :] + u[8 * m:8 * m + 8] @ nu.T                # t[row][col]
        np.testing.assert_allclose(d.reshape(8, 8), want, rtol=1e-13, atol=1e-13)


def test_row_sums_use_the_lanes_own_values_as_a_fragments():
    """r_b,i += sum_j Eh_ij beta_b,j over the 8 columns of a round: two DMMAs whose A fragments are the accumulator
    entries the lane already holds (columns 2q and 2q+1); the k slot q of the first stands for column 2q, of the second
    for column 2q+1, and the B fragments are beta_{b = g, that column}."""
    rng = np.random.default_rng(1)
    E = 8
    Eh = rng.standard_normal((32, 8))
    beta = rng.standard_normal((8, E))       # [column][b]
    lane, g, q = lanes()
    for m in range(4):
        ex = Eh[8 * m + g][np.arange(32)[:, None], np.stack([2 * q, 2 * q + 1], axis=1)]   # the lane's two values
        r = np.zeros((32, 2))
        r = dmma(r, ex[:, 0], beta[2 * q, g])          # B[k = q][n = g] = beta_{b = g, column 2 q}
        r = dmma(r, ex[:, 1], beta[2 * q + 1, g])      #                 = beta_{b = g, column 2 q + 1}
        want = Eh[8 * m:8 * m + 8] @ beta              # [row g][b]
        np.testing.assert_allclose(r.reshape(8, 8), want, rtol=1e-13, atol=1e-13)   # lane (g, q) holds b = 2q, 2q+1 of row g
    # state dimensions 6 and 7 run the same code: the unused output columns (b >= E) collect garbage that is never read
    for E in (6, 7):
        beta_pad = np.concatenate([rng.standard_normal((8, E)), 1e3 * rng.standard_normal((8, 8 - E))], axis=1)
        ex = Eh[g][np.arange(32)[:, None], np.stack([2 * q, 2 * q + 1], axis=1)]
        r = dmma(dmma(np.zeros((32, 2)), ex[:, 0], beta_pad[2 * q, g]), ex[:, 1], beta_pad[2 * q + 1, g])
        np.testing.assert_allclose(r.reshape(8, 8)[:, :E], Eh[:8] @ beta_pad[:, :E], rtol=1e-13, atol=1e-13)


def shfl_xor(v, mask):
    return v[np.arange(32) ^ mask]


def test_column_sums_end_in_lanes_0_3_and_16_19():
    rng = np.random.default_rng(2)
    lane, g, q = lanes()
    tile = rng.standard_normal((32, 8))                            # w of one round: 32 rows x 8 columns
    # per lane: v0, v1 = its two columns summed over its 4 row groups (rows 8 m + g)
    v0 = sum(tile[8 * m + g, 2 * q] for m in range(4))
    v1 = sum(tile[8 * m + g, 2 * q + 1] for m in range(4))
    up = (lane & 16) != 0
    a = np.where(up, v1, v0) + shfl_xor(np.where(up, v0, v1), 16)  # lanes 0-15 keep column 2q, lanes 16-31 column 2q+1
    a = a + shfl_xor(a, 8)
    a = a + shfl_xor(a, 4)
    writers = (lane & 12) == 0
    cols = 2 * q + (lane >> 4)
    got = np.zeros(8)
    got[cols[writers]] = a[writers]
    np.testing.assert_allclose(got, tile.sum(axis=0), rtol=1e-13, atol=1e-13)
    assert sorted(cols[writers]) == list(range(8))


def test_reduce_over_row_groups_leaves_index_2g_in_lane_g():
    """uni_reduce_over_g16: 16 values per lane (index 2 a + k), summed over the 8 row groups g (lanes with equal q) with
    halving exchanges at 16, 8, 4; lane (g, q) ends with the totals of index 2 g and 2 g + 1 for its q."""
    rng = np.random.default_rng(3)
    v = rng.standard_normal((32, 16))
    lane, g, q = lanes()
    u16, u8, u4 = (lane & 16) != 0, (lane & 8) != 0, (lane & 4) != 0
    a = np.stack([np.where(u16, v[:, k + 8], v[:, k]) + shfl_xor(np.where(u16, v[:, k], v[:, k + 8]), 16) for k in range(8)], axis=1)
    b = np.stack([np.where(u8, a[:, k + 4], a[:, k]) + shfl_xor(np.where(u8, a[:, k], a[:, k + 4]), 8) for k in range(4)], axis=1)
    o0 = np.where(u4, b[:, 2], b[:, 0]) + shfl_xor(np.where(u4, b[:, 0], b[:, 2]), 4)
    o1 = np.where(u4, b[:, 3], b[:, 1]) + shfl_xor(np.where(u4, b[:, 1], b[:, 3]), 4)
    for qq in range(4):
        tot = v[q == qq].sum(axis=0)                                # over the 8 lanes that share q
        sel = q == qq
        np.testing.assert_allclose(o0[sel], tot[2 * g[sel]], rtol=1e-13, atol=1e-13)
        np.testing.assert_allclose(o1[sel], tot[2 * g[sel] + 1], rtol=1e-13, atol=1e-13)


def test_half_weight_diagonal_tiles_of_32_rows():
    """Symmetric w: sweeping the tiles on or above the diagonal, the diagonal tile in full with half weights, gives every
    consumer of (rho_i + gam_i) what the full sweep gives -- with 32-row tiles as with 64-row ones."""
    rng = np.random.default_rng(4)
    n = 96
    w = rng.standard_normal((n, n)); w = w + w.T
    full = w.sum(axis=1)                                            # row sums of the full matrix
    for T in (32, 64):
        if n % T:
            continue
        g = np.zeros(n)
        for I in range(n // T):
            rows = slice(T * I, T * I + T)
            g[rows] += 0.5 * w[rows, rows].sum(axis=1)              # rho of the diagonal tile, half weight
            g[rows] += 0.5 * w[rows, rows].sum(axis=0)              # gam of the diagonal tile, half weight
            for J in range(I + 1, n // T):
                cols = slice(T * J, T * J + T)
                g[rows] += w[rows, cols].sum(axis=1)                # rho
                g[cols] += w[rows, cols].sum(axis=0)                # gam: the mirror tile's row sums
        np.testing.assert_allclose(g, full, rtol=1e-12, atol=1e-12)
