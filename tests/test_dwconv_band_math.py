"""CPU: the index algebra of the tensor-pipe depthwise 7x7 kernels (linnaeus_b200/csrc/lnx_dwconv_mma.cu), emulated lane by lane in
numpy with the fragment ownership of `mma.sync.m16n8k16` (PTX ISA: A row-major 16x16, B column-major 16x8, C/D 16x8), against a
direct cross-correlation.  It pins, without a GPU, the three choices the kernels rest on:

* forward / data gradient: MMA rows g, g+8 <-> tile rows 2g, 2g+1; K permutation k = 2t, 2t+1, 2t+8, 2t+9 <-> window columns
  t, t+4, t+8, t+12; band entries b0 = {w[t-g], w[t+4-g]}, b1 = {w[t+8-g], w[t+12-g]} (zero outside 0..6);
* weight gradient: A[kx][x] = in_h[yi][x+kx] (rows 8..15: input row yi+1), B[x][n] = dy[yi-n][x] (n = 7: dy[yi+1][x]);
  D[kx][n] -> dW[n][kx], D[8+kx][n] -> dW[n+1][kx] (n <= 5), D[8+kx][7] -> dW[0][kx]; row 7 of A = ones, D[7][2] + D[7][3] = the bias gradient
  when the padded input rows are stepped over the image rows only;
* the shared-memory swizzle: the 8 lanes of every phase of a 128-bit fragment load hit 8 distinct 16-byte bank groups.

The GPU parity of the kernels themselves is tests/test_gpu_dwconv_mma.py.  R/models/blocks/convnext.py:56-58."""
import numpy as np

LANES = [(lane >> 2, lane & 3) for lane in range(32)]  # (g, t)


def xo(k):
    """window column of MMA k index: k = 2t, 2t+1, 2t+8, 2t+9 -> t, t+4, t+8, t+12"""
    t, r = (k % 8) // 2, k % 2
    return t + 4 * r + 8 * (k // 8)


def mma_from_fragments(a, b):
    """a[lane] = (a0, a1, a2, a3), b[lane] = (b0, b1), each register a pair (lo, hi) -> D[16, 8] as the hardware computes it."""
    A = np.zeros((16, 16))
    B = np.zeros((16, 8))
    for (g, t), ar, br in zip(LANES, a, b):
        for reg, (row, kb) in enumerate(((g, 0), (g + 8, 0), (g, 8), (g + 8, 8))):
            A[row, kb + 2 * t], A[row, kb + 2 * t + 1] = ar[reg]
        for reg, kb in enumerate((0, 8)):
            B[kb + 2 * t, g], B[kb + 2 * t + 1, g] = br[reg]
    return A @ B


def test_k_permutation_is_a_bijection():
    assert sorted(xo(k) for k in range(16)) == list(range(16))
    for t in range(4):
        assert [xo(2 * t), xo(2 * t + 1), xo(2 * t + 8), xo(2 * t + 9)] == [t, t + 4, t + 8, t + 12]


def test_forward_tile_equals_direct_correlation():
    rng = np.random.default_rng(0)
    rows, cols = 16, 24  # one 16-row tile, three 8-column blocks
    x = rng.standard_normal((rows + 6, cols + 6 + 8))  # padded ("halo") input of one channel; slack columns meet zero band entries
    x[:, cols + 6:] = rng.standard_normal((rows + 6, 8)) * 100.0  # ... whatever they hold
    w = rng.standard_normal((7, 7))
    bias = 0.3
    wp = lambda ky, i: w[ky, i] if 0 <= i <= 6 else 0.0
    out = np.full((rows, cols), np.nan)
    for j in range(cols // 8):
        D = np.full((16, 8), bias)
        for ky in range(7):
            a, b = [], []
            for g, t in LANES:
                lo = lambda r: (x[2 * g + r, 8 * j + t], x[2 * g + r, 8 * j + t + 4])
                hi = lambda r: (x[2 * g + r, 8 * j + t + 8], x[2 * g + r, 8 * j + t + 12])
                a.append((lo(ky), lo(ky + 1), hi(ky), hi(ky + 1)))
                d = t - g
                b.append(((wp(ky, d), wp(ky, d + 4)), (wp(ky, d + 8), wp(ky, d + 12))))
            D = D + mma_from_fragments(a, b)
        for g, t in LANES:  # c0, c1 = D[g][2t], D[g][2t+1] -> tile row 2g; c2, c3 = D[g+8][...] -> tile row 2g+1
            for hh in range(2):
                for xx in range(2):
                    out[2 * g + hh, 8 * j + 2 * t + xx] = D[g + 8 * hh, 2 * t + xx]
    ref = np.array([[bias + (x[r:r + 7, c:c + 7] * w).sum() for c in range(cols)] for r in range(rows)])
    assert np.allclose(out, ref, rtol=1e-12, atol=1e-10)


def test_weight_gradient_band_equals_direct_sum():
    rng = np.random.default_rng(1)
    H, W = 12, 32  # image; bands of 8 input rows stepped over the image rows only; two 16-column steps
    xin = rng.standard_normal((H, W))
    dy = rng.standard_normal((H, W))
    # padded input in_h[yh][xh] = xin[yh-3][xh-3]; dy rows / columns outside the image are zero
    in_h = np.zeros((H + 6 + 16, W + 6 + 32))
    in_h[3:3 + H, 3:3 + W] = xin
    in_h[3 + H + 3:, :] = 0.0
    dyz = lambda y, xx: dy[y, xx] if 0 <= y < H and 0 <= xx < W else 0.0
    dW = np.zeros((7, 7))
    db = 0.0
    for band in range((H + 7) // 8):
        for rp in range(4):
            yi = 3 + 8 * band + 2 * rp  # padded row index of the lower input row of this pair
            for xc in range(W // 16):
                a, b = [], []
                for g, t in LANES:
                    col = 16 * xc + t + g
                    row = lambda r: [(in_h[r, col + 8 * h], in_h[r, col + 8 * h + 4]) for h in range(2)]
                    r0, r1 = row(yi), row(yi + 1)
                    if g == 7:
                        r0 = [(1.0, 1.0), (1.0, 1.0)]  # row m = 7 of A: ones
                    a.append((r0[0], r1[0], r0[1], r1[1]))
                    ydy = yi - g if g < 7 else yi + 1
                    bcol = lambda h: (dyz(ydy, 16 * xc + t + 8 * h), dyz(ydy, 16 * xc + t + 8 * h + 4))
                    b.append((bcol(0), bcol(1)))
                D = mma_from_fragments(a, b)
                for kx in range(7):
                    for n in range(7):
                        dW[n, kx] += D[kx, n]
                    for n in range(6):
                        dW[n + 1, kx] += D[8 + kx, n]
                    dW[0, kx] += D[8 + kx, 7]
                db += D[7, 2] + D[7, 3]
    # dy is indexed by OUTPUT row y with in_h row yh = y + ky: dy[yi - n] above uses padded-row arithmetic (yi - n is an output row)
    ref = np.zeros((7, 7))
    for ky in range(7):
        for kx in range(7):
            ref[ky, kx] = sum(dy[y, xx] * in_h[y + ky, xx + kx] for y in range(H) for xx in range(W))
    assert np.allclose(dW, ref, rtol=1e-11, atol=1e-9)
    assert np.isclose(db, dy.sum(), rtol=1e-11)


def _swz(p, cq):
    return p * 64 + ((cq ^ ((p >> 1) & 3)) << 4)


def test_fragment_loads_are_bank_conflict_free():
    """A 128-bit shared load is served in four phases of 8 lanes; a phase is conflict free when its 8 addresses fall into 8 distinct
    16-byte groups of the 128-byte bank row (or coincide).  Forward: tile row pitch RP = 2 mod 4; weight gradient: any input pitch,
    dy pitch RPG = 4 mod 8."""
    def phases_ok(pixels, cq):
        for ph in range(4):
            addrs = {_swz(pixels[l], cq) for l in range(8 * ph, 8 * ph + 8)}
            if len({(a // 16) % 8 for a in addrs}) != len(addrs):
                return False
        return True

    for RP in (10, 22, 34, 38, 62):  # forward halo tiles
        for j in range(3):
            for r in range(8):
                for i in range(4):
                    px = [(2 * g + r) * RP + 8 * j + t + 4 * i for g, t in LANES]
                    assert all(phases_ok(px, cq) for cq in range(4)), (RP, j, r, i)
    for RP in (20, 34, 62, 63):  # weight gradient, input band (lanes read pixel t + g: neighbours coincide)
        for rp in range(4):
            for i in range(4):
                px = [2 * rp * RP + t + g + 4 * i for g, t in LANES]
                assert all(phases_ok(px, cq) for cq in range(4))
    for RPG in (20, 36, 68):  # weight gradient, dy band
        for rp in range(4):
            for i in range(4):
                px = [((2 * rp + 6 - g) if g < 7 else (2 * rp + 7)) * RPG + t + 4 * i for g, t in LANES]
                assert all(phases_ok(px, cq) for cq in range(4))
