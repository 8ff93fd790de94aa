"""CPU emulation of the row schedule of k_iterate_t2 (csrc/tvl1_kernels.cuh): two iterations per launch,
the second following the first one row behind, three rotating row blocks per warp strip, 120 owned columns + one
halo group of four either side.  The per-pixel arithmetic is replaced by simple stand-ins with the SAME data
dependencies and boundary rules as src/tvl1flow.cpp:114-181 / src/operators.cpp:35-125:

    u'(x,y) = F(u, p11(x,y) - p11(x-1,y), p12(x,y) - p12(x,y-1))      p[-1] = 0; "+p" dropped on the last column / row
    p'(x,y) = G(p, u'(x+1,y) - u'(x,y), u'(x,y+1) - u'(x,y))           differences are 0 on the last column / row

so that the emulation checks what the kernel's control flow has to get right: which rows and columns of which
iteration are available when, for every strip position (first / last strips, strips that end inside the image, images
narrower or shorter than a strip), and that assembling the strips' stores gives exactly two global iterations."""
import numpy as np
import pytest

W_OWN, GROUP = 120, 4          # kT2W, pixels per lane
LANES = 32


def primal_rows(u, p11, p12, p11_left, p12_above, last_col, last_row):
    """u' of one row segment (vectors over x).  p11_left / p12_above: the neighbours' values (0 where the rule says so)."""
    a = np.where(last_col, 0.0, p11) - p11_left
    b = (0.0 if last_row else p12) - p12_above
    return u + 0.25 * a + 0.5 * b


def dual_rows(p11, p12, uc, u_right, u_below, last_col, has_below):
    dx = np.where(last_col, 0.0, u_right - uc)
    dy = (u_below - uc) if has_below else np.zeros_like(uc)
    return 0.75 * p11 + 0.125 * dx, 0.5 * p12 + 0.25 * dy


def reference_iteration(u, p11, p12):
    ny, nx = u.shape
    un = np.empty_like(u)
    for y in range(ny):
        left = np.concatenate([[0.0], p11[y, :-1]])
        above = p12[y - 1] if y > 0 else np.zeros(nx)
        last_col = np.arange(nx) >= nx - 1
        un[y] = primal_rows(u[y], p11[y], p12[y], left, above, last_col, y == ny - 1)
    q11, q12 = np.empty_like(p11), np.empty_like(p12)
    for y in range(ny):
        right = np.concatenate([un[y, 1:], [0.0]])
        below = un[y + 1] if y + 1 < ny else np.zeros(nx)
        last_col = np.arange(nx) >= nx - 1
        q11[y], q12[y] = dual_rows(p11[y], p12[y], un[y], right, below, last_col, y + 1 < ny)
    return un, q11, q12


class Block:
    """One register block of a warp: a row segment of u, p11, p12 over the warp's 128 columns (32 lanes x 4)."""
    def __init__(self, n):
        self.u = np.zeros(n); self.p11 = np.zeros(n); self.p12 = np.zeros(n)


def emulate_strip(u, p11, p12, out, bx, ys, R):
    """iterate_t2_pair for the warp strip (column group bx, rows ys .. ys+R): the kernel's stage order, literally."""
    ny, nx = u.shape
    ye = min(ys + R, ny)
    x0 = bx * W_OWN - GROUP                       # lane 0's first column
    xs = x0 + np.arange(LANES * GROUP)            # columns held by the warp
    in_img = (xs >= 0) & (xs < nx)                # (the kernel's in_alloc, with pitch == nx here)
    own = (xs >= x0 + GROUP) & (xs < x0 + GROUP + W_OWN) & (xs < nx)
    last_col = xs >= nx - 1
    outside_left = xs < 0

    def load_row(y, blk):
        for name, src in (("u", u), ("p11", p11), ("p12", p12)):
            v = np.zeros(LANES * GROUP)
            v[in_img] = src[y, xs[in_img]]
            setattr(blk, name, v)

    def shift_from_left(v):                       # previous pixel's value; lane 0's first pixel gets garbage -> use nan
        return np.concatenate([[np.nan], v[:-1]])

    def shift_from_right(v):
        return np.concatenate([v[1:], [np.nan]])

    def primal(y, blk, a12):
        blk.u = primal_rows(blk.u, blk.p11, blk.p12, shift_from_left(blk.p11), a12, last_col, y == ny - 1)

    def dual(y, c, d):
        has_below = y + 1 < ny
        return dual_rows(c.p11, c.p12, c.u, shift_from_right(c.u), d.u if has_below else np.zeros_like(c.u), last_col, has_below)

    def stage_a(y, blk, a12):
        load_row(y, blk)
        primal(y, blk, a12)

    def stage_b(y, c, d):
        q11, q12 = dual(y, c, d)
        c.p11 = np.where(outside_left, 0.0, q11)
        c.p12 = np.where(outside_left, 0.0, q12)

    def stage_d(y, c, d):
        q11, q12 = dual(y, c, d)
        out["u"][y, xs[own]] = c.u[own]
        out["p11"][y, xs[own]] = q11[own]
        out["p12"][y, xs[own]] = q12[own]

    def step(t, c, d, e):
        if t + 2 < ny:
            stage_a(t + 2, e, d.p12)
        if t + 1 < ny:
            stage_b(t + 1, d, e)
            primal(t + 1, d, c.p12)               # C
        stage_d(t, c, d)

    n = LANES * GROUP
    r0, r1, r2 = Block(n), Block(n), Block(n)
    a12 = np.zeros(n)
    if ys > 0:
        if ys > 1:
            a12[in_img] = p12[ys - 2, xs[in_img]]
        stage_a(ys - 1, r2, a12)
        a12 = r2.p12
    stage_a(ys, r0, a12)
    if ys + 1 < ny:
        stage_a(ys + 1, r1, r0.p12)
    else:
        r1.u, r1.p11, r1.p12 = r0.u.copy(), r0.p11.copy(), r0.p12.copy()
    if ys > 0:
        stage_b(ys - 1, r2, r0)
    stage_b(ys, r0, r1)
    primal(ys, r0, r2.p12)                        # C(ys); r2.p12 is zero on the first image row
    blocks = [r0, r1, r2]
    for t in range(ys, ye):
        k = t - ys
        step(t, blocks[k % 3], blocks[(k + 1) % 3], blocks[(k + 2) % 3])


@pytest.mark.parametrize("nx,ny,R", [(250, 40, 16), (120, 33, 16), (121, 17, 16), (119, 64, 32), (7, 5, 16), (1, 9, 16),
                                     (300, 1, 16), (241, 70, 32), (360, 2, 16), (5, 35, 32)])
def test_two_iteration_march_equals_two_global_iterations(nx, ny, R):
    rs = np.random.RandomState(nx * 1000 + ny)
    u, p11, p12 = (rs.uniform(-1, 1, (ny, nx)) for _ in range(3))
    ref = reference_iteration(*reference_iteration(u, p11, p12))
    out = {k: np.full((ny, nx), np.nan) for k in ("u", "p11", "p12")}
    for bx in range(-(-nx // W_OWN)):
        for ys in range(0, ny, R):
            emulate_strip(u, p11, p12, out, bx, ys, R)
    for k, r in zip(("u", "p11", "p12"), ref):
        assert not np.isnan(out[k]).any(), k                      # every pixel stored, no garbage lane value consumed
        assert np.array_equal(out[k], r), k
