"""The reference's own unit tests (tests/gtest_*.cc of flanggut/smvs), ported
to a dependency-free runner and applied to every CPU oracle of this repo:
the compiled-verbatim reference (oracle/_ref) and the plain restatement
(oracle/port). They pin the oracle before anything is compared with it.

Each test cites the gtest it restates."""
import numpy as np
import pytest

from oracle import ref as oref

IMPLS = []
if oref.available():
    IMPLS.append(oref.Units)
try:
    from oracle import port as oport
    if oport.available():
        IMPLS.append(oport.Units)
except ImportError:
    pass

pytestmark = pytest.mark.skipif(not IMPLS, reason="no CPU oracle built")


@pytest.fixture(params=IMPLS, ids=lambda u: u.name)
def U(request):
    return request.param


def nodes(*rows):
    return np.array(rows, dtype=np.float64).reshape(16)


# --- gtest_bicubic_patch.cc:16-162 (tolerance 1e-20 = exact) ---------------

def test_bicubic_linear_x(U):
    n = nodes([0, 1, 0, 0], [1, 1, 0, 0], [0, 1, 0, 0], [1, 1, 0, 0])
    assert np.array_equal(U.bicubic_eval(n, 0.5, 0.5), [0.5, 1, 0, 0, 0, 0])


def test_bicubic_linear_y(U):
    n = nodes([0, 0, 1, 0], [0, 0, 1, 0], [1, 0, 1, 0], [1, 0, 1, 0])
    assert np.array_equal(U.bicubic_eval(n, 0.5, 0.5), [0.5, 0, 1, 0, 0, 0])


def test_bicubic_linear_xy(U):
    n = nodes([0, .5, .5, 0], [.5, .5, .5, 0], [.5, .5, .5, 0], [1, .5, .5, 0])
    assert np.array_equal(U.bicubic_eval(n, 0.5, 0.5), [0.5, .5, .5, 0, 0, 0])


def test_bicubic_quadratic(U):
    n = nodes([10, 4, 4, -8], [10, -4, 4, -8], [10, 4, -4, -8], [10, -4, -4, -8])
    assert U.bicubic_eval(n, 0.5, 0.0)[0] == 11.0
    assert U.bicubic_eval(n, 0.0, 0.5)[0] == 11.0
    c = U.bicubic_eval(n, 0.5, 0.5)
    assert np.array_equal(c, [12, 0, 0, -2, -8, -8])
    assert U.bicubic_eval(n, 0.5, 0.0)[1] == 0.0
    assert U.bicubic_eval(n, 0.0, 0.5)[2] == 0.0


# --- gtest_bicubic_patch.cc:164-615: finite differences of all 4 x 24
#     node_derivatives entries at (0.9, 0.3), patch_to_pixel 0.2 ------------

def test_node_derivatives_fd(U):
    base = nodes([1, 2, 2, -4], [1, -2, 2, -4], [1, 2, -2, -4], [1, -2, -2, -4])
    cx, cy, p2p = 0.9, 0.3, 0.2
    scale = np.array([1, p2p, p2p, p2p * p2p, p2p * p2p, p2p * p2p])
    d = U.node_derivatives(cx, cy, 1.0 / p2p).reshape(4, 6, 4)
    v0 = U.bicubic_eval(base, cx, cy) * scale
    delta = 1e-4
    for node in range(4):
        for c in range(4):
            n2 = base.copy()
            n2[node * 4 + c] += delta
            fd = (U.bicubic_eval(n2, cx, cy) * scale - v0) / delta
            # layout per node: [f(4) dx(4) dy(4) dxy(4) dxx(4) dyy(4)]
            np.testing.assert_allclose(d[node, :, c], fd, atol=1e-8)


# --- gtest_correspondence.cc:17-260: d(corr)/d(node param), delta 1e-8,
#     tolerance 1e-4, golden M, t of :49-52 ---------------------------------

MM = [-0.997402, -0.0167178, 626.197, -0.0269324, -1.01116, 174.093,
      7.05365e-06, -0.000139764, 1.00931]
TT = [3.78737, 168.604, 0.0117067]


def test_correspondence_derivatives_fd(U):
    n = nodes([1.2, .2, .2, -.1], [1.4, -.3, .3, -.2], [1.1, .4, -.4, -.1],
              [1.3, -.2, -.2, -.1])
    u, v = 0.7, 0.4
    dn = U.node_derivatives(u, v)
    w = U.bicubic_eval(n, u, v)[0]
    base = U.correspondence(MM, TT, 100, 100, w, dn=dn)
    delta = 1e-8
    for col in range(16):
        n2 = n.copy()
        n2[col] += delta
        w2 = U.bicubic_eval(n2, u, v)[0]
        new = U.correspondence(MM, TT, 100, 100, w2)
        fd = (new["proj"] - base["proj"]) / delta
        np.testing.assert_allclose(base["c_dn"][col], fd, atol=1e-4)


# --- gtest_correspondence.cc:286-363: warp Jacobian vs finite differences --

def test_correspondence_jacobian_fd(U):
    n = nodes([8.2, .2, .2, 0], [9.4, -.3, .3, -.2], [10.1, .4, -.4, .1],
              [3.3, -.2, -.2, -.1])
    u, v, p2p = 0.2, 0.8, 0.8
    e = U.bicubic_eval(n, u, v)
    w, wx, wy = e[0], e[1] * p2p, e[2] * p2p
    x, y = 300.5, 200.5
    base = U.correspondence(MM, TT, x, y, w, wx, wy)
    delta, eps = 1e-8, 1e-5
    w2 = U.bicubic_eval(n, u + delta * p2p, v)[0]
    new = U.correspondence(MM, TT, x + delta, y, w2, wx, wy)
    fd = (new["proj"] - base["proj"]) / delta
    np.testing.assert_allclose(base["jac"][0:2], fd, atol=eps)
    w2 = U.bicubic_eval(n, u, v + delta * p2p)[0]
    new = U.correspondence(MM, TT, x, y + delta, w2, wx, wy)
    fd = (new["proj"] - base["proj"]) / delta
    np.testing.assert_allclose(base["jac"][2:4], fd, atol=eps)


# --- gtest_correspondence.cc:365-492: d(J grad)/d(node param) --------------

def test_correspondence_jacobian_derivative_grad_fd(U):
    n = nodes([8.2, .2, .2, 0], [9.4, -.3, .3, -.2], [10.1, .4, -.4, .1],
              [3.3, -.2, -.2, -.1])
    u, v, p2p = 0.2, 0.8, 0.8
    dn = U.node_derivatives(u, v, 1.0 / p2p)
    grad = np.array([0.3, -0.7])
    x, y = 300.5, 200.5

    def jac_of(nn):
        e = U.bicubic_eval(nn, u, v)
        return U.correspondence(MM, TT, x, y, e[0], e[1] * p2p, e[2] * p2p,
                                grad=grad, dn=dn)

    base = jac_of(n)
    jb = base["jac"].reshape(2, 2)
    delta = 1e-8
    for col in range(16):
        n2 = n.copy()
        n2[col] += delta
        jn = jac_of(n2)["jac"].reshape(2, 2)
        fd = ((jn - jb) / delta) @ grad
        np.testing.assert_allclose(base["jac_dn"][col], fd, atol=1e-4)


# --- gtest_surface_deriv.cc:208-375 normal_derivative, :377-468
#     normal_divergence vs FD of fill_normal, :502-666 normal_divergence_deriv

def _surf_setup(U, n, px, py, p2p):
    e = U.bicubic_eval(n, px, py)
    return (e[0], e[1] * p2p, e[2] * p2p, e[3] * p2p * p2p, e[4] * p2p * p2p,
            e[5] * p2p * p2p)


def test_normal_derivative_fd(U):
    n = nodes([10, 4, 4, -8], [10, -4, 4, -8], [10, 4, -4, -8], [10, -4, -4, -8])
    px, py, p2p = 0.7, 0.2, 0.2
    x, y, f = 100.0, 200.0, 500.0
    dn = U.node_derivatives(px, py, 1.0 / p2p)
    q = _surf_setup(U, n, px, py, p2p)
    base = U.surface_derivatives(dn, x, y, f, *q)
    delta = 1e-5
    for col in range(16):
        n2 = n.copy()
        n2[col] += delta
        q2 = _surf_setup(U, n2, px, py, p2p)
        new = U.surface_derivatives(dn, x, y, f, *q2)
        fd = (new["normal"] - base["normal"]) / delta
        got = base["normal_deriv"].reshape(3, 16)[:, col]
        np.testing.assert_allclose(got, fd, atol=1e-5)


def test_normal_divergence_fd(U):
    n = nodes([10, 4, 4, -8], [10, -4, 4, -8], [10, 4, -4, -8], [10, -4, -4, -8])
    px, py, ps = 0.7, 0.2, 5.0
    p2p = 1.0 / ps
    x, y, f = 100.0, 200.0, 500.0
    dn = U.node_derivatives(px, py, ps)
    base = U.surface_derivatives(dn, x, y, f, *_surf_setup(U, n, px, py, p2p))
    delta = 1e-6
    # move one pixel-fraction in x / y: patch coordinate moves delta / ps
    nx = U.surface_derivatives(dn, x + delta, y, f,
                               *_surf_setup(U, n, px + delta * p2p, py, p2p))
    ny = U.surface_derivatives(dn, x, y + delta, f,
                               *_surf_setup(U, n, px, py + delta * p2p, p2p))
    np.testing.assert_allclose(base["div"][0:3], (nx["normal"] - base["normal"]) / delta,
                               atol=1e-5)
    np.testing.assert_allclose(base["div"][3:6], (ny["normal"] - base["normal"]) / delta,
                               atol=1e-5)


def test_normal_divergence_deriv_fd(U):
    n = nodes([10, 4, 4, -8], [10, -4, 4, -8], [10, 4, -4, -8], [10, -4, -4, -8])
    px, py, p2p = 0.7, 0.2, 0.2
    x, y, f = 100.0, 200.0, 500.0
    dn = U.node_derivatives(px, py, 1.0 / p2p)
    base = U.surface_derivatives(dn, x, y, f, *_surf_setup(U, n, px, py, p2p))
    delta = 1e-7
    for col in range(16):
        n2 = n.copy()
        n2[col] += delta
        new = U.surface_derivatives(dn, x, y, f, *_surf_setup(U, n2, px, py, p2p))
        fd = (new["div"] - base["div"]) / delta
        got = base["div_deriv"].reshape(6, 16)[:, col]
        np.testing.assert_allclose(got, fd, atol=1e-5)


# --- gtest_spherical_harmonics.cc:17-60 -------------------------------------

def test_sh_derivative_fd(U):
    nrm = np.array([0.2, 0.3, 0.4])
    nrm /= np.linalg.norm(nrm)
    base, d = U.sh_4band(nrm)
    d = d.reshape(16, 3)
    delta = 1e-7
    for c in range(3):
        n2 = nrm.copy()
        n2[c] += delta
        fd = (U.sh_4band(n2)[0] - base) / delta
        np.testing.assert_allclose(d[:, c], fd, atol=1e-5)


# --- gtest_matrix_vector.cc:16-31 -------------------------------------------

def test_ldl_inverse_known_answer(U):
    A = U.ldl_inverse(np.array([[2, -1, 0], [-1, 2, -1], [0, -1, 1]], dtype=np.float64))
    np.testing.assert_allclose(A, [[1, 1, 1], [1, 2, 2], [1, 2, 3]], atol=1e-15)


def test_ldl_inverse_zero_pivot_leaves_input(U):
    """lib/ldl_decomposition.h:60-61: early return on an exactly-zero pivot."""
    A0 = np.array([[0.0, 1], [1, 0]])
    np.testing.assert_array_equal(U.ldl_inverse(A0), A0)


# --- gtest_matrix_vector.cc:33-195: SSEVector ----------------------------------

def _la(U):
    if not getattr(U, "has_linear_algebra", False):
        pytest.skip("SSEVector / BlockSparseMatrix are the reference's own classes: "
                    "only the compiled reference has them")
    return U


def test_ssevector_dot(U):
    U = _la(U)
    a = [1.0, 2.0, 3.0, 1.0, 2.0]
    assert U.ssevector("dot", a, a) == 19.0


@pytest.mark.parametrize("op,a,b,factor", [
    ("add", [4.0, 2, 3, 1, 2], [1.0, 7, 3, 8, 2], 0.0),
    ("subtract", [1.0, 2, 3, 10, 2], [11.0, 2, 30, 1, 29], 0.0),
    ("multiply", [12.0, 2, 3, 10, 2], None, 4.3),
    ("multiply_add", [12.0, 2, 3, 10, 2], [11.0, 2, 30, 1, 29], 4.3),
    ("multiply_sub", [12.0, 2, 3, 10, 2], [11.0, 2, 30, 1, 29], 4.3)])
def test_ssevector_elementwise(U, op, a, b, factor):
    U = _la(U)
    a = np.array(a)
    bb = None if b is None else np.array(b)
    want = {"add": lambda: a + bb, "subtract": lambda: a - bb,
            "multiply": lambda: a * factor, "multiply_add": lambda: a + bb * factor,
            "multiply_sub": lambda: a - bb * factor}[op]()
    np.testing.assert_array_equal(U.ssevector(op, a, bb, factor), want)   # EXPECT_DOUBLE_EQ


@pytest.mark.parametrize("op,ma,mb", [("multiply_add", 3, 7), ("multiply_sub", 7, 3)])
def test_ssevector_large(U, op, ma, mb):
    """multiply_add_large / multiply_sub_large: 1e6 entries, odd tail included."""
    U = _la(U)
    i = np.arange(1000000)
    a, b = (i % ma).astype(np.float64), (i % mb).astype(np.float64)
    want = a + b * 0.3 if op == "multiply_add" else a - b * 0.3
    np.testing.assert_array_equal(U.ssevector(op, a, b, 0.3), want)


# --- gtest_matrix_vector.cc:197-356: BlockSparseMatrix<2> ---------------------------

V1, V2 = [1.0, 2, 3, 4], [4.0, 3, 2, 1]


def test_bsm_set_from_blocks(U):
    U = _la(U)
    assert U.bsm2(4, blocks=[(0, 0, V1), (2, 2, V2)])[0] == 2
    assert U.bsm2(4, blocks=[(2, 2, V2), (0, 0, V1)])[0] == 2


def test_bsm_set_from_triplets(U):
    U = _la(U)
    trips = [(0, 0, 11), (0, 1, 12), (0, 2, 13), (0, 3, 14), (1, 1, 22), (1, 2, 23),
             (2, 2, 33), (2, 3, 34), (2, 4, 35), (2, 5, 36), (3, 3, 44), (3, 4, 45),
             (4, 5, 56), (5, 5, 66)]
    assert U.bsm2(6, triplets=trips)[0] == 5


def test_bsm_multiply(U):
    U = _la(U)
    v2 = [5.0, 3, 2, 0]
    ones = np.ones(4)
    for blocks in ([(0, 0, V1), (2, 2, v2)], [(2, 2, v2), (0, 0, V1)]):
        np.testing.assert_array_equal(U.bsm2(4, blocks=blocks, x=ones)[1], [3, 7, 8, 2])
    y = U.bsm2(4, blocks=[(2, 2, v2), (0, 2, V1), (0, 0, V1)], x=ones)[1]
    np.testing.assert_array_equal(y, [6, 14, 8, 2])


def test_bsm_set_from_triplets_multiply(U):
    U = _la(U)
    trips = [(0, 0, 1), (0, 1, 2), (1, 0, 3), (1, 1, 4), (2, 2, 5), (2, 3, 3),
             (3, 2, 2), (3, 3, 0)]
    ones = np.ones(4)
    np.testing.assert_array_equal(U.bsm2(4, triplets=trips, x=ones)[1], [3, 7, 8, 2])
    trips += [(2, 0, 2), (2, 1, 7), (3, 0, 4), (3, 1, 1)]
    y1 = U.bsm2(4, triplets=trips, x=ones)[1]
    y2 = U.bsm2(4, triplets=trips, x=ones)[1]
    np.testing.assert_array_equal(y1, y2)
    np.testing.assert_array_equal(y1, [3, 7, 17, 7])      # rows 2, 3 gained 2+7, 4+1


def test_bsm_block_invert(U):
    U = _la(U)
    two = [2.0, 0, 0, 2]
    y = U.bsm2(4, blocks=[(0, 0, two), (2, 2, two)], invert=True, x=np.ones(4))[1]
    np.testing.assert_allclose(y, 0.5, atol=1e-10)
