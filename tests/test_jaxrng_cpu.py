"""unidom_b200.jaxrng (NumPy threefry2x32 in the layout of the jax 0.3.14 the reference pins) against known answers:
the Random123 Threefry-2x32 vectors, and the key splits / draws JAX's own documentation prints for PRNGKey(0) and
PRNGKey(42).  JAX itself is not installable in this image; these constants are the published ones."""
import numpy as np

from unidom_b200 import jaxrng as R


def test_threefry2x32_known_answers():
    def kat(k, c):
        y0, y1 = R.threefry2x32(np.array(k, np.uint32), np.array([c[0]], np.uint32), np.array([c[1]], np.uint32))
        return int(y0[0]), int(y1[0])
    assert kat((0, 0), (0, 0)) == (0x6B200159, 0x99BA4EFE)
    assert kat((0xFFFFFFFF, 0xFFFFFFFF), (0xFFFFFFFF, 0xFFFFFFFF)) == (0x1CB996FC, 0xBB002BE7)
    assert kat((0x13198A2E, 0x03707344), (0x243F6A88, 0x85A308D3)) == (0xC4923A9C, 0x483DF7A0)


def test_split_uniform_normal_match_the_documented_draws():
    k = R.PRNGKey(0)
    assert k.tolist() == [0, 0] and R.PRNGKey(42).tolist() == [0, 42]
    assert R.split(k).tolist() == [[4146024105, 967050713], [2718843009, 1272950319]]
    assert R.split(k, 3).shape == (3, 2)
    assert abs(float(R.uniform(k, ())) - 0.41845703) < 1e-8          # exact integer path
    assert abs(float(R.normal(k, ())) - (-0.20584226)) < 2e-7
    np.testing.assert_allclose(R.normal(k, (3,)), [1.8160863, -0.48262316, 0.33988908], rtol=0, atol=3e-7)
    assert abs(float(R.normal(R.PRNGKey(42), ())) - (-0.18471177)) < 2e-7
    u = R.uniform(R.PRNGKey(7), (1001, 3))                            # odd count: the padded lane
    assert u.dtype == np.float32 and u.shape == (1001, 3) and float(u.min()) >= 0.0 and float(u.max()) < 1.0
    assert abs(float(u.mean()) - 0.5) < 0.02
    n = R.normal(R.PRNGKey(7), (20000,))
    assert abs(float(n.mean())) < 0.03 and abs(float(n.std()) - 1.0) < 0.03 and np.isfinite(n).all()


def test_cloth_reset_shift_is_the_reference_stream():
    """cloth_env.py:181-185 with apg.py's PRNGKey(0): key, _ = split(key); shift = normal(key, (2,)) * 0.05."""
    key = R.split(R.PRNGKey(0))[0]
    shift = R.normal(key, (2,)) * np.float32(0.05)
    assert key.tolist() == [4146024105, 967050713]
    assert shift.dtype == np.float32 and np.all(np.abs(shift) < 0.25)


def test_unfold_reset_noise_and_fold_points_are_the_reference_draws():
    """unfold_cloth1_env.py:74-79: key, _ = split(PRNGKey(1)); x = lattice + normal(key, x.shape) * 1e-4, then the fold end
    points from np.random -- against what the reference's own UnfoldCloth1Env.reset drew under the shim
    (tests/golden/ref_clothenv_unfold1.npz, gen_golden.py::cloth_unfold_case), bit for bit."""
    import os
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_clothenv_unfold1.npz"))
    k = R.split(R.PRNGKey(int(d["key_seed"])))[0]
    noise = R.normal(k, d["lattice_x"].shape) * np.float32(0.0001)
    assert noise.dtype == np.float32
    assert np.array_equal((d["lattice_x"] + noise).astype(np.float32), d["noisy_x"])
    rs = np.random.RandomState(int(d["seed"]))
    P = d["lattice_x"].shape[1]
    assert np.array_equal(rs.randint(0, P, size=(2,)), d["st_point"]) and np.array_equal(rs.randint(0, P, size=(2,)), d["ed_point"])
