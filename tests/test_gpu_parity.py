"""GPU parity tests: libartes_gpu (through the C-ABI) against the CPU oracle and the committed
golden fixtures.  Bar (BASELINE.json north_star): cell-crossing sequences bit-exact for an injected
random stream; images within Monte Carlo tolerance (3 sigma of the combined photon noise)."""
import math
import os

import numpy as np
import pytest

from artes_b200 import abi, host
from artes_b200.abi import make_launch
from tools import atmospheres as A
from test_oracle import random_interior_points

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MODES = [abi.MODE_FAITHFUL, abi.MODE_FAST]


def oblate_pair(atm, oblateness):
    """oracle + gpu contexts with an oblate planet (src/ARTES.f90:469-471)."""
    from oracle_lib import Oracle
    from artes_b200.lib import GpuTransport
    ox = 1.0 / (1.0 - oblateness)
    o = Oracle(); depth = o.set_atmosphere(atm, oblateness=oblateness)
    g = GpuTransport((0,))
    g.set_grid(atm.rfront, atm.thetafront(), atm.thetaplane(), atm.phifront(), (ox, ox, 1.0))
    g.set_wavelength(atm.k_sca[0], atm.k_abs[0], atm.uniq[0], atm.cell_to_uniq[0], depth)
    return o, g


# ---- cell_face ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["c1_template_rayleigh", "c2_hg_deck", "c4_mie_patches"])
def test_cell_face_bit_exact(atmospheres, oracle_factory, gpu_factory, name):
    atm = atmospheres(name)
    o, _ = oracle_factory(atm)
    g, _ = gpu_factory(atm)
    pos, d, face, cell = random_interior_points(atm, 100000, 11)
    oi, od = o.cell_face(pos, d, face, cell)
    gi, gd = g.cell_face(pos, d, face, cell, mode=abi.MODE_FAITHFUL)
    np.testing.assert_array_equal(oi, gi)
    np.testing.assert_array_equal(od, gd)          # bit-exact distances
    # second step: start ON the face just reached, in the next cell
    ok = (oi[:, 6] == 0) & (oi[:, 5] == 0) & (oi[:, 2] >= 0)
    pos2 = (pos + od[:, None] * d)[ok]
    oi2, od2 = o.cell_face(pos2, d[ok], oi[ok, 0:2].copy(), oi[ok, 2:5].copy())
    gi2, gd2 = g.cell_face(pos2, d[ok], oi[ok, 0:2].copy(), oi[ok, 2:5].copy(), mode=abi.MODE_FAITHFUL)
    np.testing.assert_array_equal(oi2, gi2)
    np.testing.assert_array_equal(od2, gd2)
    # fast mode: same topology, distances to rounding
    fi, fd = g.cell_face(pos, d, face, cell, mode=abi.MODE_FAST)
    same = np.all(fi == oi, axis=1)
    assert same.mean() > 0.9999
    np.testing.assert_allclose(fd[same], od[same], rtol=1e-6)


def test_cell_face_oblate_bit_exact(atmospheres):
    atm = atmospheres("c4_mie_patches")
    o, g = oblate_pair(atm, 0.06)
    pos, d, face, cell = random_interior_points(atm, 50000, 12)
    pos[:, 0:2] *= 1.0 / (1.0 - 0.06)
    oi, od = o.cell_face(pos, d, face, cell)
    gi, gd = g.cell_face(pos, d, face, cell, mode=abi.MODE_FAITHFUL)
    np.testing.assert_array_equal(oi, gi)
    np.testing.assert_array_equal(od, gd)


# ---- injected-stream walk parity -------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["c1", "c2", "c4"])
@pytest.mark.parametrize("mode", MODES)
def test_trace_matches_committed_golden(atmospheres, gpu_factory, name, mode):
    from golden.make_golden import build_case
    gold = np.load(os.path.join(GOLDEN, f"golden_{name}.npz"))
    atm, lt, lr, xi = build_case(name, atmospheres)
    g, depth = gpu_factory(atm)
    assert depth == int(gold["cell_depth"])
    lt.mode = mode
    t = g.trace(lt, xi)
    np.testing.assert_array_equal(t["len"], gold["seq_len"])
    np.testing.assert_array_equal(t["hash"], gold["seq_hash"])
    np.testing.assert_allclose(t["fstate"][:, 3:7], gold["fstate"][:, 3:7], rtol=1e-6, atol=1e-12)
    lr.mode = mode
    r = g.run(lr)
    np.testing.assert_array_equal(r["det"][2], gold["det"][2])                       # counts: integers
    np.testing.assert_allclose(r["det"][0].sum(axis=(1, 2)), gold["det"][0].sum(axis=(1, 2)), rtol=1e-6, atol=1e-9)
    assert r["stats"]["n_cell_face"] == int(gold["n_cell_face"]) and r["stats"]["n_scatter"] == int(gold["n_scatter"])


CASES_LIVE = [
    ("c1_template_rayleigh", dict(), 40000),
    ("c2_hg_deck", dict(nx=1, ny=1, det_phi=math.radians(140.0)), 40000),
    ("c4_mie_patches", dict(nx=64, ny=64, det_phi=math.radians(60.0)), 60000),
    ("c4_mie_patches", dict(surface_albedo=0.7, det_phi=math.radians(20.0)), 20000),
    ("c4_mie_patches", dict(stellar_direction=1, theta_star=math.radians(70.0), phi_star=math.radians(33.0)), 20000),
    ("c2_hg_deck", dict(limb_emission=1, nx=1, ny=1, det_phi=math.radians(175.0)), 20000),
    ("c3_molecular", dict(nx=1, ny=1), 100000),      # SURVEY 8d gate: >= 1e5 photons per grid class (1-D here; 2-D: 100k, 3-D: 106k)
    ("c5_scale", dict(nx=16, ny=16, det_phi=math.radians(60.0)), 6000),
]


@pytest.mark.parametrize("name,kw,n", CASES_LIVE)
@pytest.mark.parametrize("mode", MODES)
def test_crossing_sequences_bit_exact_vs_oracle(atmospheres, oracle_factory, gpu_factory, name, kw, n, mode):
    """>= 1e5 injected-stream photons over the grid classes (1-D, 2-D with the equatorial plane face,
    3-D with phi wrap): every (face, cell) sequence identical."""
    atm = atmospheres(name)
    o, _ = oracle_factory(atm)
    g, _ = gpu_factory(atm)
    xm = 1.3 * atm.rfront[-1]
    xi = np.random.RandomState(hash(name) % 1000 + n).random_sample((n, 160))
    L = make_launch(mode=mode, n_photons=n, x_max=xm, y_max=xm, fstop=0.03, **kw)
    ro = o.trace(L, xi, max_rec=4)
    rg = g.trace(L, xi, max_rec=4)
    # fast mode walks the ray/event engine (the production path); the faithful mode the persistent-lane engine
    assert g.last_engine() == (2 if mode == abi.MODE_FAST else 3)
    same = (ro["hash"] == rg["hash"]) & (ro["len"] == rg["len"])
    assert same.all(), f"{(~same).sum()} of {n} sequences differ"
    np.testing.assert_array_equal(ro["head"], rg["head"])
    assert ro["len"].mean() > 5
    np.testing.assert_allclose(rg["fstate"][:, 3], ro["fstate"][:, 3], rtol=1e-5, atol=1e-14)   # Stokes I
    assert (rg["fstate"][:, 7] == ro["fstate"][:, 7]).all()                                        # n_scatter


def test_trace_thermal_source(atmospheres, oracle_factory):
    from artes_b200.lib import GpuTransport
    atm = atmospheres("c3_molecular")
    o, depth = oracle_factory(atm, 0, 2)
    vol = host.cell_volume(atm.rfront, atm.thetafront(), atm.phifront())
    cw, lum, cdf = host.thermal_tables(depth, atm.k_abs[0], atm.temperature, vol, atm.wavelengths[0] * 1e-6,
                                       atm.nr, atm.ntheta, atm.nphi)
    o.set_wavelength(atm.k_sca[0], atm.k_abs[0], atm.uniq[0], atm.cell_to_uniq[0], depth, cw, cdf)
    g = GpuTransport((0,))
    g.set_grid(atm.rfront, atm.thetafront(), atm.thetaplane(), atm.phifront())
    g.set_wavelength(atm.k_sca[0], atm.k_abs[0], atm.uniq[0], atm.cell_to_uniq[0], depth, cw, cdf)
    xm = 1.3 * atm.rfront[-1]
    n = 20000
    xi = np.random.RandomState(5).random_sample((n, 160))
    for emission in (1, 2):
        for mode in MODES:
            L = make_launch(mode=mode, n_photons=n, x_max=xm, y_max=xm, fstop=0.03, photon_source=2,
                            photon_emission=emission, nx=1, ny=1)
            ro, rg = o.trace(L, xi), g.trace(L, xi)
            assert g.last_engine() == (2 if mode == abi.MODE_FAST else 3)
            assert ((ro["hash"] == rg["hash"]) & (ro["len"] == rg["len"])).all()
        L = make_launch(mode=abi.MODE_FAST, n_photons=n, x_max=xm, y_max=xm, seed=3, photon_source=2,
                        photon_emission=emission, nx=1, ny=1)
        a, b = o.run(L), g.run(L)
        np.testing.assert_allclose(b["flux"], a["flux"], rtol=1e-9)
        np.testing.assert_allclose(b["det"][0, 0], a["det"][0, 0], rtol=1e-6)


def test_trace_oblate_planet(atmospheres):
    atm = atmospheres("c4_mie_patches")
    o, g = oblate_pair(atm, 0.06)
    xm = 1.06 * 1.3 * atm.rfront[-1]
    n = 15000
    xi = np.random.RandomState(8).random_sample((n, 160))
    for mode in MODES:
        L = make_launch(mode=mode, n_photons=n, x_max=xm, y_max=xm, fstop=0.03, surface_albedo=0.5)
        ro, rg = o.trace(L, xi), g.trace(L, xi)
        assert ((ro["hash"] == rg["hash"]) & (ro["len"] == rg["len"])).all()


# ---- images -----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,kw", [("c1_template_rayleigh", dict()),
                                     ("c4_mie_patches", dict(nx=64, ny=64, det_phi=math.radians(60.0), surface_albedo=0.2))])
@pytest.mark.parametrize("mode", MODES)
def test_same_stream_images_agree(atmospheres, oracle_factory, gpu_factory, name, kw, mode):
    """Same Philox stream on both sides: identical event counts, images equal to rounding."""
    atm = atmospheres(name)
    o, _ = oracle_factory(atm)
    g, _ = gpu_factory(atm)
    xm = 1.3 * atm.rfront[-1]
    L = make_launch(mode=mode, n_photons=60000, x_max=xm, y_max=xm, seed=21, flow_theta=1, flow_global=1, **kw)
    a = o.run(L, flows=True)
    b = g.run(L, flows=True)
    for k in ("n_emit", "n_cell_face", "n_scatter", "n_peel", "n_surface", "n_draws", "n_error"):
        assert a["stats"][k] == b["stats"][k], k
    np.testing.assert_array_equal(a["det"][2], b["det"][2])
    # libm differences (CUDA vs glibc, <= 2 ulp) are amplified along a multiple-scattering path, so single
    # deposits agree to ~1e-5 relative; the sums agree far better.
    scale = np.abs(a["det"][0]).max()
    np.testing.assert_allclose(b["det"][0], a["det"][0], rtol=1e-4, atol=1e-7 * scale)
    np.testing.assert_allclose(b["det"][1], a["det"][1], rtol=1e-4, atol=1e-9 * scale * scale)
    np.testing.assert_allclose(b["det"][0].sum(axis=(1, 2)), a["det"][0].sum(axis=(1, 2)), rtol=1e-7, atol=1e-9 * scale)
    np.testing.assert_array_equal(a["err"], b["err"])
    np.testing.assert_allclose(b["flow4"], a["flow4"], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(b["flow3"], a["flow3"], rtol=1e-6, atol=1e-3 * np.abs(a["flow3"]).max())


# The fast mode's production path is the ray/event engine (artes_b200/csrc/engine2.cuh): star source, black
# surface, no flow counters.  Same Philox stream and draw order as the oracle, so the trajectories agree to
# rounding: identical event counts up to the odd threshold decision that falls within an ulp, images to ~1e-5.
E2_CASES = [
    ("c1_template_rayleigh", dict(), 60000),
    ("c2_hg_deck", dict(nx=1, ny=1, det_phi=math.radians(140.0)), 60000),
    ("c2_hg_deck", dict(nx=8, ny=8, limb_emission=1, det_phi=math.radians(175.0)), 30000),
    ("c3_molecular", dict(nx=1, ny=1), 20000),
    ("c4_mie_patches", dict(nx=64, ny=64, det_phi=math.radians(60.0)), 60000),
    ("c4_mie_patches", dict(nx=16, ny=16, stellar_direction=1, theta_star=math.radians(70.0), phi_star=math.radians(33.0)), 30000),
    ("c5_scale", dict(nx=64, ny=64, det_phi=math.radians(60.0)), 8000),
]


@pytest.mark.parametrize("name,kw,n", E2_CASES)
def test_ray_event_engine_same_stream_vs_oracle(atmospheres, oracle_factory, gpu_factory, name, kw, n):
    atm = atmospheres(name)
    o, _ = oracle_factory(atm)
    g, _ = gpu_factory(atm)
    xm = 1.3 * atm.rfront[-1]
    L = make_launch(mode=abi.MODE_FAST, n_photons=n, x_max=xm, y_max=xm, seed=33, **kw)
    a, b = o.run(L), g.run(L)
    assert g.last_engine() == 2
    for k in ("n_emit", "n_cell_face", "n_scatter", "n_peel", "n_surface", "n_draws"):
        assert abs(a["stats"][k] - b["stats"][k]) <= max(2, 2e-5 * a["stats"][k]), (k, a["stats"][k], b["stats"][k])
    assert b["stats"]["n_error"] <= a["stats"]["n_error"] + 2
    assert np.abs(a["det"][2] - b["det"][2]).sum() <= max(4, 2e-5 * a["det"][2].sum())
    scale = np.abs(a["det"][0]).max()
    same = a["det"][2] == b["det"][2]
    np.testing.assert_allclose(b["det"][0][same], a["det"][0][same], rtol=2e-4, atol=1e-6 * scale)
    np.testing.assert_allclose(b["det"][0].sum(axis=(1, 2)), a["det"][0].sum(axis=(1, 2)), rtol=2e-5, atol=1e-7 * scale)
    np.testing.assert_allclose(b["det"][1].sum(axis=(1, 2)), a["det"][1].sum(axis=(1, 2)), rtol=2e-4, atol=1e-9 * scale * scale)


def test_ray_event_engine_surface_flows_thermal(atmospheres, oracle_factory, gpu_factory):
    """The general paths of the ray/event engine: Lambert surface + peel_surface, latitudinal flow counters,
    thermal source + peel_thermal -- same Philox stream as the oracle."""
    from artes_b200.lib import GpuTransport
    atm = atmospheres("c4_mie_patches")
    o, _ = oracle_factory(atm)
    g, _ = gpu_factory(atm)
    xm = 1.3 * atm.rfront[-1]
    L = make_launch(mode=abi.MODE_FAST, n_photons=50000, x_max=xm, y_max=xm, seed=17, nx=32, ny=32, det_phi=math.radians(40.0),
                    surface_albedo=0.4, flow_theta=1)
    a, b = o.run(L, flows=True), g.run(L, flows=True)
    assert g.last_engine() == 2
    for k in ("n_emit", "n_cell_face", "n_scatter", "n_peel", "n_surface", "n_draws"):
        assert abs(a["stats"][k] - b["stats"][k]) <= max(2, 2e-5 * a["stats"][k]), (k, a["stats"][k], b["stats"][k])
    assert a["stats"]["n_surface"] > 1000 and a["stats"]["n_peel"] > a["stats"]["n_scatter"]      # surface peels happened
    scale = np.abs(a["det"][0]).max()
    np.testing.assert_allclose(b["det"][0].sum(axis=(1, 2)), a["det"][0].sum(axis=(1, 2)), rtol=5e-5, atol=1e-7 * scale)
    assert np.abs(a["det"][2] - b["det"][2]).sum() <= max(4, 5e-5 * a["det"][2].sum())
    np.testing.assert_allclose(b["flow4"], a["flow4"], rtol=1e-4, atol=2e-4 * np.abs(a["flow4"]).max())
    assert np.abs(a["flow4"]).sum() > 0
    # thermal source
    atm3 = atmospheres("c3_molecular")
    o3, depth = oracle_factory(atm3, 0, 2)
    vol = host.cell_volume(atm3.rfront, atm3.thetafront(), atm3.phifront())
    cw, lum, cdf = host.thermal_tables(depth, atm3.k_abs[0], atm3.temperature, vol, atm3.wavelengths[0] * 1e-6,
                                       atm3.nr, atm3.ntheta, atm3.nphi)
    o3.set_wavelength(atm3.k_sca[0], atm3.k_abs[0], atm3.uniq[0], atm3.cell_to_uniq[0], depth, cw, cdf)
    g3 = GpuTransport((0,))
    g3.set_grid(atm3.rfront, atm3.thetafront(), atm3.thetaplane(), atm3.phifront())
    g3.set_wavelength(atm3.k_sca[0], atm3.k_abs[0], atm3.uniq[0], atm3.cell_to_uniq[0], depth, cw, cdf)
    xm3 = 1.3 * atm3.rfront[-1]
    for emission in (1, 2):
        L3 = make_launch(mode=abi.MODE_FAST, n_photons=30000, x_max=xm3, y_max=xm3, seed=5, photon_source=2, photon_emission=emission,
                         nx=4, ny=4, flow_theta=1)
        a3, b3 = o3.run(L3, flows=True), g3.run(L3, flows=True)
        assert g3.last_engine() == 2
        np.testing.assert_allclose(b3["flux"], a3["flux"], rtol=1e-8)
        np.testing.assert_allclose(b3["det"][0, 0].sum(), a3["det"][0, 0].sum(), rtol=2e-5)
        np.testing.assert_allclose(b3["flow4"], a3["flow4"], rtol=1e-4, atol=2e-4 * np.abs(a3["flow4"]).max())
        assert abs(a3["stats"]["n_cell_face"] - b3["stats"]["n_cell_face"]) <= max(2, 2e-5 * a3["stats"]["n_cell_face"])


def test_fast_mode_oblate_planet_same_stream(atmospheres):
    """Oblate launches are outside the ray/event engine's scope (launchers.inc) and run on the persistent-lane engine."""
    atm = atmospheres("c4_mie_patches")
    o, g = oblate_pair(atm, 0.06)
    xm = 1.06 * 1.3 * atm.rfront[-1]
    L = make_launch(mode=abi.MODE_FAST, n_photons=40000, x_max=xm, y_max=xm, seed=9, nx=32, ny=32, det_phi=math.radians(100.0))
    a, b = o.run(L), g.run(L)
    assert g.last_engine() == 3
    assert abs(a["stats"]["n_cell_face"] - b["stats"]["n_cell_face"]) <= max(2, 2e-5 * a["stats"]["n_cell_face"])
    assert a["stats"]["n_scatter"] == b["stats"]["n_scatter"] or abs(a["stats"]["n_scatter"] - b["stats"]["n_scatter"]) <= 3
    np.testing.assert_allclose(b["det"][0].sum(axis=(1, 2)), a["det"][0].sum(axis=(1, 2)), rtol=2e-5, atol=1e-9)


# (the independent-stream statistical gate lives in tests/test_gpu_statistical.py: all five configurations, SURVEY 8d thresholds)


def test_lambert_sphere_on_gpu(gpu_factory):
    atm = A.lambert_sphere()
    g, _ = gpu_factory(atm)
    n = 400000
    xm = 1.3 * atm.rfront[-1]
    for a_deg in (30.0, 110.0):
        a = math.radians(a_deg)
        for mode in MODES:
            r = g.run(make_launch(mode=mode, n_photons=n, x_max=xm, y_max=xm, seed=3, surface_albedo=1.0, det_phi=a, nx=1, ny=1))
            hit = (atm.rfront[0] / atm.rfront[-1]) ** 2
            expect = hit * (2.0 / (3.0 * math.pi)) * (math.sin(a) + (math.pi - a) * math.cos(a)) / math.pi
            assert abs(r["det"][0, 0].sum() / n / expect - 1.0) < 0.008


# ---- API behaviour ---------------------------------------------------------------------------------------------
def test_dense_matrix_entry_equals_compact(atmospheres, gpu_factory):
    atm = atmospheres("c2_hg_deck")
    g, depth = gpu_factory(atm)
    xm = 1.3 * atm.rfront[-1]
    L = make_launch(mode=abi.MODE_FAITHFUL, n_photons=30000, x_max=xm, y_max=xm, seed=4)
    a = g.run(L)
    g.set_wavelength_dense(atm.k_sca[0], atm.k_abs[0], atm.dense_matrix(0), depth)
    b = g.run(L)
    np.testing.assert_array_equal(a["det"][2], b["det"][2])
    np.testing.assert_allclose(a["det"][0], b["det"][0], rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("name,px", [("c1_template_rayleigh", 25), ("c4_mie_patches", 32)])
def test_eight_element_matrix_table_equals_the_full_matrices(atmospheres, gpu_factory, name, px):
    """Block-diagonal scattering matrices (all the reference's opacity tools make them) are read from an eight-element copy
    (DevTables::Mc, artes_gpu.cu); artes_gpu_test_full_matrix forces the 16-element path.  The terms the short path leaves out are
    products with exact zeros: same photons, same pixels, sums equal up to the order of the detector's atomic additions.  A matrix
    with one non-zero element outside the diagonal quarters must switch the short path off by itself."""
    atm = atmospheres(name)
    xm = 1.3 * atm.rfront[-1]
    L = make_launch(mode=abi.MODE_FAST, n_photons=60000, x_max=xm, y_max=xm, nx=px, ny=px, seed=9)
    g, depth = gpu_factory(atm)
    a = g.run(L)
    assert a["stats"]["n_scatter"] > 100000
    g.lib.artes_gpu_test_full_matrix(1)
    try:
        g.set_wavelength(atm.k_sca[0], atm.k_abs[0], atm.uniq[0], atm.cell_to_uniq[0], depth)
        b = g.run(L)
    finally:
        g.lib.artes_gpu_test_full_matrix(0)
    for k in ("n_emit", "n_cell_face", "n_scatter", "n_peel"):
        assert a["stats"][k] == b["stats"][k]
    np.testing.assert_array_equal(a["det"][2], b["det"][2])
    np.testing.assert_allclose(a["det"][0], b["det"][0], rtol=1e-10, atol=1e-13 * np.abs(b["det"][0]).max())
    # F13 != 0 somewhere: not block-diagonal any more -> the general path, and a different answer from the zero-F13 tables
    uniq = np.array(atm.uniq[0], dtype=np.float64, copy=True).reshape(-1, 180, 16)
    uniq[0, 40, 2] = 1e-3 * uniq[0, 40, 0]
    g.set_wavelength(atm.k_sca[0], atm.k_abs[0], uniq.reshape(np.shape(atm.uniq[0])), atm.cell_to_uniq[0], depth)
    c = g.run(L)
    g.set_wavelength(atm.k_sca[0], atm.k_abs[0], atm.uniq[0], atm.cell_to_uniq[0], depth)
    d = g.run(L)
    assert np.abs(c["det"][0] - b["det"][0]).max() > 0.0
    np.testing.assert_array_equal(d["det"][2], a["det"][2])
    np.testing.assert_allclose(d["det"][0], a["det"][0], rtol=1e-10, atol=1e-13 * np.abs(a["det"][0]).max())


def test_dense_whole_array_entry_picks_the_wavelength_and_dedups_on_the_device(atmospheres):
    """artes_gpu_set_wavelength_dense_wl takes the reference's WHOLE arrays (cells, n_wl[, 16, 180]) as they sit in memory
    and picks one wavelength with strides (no Fortran slice copy); the (element, angle) planes are hashed and verified on
    the device.  Results must equal the compact entry for that wavelength."""
    from artes_b200.lib import GpuTransport
    atm = A.c5_scale(nr=12, ntheta=10, nphi=16, nl=3)
    nl = len(atm.wavelengths)
    dense_all = np.ascontiguousarray(np.stack([atm.dense_matrix(l) for l in range(nl)], axis=2))     # (180, 16, nl, nphi, ntheta, nr)
    xm = 1.3 * atm.rfront[-1]
    for l in (0, 2):
        depth = host.cell_depth(atm.rfront, atm.k_sca[l], atm.k_abs[l], atm.nr, atm.ntheta, atm.nphi, 1)
        ga, gb = GpuTransport((0,)), GpuTransport((0,))
        for g in (ga, gb):
            g.set_grid(atm.rfront, atm.thetafront(), atm.thetaplane(), atm.phifront())
        ga.set_wavelength(atm.k_sca[l], atm.k_abs[l], atm.uniq[l], atm.cell_to_uniq[l], depth)
        # wavelength 0 through the resident path, wavelength 2 through the chunked two-pass path (planes that do not fit into HBM)
        gb.lib.artes_gpu_test_ingest_chunk(0 if l == 0 else 97)
        gb.set_wavelength_dense_wl(atm.k_sca, atm.k_abs, dense_all, l, depth)
        gb.lib.artes_gpu_test_ingest_chunk(0)
        L = make_launch(mode=abi.MODE_FAST, n_photons=30000, x_max=xm, y_max=xm, seed=4, nx=16, ny=16, det_phi=math.radians(50.0))
        a, b = ga.run(L), gb.run(L)
        np.testing.assert_array_equal(a["det"][2], b["det"][2])
        np.testing.assert_allclose(a["det"][0], b["det"][0], rtol=1e-9, atol=1e-12 * np.abs(a["det"][0]).max())
        for k in ("n_cell_face", "n_scatter", "n_draws"):
            assert a["stats"][k] == b["stats"][k]
        ga.close(); gb.close()


def test_trace_hook_walks_the_tables_of_its_wavelength(atmospheres):
    """After artes_gpu_set_wavelengths the trace hook must walk the tables of launch.wl_index (it used to walk wavelength 0):
    crossing sequences equal the oracle loaded with that wavelength, and differ from wavelength 0's."""
    from oracle_lib import Oracle
    atm = atmospheres("c3_molecular")
    wls = [0, 9, 31]
    t = host.Transport(atm, host.Params(nx=1, ny=1), mode=abi.MODE_FAST)
    t.set_all_wavelengths(wls)
    xi = np.random.RandomState(5).random_sample((3000, 160))
    hashes = []
    for k, l in enumerate(wls):
        o = Oracle(); depth = o.set_atmosphere(atm, l)
        assert depth == t.depths[k]
        for mode in MODES:
            L = make_launch(mode=mode, n_photons=3000, x_max=t.x_max, y_max=t.x_max, fstop=0.03, nx=1, ny=1, wl_index=k)
            ro, rg = o.trace(L, xi), t.gpu.trace(L, xi)
            assert ((ro["hash"] == rg["hash"]) & (ro["len"] == rg["len"])).all(), (l, mode)
        hashes.append(rg["hash"])
    assert (hashes[0] != hashes[2]).mean() > 0.5
    t.close()


def test_async_equals_sync_and_empty_launch(atmospheres, gpu_factory):
    atm = atmospheres("c1_template_rayleigh")
    g, _ = gpu_factory(atm)
    xm = 1.3 * atm.rfront[-1]
    L = make_launch(mode=abi.MODE_FAST, n_photons=20000, x_max=xm, y_max=xm, seed=6)
    a = g.run(L)
    g.run_async(L)
    b = g.wait()
    np.testing.assert_array_equal(a["det"][2], b["det"][2])
    np.testing.assert_allclose(a["det"][0], b["det"][0], rtol=1e-9, atol=1e-12)
    z = g.run(make_launch(mode=abi.MODE_FAST, n_photons=0, x_max=xm, y_max=xm))
    assert z["det"].sum() == 0 and z["stats"]["n_emit"] == 0


def test_photon_id_sharding_is_additive(atmospheres, gpu_factory):
    """Two shards with disjoint photon-id ranges = one launch over the union (multi-GPU invariant)."""
    atm = atmospheres("c2_hg_deck")
    g, _ = gpu_factory(atm)
    xm = 1.3 * atm.rfront[-1]
    kw = dict(mode=abi.MODE_FAST, x_max=xm, y_max=xm, seed=12, nx=8, ny=8)
    full = g.run(make_launch(n_photons=50000, **kw))
    s1 = g.run(make_launch(n_photons=20000, photon_id_base=0, **kw))
    s2 = g.run(make_launch(n_photons=30000, photon_id_base=20000, **kw))
    np.testing.assert_array_equal(full["det"][2], s1["det"][2] + s2["det"][2])
    np.testing.assert_allclose(full["det"][0], s1["det"][0] + s2["det"][0], rtol=1e-9, atol=1e-12)
    assert full["stats"]["n_cell_face"] == s1["stats"]["n_cell_face"] + s2["stats"]["n_cell_face"]


def _phase_launches(n, n_photons, det_phis, **kw):
    out = []
    for phi in det_phis[:n]:
        out.append(make_launch(n_photons=n_photons, det_phi=math.radians(phi), limb_emission=int(phi >= 170.0), **kw))
    return out


@pytest.mark.parametrize("name,extra", [("c2_hg_deck", dict(nx=1, ny=1)), ("c4_mie_patches", dict(nx=16, ny=16)),
                                        ("c4_mie_patches", dict(nx=8, ny=8, surface_albedo=0.5))])
def test_batched_launches_equal_single_launches(atmospheres, gpu_factory, name, extra):
    """artes_gpu_run_batch (the phase-curve loop :215-245 as one kernel): image k of the batch equals the single
    launch with det_phi_k and photon_id_base = k * n_photons, up to the order of the floating-point sums."""
    atm = atmospheres(name)
    g, _ = gpu_factory(atm)
    xm = 1.3 * atm.rfront[-1]
    kw = dict(mode=abi.MODE_FAST, x_max=xm, y_max=xm, seed=21, **extra)
    phis = [0.0, 35.0, 90.0, 145.0, 175.0, 180.0, 250.0]
    P = 30000
    Ls = _phase_launches(len(phis), P, phis, **kw)
    b = g.run_batch(Ls)
    assert g.last_engine() == 2 and b["stats"]["reserved"] == 1          # one kernel for the whole batch
    tot = dict(n_emit=0, n_cell_face=0, n_scatter=0, n_peel=0, n_draws=0)
    for k, L in enumerate(Ls):
        L.photon_id_base = k * P
        a = g.run(L)
        for key in tot:
            tot[key] += a["stats"][key]
        np.testing.assert_array_equal(b["det"][k][2], a["det"][2])
        scale = np.abs(a["det"][0]).max()
        # the batched and the plain kernel are two instantiations: FMA contraction may differ by an ulp, which the
        # cos of a grazing surface normal (:4609-4634, a cancelling sum) amplifies to ~1e-8 of a faint limb pixel
        np.testing.assert_allclose(b["det"][k][0], a["det"][0], rtol=1e-9, atol=1e-8 * scale)
        np.testing.assert_allclose(b["det"][k][1], a["det"][1], rtol=1e-9, atol=1e-8 * scale * scale)
        np.testing.assert_allclose(b["flux"][k], a["flux"], rtol=1e-12)
    for key in tot:
        assert b["stats"][key] == tot[key], key
    assert b["stats"]["n_emit"] == len(phis) * P


def test_batched_phase_curve_vs_oracle(atmospheres, oracle_factory, gpu_factory):
    """Phase-curve points out of ONE batched launch against the oracle replaying the same Philox stream launch by launch
    (the reference's loop :215-245 with limb-biased emission from 170 deg on): Stokes sums, counts and their errors."""
    atm = atmospheres("c2_hg_deck")
    o, _ = oracle_factory(atm)
    g, _ = gpu_factory(atm)
    xm = 1.3 * atm.rfront[-1]
    P = 30000
    kw = dict(mode=abi.MODE_FAST, x_max=xm, y_max=xm, seed=77, nx=1, ny=1)
    phis = [1e-5, 30.0, 90.0, 150.0, 172.5, 180.0 - 1e-5]
    Ls = _phase_launches(len(phis), P, phis, **kw)
    b = g.run_batch(Ls)
    assert g.last_engine() == 2 and b["stats"]["reserved"] == 1
    tot_cf = 0
    for k, L in enumerate(Ls):
        L.photon_id_base = k * P
        a = o.run(L)
        tot_cf += a["stats"]["n_cell_face"]
        assert np.abs(a["det"][2] - b["det"][k][2]).sum() <= max(4, 2e-5 * a["det"][2].sum())
        scale = np.abs(a["det"][0]).max()
        np.testing.assert_allclose(b["det"][k][0].sum(axis=(1, 2)), a["det"][0].sum(axis=(1, 2)), rtol=5e-5, atol=1e-7 * scale)
        np.testing.assert_allclose(b["det"][k][1].sum(axis=(1, 2)), a["det"][1].sum(axis=(1, 2)), rtol=5e-4, atol=1e-9 * scale * scale)
    assert abs(tot_cf - b["stats"]["n_cell_face"]) <= max(2, 2e-5 * tot_cf)
    I = b["det"][:, 0, 0, 0, 0]
    assert I[0] > I[2] > I[-1] > 0                                     # full phase brighter than quadrature brighter than new phase


def test_batched_launches_thermal_and_fallbacks(atmospheres, gpu_factory):
    """Batch of thermal-source launches (per-launch emitted / emergent flux), and the sequential fall-back of the
    faithful mode with the same contract."""
    from artes_b200.lib import ArtesGpuError, GpuTransport
    atm3 = atmospheres("c3_molecular")
    depth = host.cell_depth(atm3.rfront, atm3.k_sca[0], atm3.k_abs[0], atm3.nr, atm3.ntheta, atm3.nphi, 2)
    vol = host.cell_volume(atm3.rfront, atm3.thetafront(), atm3.phifront())
    cw, lum, cdf = host.thermal_tables(depth, atm3.k_abs[0], atm3.temperature, vol, atm3.wavelengths[0] * 1e-6,
                                       atm3.nr, atm3.ntheta, atm3.nphi)
    g3 = GpuTransport((0,))
    g3.set_grid(atm3.rfront, atm3.thetafront(), atm3.thetaplane(), atm3.phifront())
    g3.set_wavelength(atm3.k_sca[0], atm3.k_abs[0], atm3.uniq[0], atm3.cell_to_uniq[0], depth, cw, cdf)
    xm = 1.3 * atm3.rfront[-1]
    P = 20000
    kw = dict(mode=abi.MODE_FAST, x_max=xm, y_max=xm, seed=9, photon_source=2, photon_emission=1, nx=4, ny=4)
    Ls = _phase_launches(3, P, [10.0, 100.0, 200.0], **kw)
    b = g3.run_batch(Ls)
    assert g3.last_engine() == 2 and b["stats"]["reserved"] == 1
    for k, L in enumerate(Ls):
        L.photon_id_base = k * P
        a = g3.run(L)
        np.testing.assert_allclose(b["flux"][k], a["flux"], rtol=1e-10)
        np.testing.assert_array_equal(b["det"][k][2], a["det"][2])
        np.testing.assert_allclose(b["det"][k][0], a["det"][0], rtol=1e-9, atol=1e-12 * np.abs(a["det"][0]).max())
    # faithful mode: no batched kernel, same results through the sequential path
    kwf = dict(mode=abi.MODE_FAITHFUL, x_max=xm, y_max=xm, seed=9, nx=4, ny=4)
    Lf = _phase_launches(2, 5000, [30.0, 120.0], **kwf)
    bf = g3.run_batch(Lf)
    assert bf["stats"]["reserved"] == 2
    for k, L in enumerate(Lf):
        L.photon_id_base = k * 5000
        a = g3.run(L)
        np.testing.assert_array_equal(bf["det"][k][2], a["det"][2])
        np.testing.assert_allclose(bf["det"][k][0], a["det"][0], rtol=1e-9, atol=1e-12 * np.abs(a["det"][0]).max())
    # launches that differ in more than the detector direction are refused
    bad = _phase_launches(2, 100, [0.0, 10.0], **kwf)
    bad[1].fstop = 0.5
    with pytest.raises(ArtesGpuError):
        g3.run_batch(bad)


def _stacked_tables(atm, wls):
    """Per-wavelength compact tables of `atm` stacked for artes_gpu_set_wavelengths: one common matrix list."""
    uniq, c2u, off = [], [], 0
    for l in wls:
        uniq.append(atm.uniq[l]); c2u.append(np.asarray(atm.cell_to_uniq[l]) + off); off += atm.uniq[l].shape[0]
    depths = [host.cell_depth(atm.rfront, atm.k_sca[l], atm.k_abs[l], atm.nr, atm.ntheta, atm.nphi, 1) for l in wls]
    return (np.stack([atm.k_sca[l] for l in wls]), np.stack([atm.k_abs[l] for l in wls]), np.concatenate(uniq), np.stack(c2u), depths)


@pytest.mark.parametrize("name,extra", [("c3_molecular", dict(nx=1, ny=1)), ("c5_scale", dict(nx=16, ny=16, surface_albedo=0.3))])
def test_wavelength_batch_equals_single_launches(atmospheres, name, extra):
    """artes_gpu_set_wavelengths + run_batch over wl_index (the spectrum loop :132-165 as one kernel): launch k equals the
    single launch on the tables of wavelength k (set_wavelength) with photon_id_base = k * n_photons; a single launch
    with wl_index = k on the stacked tables gives the same again."""
    from artes_b200.lib import GpuTransport
    atm = atmospheres(name)
    wls = list(range(min(len(atm.wavelengths), 6)))
    g = GpuTransport((0,))
    g.set_grid(atm.rfront, atm.thetafront(), atm.thetaplane(), atm.phifront())
    ks, ka, uq, c2u, depths = _stacked_tables(atm, wls)
    g.set_wavelengths(ks, ka, uq, c2u, depths)
    xm = 1.3 * atm.rfront[-1]
    P = 20000
    kw = dict(mode=abi.MODE_FAST, x_max=xm, y_max=xm, seed=33, n_photons=P, det_phi=math.radians(75.0), **extra)
    Ls = [make_launch(wl_index=k, **kw) for k in range(len(wls))]
    b = g.run_batch(Ls)
    assert g.last_engine() == 2 and b["stats"]["reserved"] == 1
    singles = [g.run(make_launch(wl_index=k, photon_id_base=k * P, **kw)) for k in range(len(wls))]
    g1 = GpuTransport((0,))
    g1.set_grid(atm.rfront, atm.thetafront(), atm.thetaplane(), atm.phifront())
    assert len({round(float(r["det"][0, 0].sum()), 6) for r in singles}) > 1          # the wavelengths really differ
    for k, l in enumerate(wls):
        g1.set_wavelength(atm.k_sca[l], atm.k_abs[l], atm.uniq[l], atm.cell_to_uniq[l], depths[k])
        a = g1.run(make_launch(photon_id_base=k * P, **kw))
        for r in (singles[k], dict(det=b["det"][k], stats=None)):
            np.testing.assert_array_equal(r["det"][2], a["det"][2])
            np.testing.assert_allclose(r["det"][0], a["det"][0], rtol=1e-9, atol=1e-8 * np.abs(a["det"][0]).max())
        assert singles[k]["stats"]["n_cell_face"] == a["stats"]["n_cell_face"]
    assert b["stats"]["n_cell_face"] == sum(r["stats"]["n_cell_face"] for r in singles)
    with pytest.raises(Exception):
        g.run(make_launch(wl_index=len(wls), **kw))                                  # wavelength outside the tables


def test_batched_spectrum_vs_oracle(atmospheres, oracle_factory):
    """Spectrum points out of ONE batched launch over wl_index against the oracle run wavelength by wavelength on the
    same Philox stream (the wl_count loop :132-165)."""
    from artes_b200.lib import GpuTransport
    atm = atmospheres("c3_molecular")
    wls = [0, 11, 23, 31]
    g = GpuTransport((0,))
    g.set_grid(atm.rfront, atm.thetafront(), atm.thetaplane(), atm.phifront())
    ks, ka, uq, c2u, depths = _stacked_tables(atm, wls)
    g.set_wavelengths(ks, ka, uq, c2u, depths)
    xm = 1.3 * atm.rfront[-1]
    P = 20000
    kw = dict(mode=abi.MODE_FAST, x_max=xm, y_max=xm, seed=88, n_photons=P, nx=1, ny=1)
    b = g.run_batch([make_launch(wl_index=k, **kw) for k in range(len(wls))])
    assert g.last_engine() == 2 and b["stats"]["reserved"] == 1
    for k, l in enumerate(wls):
        o, _ = oracle_factory(atm, l)
        a = o.run(make_launch(photon_id_base=k * P, **kw))
        assert np.abs(a["det"][2] - b["det"][k][2]).sum() <= max(4, 2e-5 * a["det"][2].sum())
        scale = np.abs(a["det"][0]).max()
        np.testing.assert_allclose(b["det"][k][0].sum(axis=(1, 2)), a["det"][0].sum(axis=(1, 2)), rtol=5e-5, atol=1e-7 * scale)
        np.testing.assert_allclose(b["det"][k][1].sum(axis=(1, 2)), a["det"][1].sum(axis=(1, 2)), rtol=5e-4, atol=1e-9 * scale * scale)


def test_wavelength_batch_thermal_source(atmospheres):
    """Thermal-emission spectrum as one batched launch: per-wavelength emissivity CDF, cell weights and cell_depth."""
    from artes_b200.lib import GpuTransport
    atm = atmospheres("c3_molecular")
    wls = [0, 7, 19, 31]
    vol = host.cell_volume(atm.rfront, atm.thetafront(), atm.phifront())
    depths, cws, cdfs = [], [], []
    for l in wls:
        d = host.cell_depth(atm.rfront, atm.k_sca[l], atm.k_abs[l], atm.nr, atm.ntheta, atm.nphi, 2)
        cw, _, cdf = host.thermal_tables(d, atm.k_abs[l], atm.temperature, vol, atm.wavelengths[l] * 1e-6, atm.nr, atm.ntheta, atm.nphi)
        depths.append(d); cws.append(np.ravel(cw)); cdfs.append(np.ravel(cdf))
    uq, c2u, off = [], [], 0
    for l in wls:
        uq.append(atm.uniq[l]); c2u.append(np.asarray(atm.cell_to_uniq[l]) + off); off += atm.uniq[l].shape[0]
    g = GpuTransport((0,))
    g.set_grid(atm.rfront, atm.thetafront(), atm.thetaplane(), atm.phifront())
    g.set_wavelengths(np.stack([atm.k_sca[l] for l in wls]), np.stack([atm.k_abs[l] for l in wls]), np.concatenate(uq), np.stack(c2u),
                      depths, np.stack(cws), np.stack(cdfs))
    xm = 1.3 * atm.rfront[-1]
    P = 20000
    kw = dict(mode=abi.MODE_FAST, x_max=xm, y_max=xm, seed=41, n_photons=P, photon_source=2, nx=4, ny=4)
    b = g.run_batch([make_launch(wl_index=k, **kw) for k in range(len(wls))])
    assert g.last_engine() == 2 and b["stats"]["reserved"] == 1
    g1 = GpuTransport((0,))
    g1.set_grid(atm.rfront, atm.thetafront(), atm.thetaplane(), atm.phifront())
    for k, l in enumerate(wls):
        g1.set_wavelength(atm.k_sca[l], atm.k_abs[l], atm.uniq[l], atm.cell_to_uniq[l], depths[k], cws[k], cdfs[k])
        a = g1.run(make_launch(photon_id_base=k * P, **kw))
        s1 = g.run(make_launch(wl_index=k, photon_id_base=k * P, **kw))
        for det, flux in ((b["det"][k], b["flux"][k]), (s1["det"], s1["flux"])):
            np.testing.assert_allclose(flux, a["flux"], rtol=1e-10)
            np.testing.assert_array_equal(det[2], a["det"][2])
            np.testing.assert_allclose(det[0], a["det"][0], rtol=1e-9, atol=1e-12 * np.abs(a["det"][0]).max())
    assert len({round(float(x[0]), 3) for x in b["flux"]}) > 1            # the wavelengths emit differently


def test_two_device_context_equals_one_device(atmospheres, gpu_factory):
    """One process, two GPUs in one context (`gpu:devices=2` of the driver; ncclCommInitAll + all-reduce inside the call):
    same photon ids, so the images equal the single-device ones up to the order of the sums -- plain and batched."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from artes_b200.lib import GpuTransport
    atm = atmospheres("c4_mie_patches")
    g1, depth = gpu_factory(atm)
    g2 = GpuTransport((0, 1))
    g2.set_grid(atm.rfront, atm.thetafront(), atm.thetaplane(), atm.phifront())
    g2.set_wavelength(atm.k_sca[0], atm.k_abs[0], atm.uniq[0], atm.cell_to_uniq[0], depth)
    xm = 1.3 * atm.rfront[-1]
    kw = dict(mode=abi.MODE_FAST, x_max=xm, y_max=xm, seed=51, nx=16, ny=16)
    L = make_launch(n_photons=60001, det_phi=math.radians(50.0), **kw)
    a, b = g1.run(L), g2.run(L)
    assert b["stats"]["reserved"] == 2                                   # one kernel per device
    np.testing.assert_array_equal(b["det"][2], a["det"][2])
    np.testing.assert_allclose(b["det"][0], a["det"][0], rtol=1e-9, atol=1e-12 * np.abs(a["det"][0]).max())
    for k in ("n_emit", "n_cell_face", "n_scatter", "n_peel", "n_draws"):
        assert a["stats"][k] == b["stats"][k], k
    Ls = _phase_launches(5, 20001, [0.0, 45.0, 90.0, 135.0, 180.0], **kw)
    ab, bb = g1.run_batch(Ls), g2.run_batch(Ls)
    assert bb["stats"]["reserved"] == 2
    np.testing.assert_array_equal(bb["det"][:, 2], ab["det"][:, 2])
    np.testing.assert_allclose(bb["det"][:, 0], ab["det"][:, 0], rtol=1e-9, atol=1e-12 * np.abs(ab["det"][:, 0]).max())
    np.testing.assert_allclose(bb["flux"], ab["flux"], rtol=1e-12)
    g2.close()


def test_errors_are_reported_not_fatal(atmospheres):
    from artes_b200.lib import ArtesGpuError, GpuTransport
    atm = atmospheres("c1_template_rayleigh")
    g = GpuTransport((0,))
    xm = 1.3 * atm.rfront[-1]
    with pytest.raises(ArtesGpuError):
        g.run(make_launch(n_photons=10, x_max=xm, y_max=xm))              # no grid yet
    g.set_grid(atm.rfront, atm.thetafront(), atm.thetaplane(), atm.phifront())
    g.set_wavelength(atm.k_sca[0], atm.k_abs[0], atm.uniq[0], atm.cell_to_uniq[0], 0)
    with pytest.raises(ArtesGpuError):
        g.run(make_launch(n_photons=10, x_max=xm, y_max=xm, photon_source=2))   # thermal tables missing
    bad = make_launch(n_photons=10, x_max=xm, y_max=xm)
    bad.struct_size = 8
    with pytest.raises(ArtesGpuError):
        g.run(bad)


def test_full_size_invariants_c4(atmospheres, gpu_factory):
    """BASELINE config C4 at its full image size and 2e6 photons: size-independent properties."""
    atm = atmospheres("c4_mie_patches")
    t = host.Transport(atm, host.Params(nx=64, ny=64, det_phi=math.radians(60.0)), mode=abi.MODE_FAST)
    det, phot, res = t.radiative_transfer(2_000_000, seed=4)
    st = res["stats"]
    assert st["n_emit"] == 2_000_000 and st["n_error"] == 0
    assert st["n_peel"] == st["n_scatter"]                      # star source, black surface: one peel per scattering
    assert abs(det[0, 0].sum() - phot[0]) <= 1e-12 * abs(phot[0])   # image sum = photometry(1)
    np.testing.assert_array_equal(det[2, 1], det[2, 2]); np.testing.assert_array_equal(det[2, 1], det[2, 3])
    assert (det[2, 0] >= det[2, 1]).all() and det[2, 0].sum() <= st["n_peel"]
    assert (det[0, 0] >= 0).all() and (np.abs(det[0, 1]) <= det[0, 0] + 1e-30).all()
    # the planet is symmetric about the equator and the detector sits in the equatorial plane:
    top, bot = det[0, 0][32:].sum(), det[0, 0][:32].sum()
    assert abs(top / bot - 1.0) < 0.02
    t.close()
