"""The C++ host driver bin/ARTES (src/host): same CLI / artes.in grammar / input tree / output files as the
reference's `program artes` (src/ARTES.f90:4232-4517, :3472-3772).  CPU part: argument handling, keyword
grammar, atmosphere.fits ingest (dry run, no GPU touched).  GPU part: a real run compared with the Python host
mirror driving the same library."""
import math
import os
import subprocess

import numpy as np
import pytest

from artes_b200 import abi, fitsio, host
from tools import atmospheres as A
from tools.make_input import write_input

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "bin", "ARTES")


@pytest.fixture(scope="module")
def driver():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "artes_b200", "csrc"), "-j4"], stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "src", "host")], stdout=subprocess.DEVNULL)
    assert os.path.exists(BIN)
    return BIN


def run(driver, cwd, *args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([driver, *args], cwd=cwd, env=e, capture_output=True, text=True, timeout=600)


def test_usage_and_missing_input(driver, tmp_path):
    r = run(driver, tmp_path)                         # :4242-4246: fewer than two arguments -> usage, exit(0)
    assert r.returncode == 0 and "./bin/ARTES [inputDirectory] [photons] -o [outputDirectory] -k [keyWord]=[value]" in r.stdout
    r = run(driver, tmp_path, "nothing", "1e3", "-o", "x")
    assert r.returncode == 0 and "Input file does not exist!" in r.stdout     # :374-378


def test_keyword_grammar_and_overrides(driver, tmp_path, atmospheres):
    atm = atmospheres("c1_template_rayleigh")
    write_input(atm, "tmpl", root=str(tmp_path))
    dry = {"ARTES_DRYRUN": "1"}
    r = run(driver, tmp_path, "tmpl", "1e6", "-o", "o1", env=dry)
    assert r.returncode == 0, r.stderr
    assert "photons=1000000" in r.stdout and "grid nr=2 ntheta=7 nphi=1 nlambda=1" in r.stdout
    assert "type=imaging_mono" in r.stdout and "pixels=25" in r.stdout and "fstop=1e-05" in r.stdout     # 1d-5 (Fortran exponent)
    # -k overrides artes.in (:4297-4303); detector angles are degrees in the file, radians inside (:4467-4474)
    r = run(driver, tmp_path, "tmpl", "2.5e4", "-o", "o1", "-k", "detector:phi=45", "-k", "detector:pixel=64", "-k", "gpu:seed=7", env=dry)
    assert "photons=25000" in r.stdout and "pixels=64" in r.stdout and "seed=7" in r.stdout
    assert f"phi={math.radians(45.0):.6f}" in r.stdout
    # phase curves force a 1x1 detector in the equatorial plane (:455-463)
    r = run(driver, tmp_path, "tmpl", "1e4", "-o", "o1", "-k", "detector:type=phase", env=dry)
    assert "pixels=1" in r.stdout
    # unknown keyword: message and exit(0) like the reference (:4494-4496)
    r = run(driver, tmp_path, "tmpl", "1e4", "-o", "o1", "-k", "photon:colour=blue", env=dry)
    assert r.returncode == 0 and "Wrong keyword found in input file" in r.stdout
    # comment lines (* - =) and blank lines are skipped (:390)
    with open(tmp_path / "input" / "tmpl" / "artes.in", "a") as f:
        f.write("\n* a comment\n----\n====\nplanet:oblateness=0.05\n")
    r = run(driver, tmp_path, "tmpl", "1e4", "-o", "o1", env=dry)
    assert "oblateness=0.05" in r.stdout


def test_driver_fails_loudly_without_gpu(driver, tmp_path, atmospheres):
    import ctypes
    try:
        has_gpu = ctypes.CDLL("libcuda.so.1").cuInit(0) == 0
    except OSError:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    write_input(atmospheres("c1_template_rayleigh"), "tmpl", root=str(tmp_path))
    r = run(driver, tmp_path, "tmpl", "1e3", "-o", "o1")
    assert r.returncode != 0 and "no CPU fallback" in r.stderr


def _read_table(path):
    rows = [l.split() for l in open(path) if l.strip() and not l.lstrip().startswith("#")]
    return np.array(rows, dtype=float)


@pytest.mark.gpu
def test_driver_imaging_mono_matches_python_host(driver, tmp_path, atmospheres):
    atm = atmospheres("c1_template_rayleigh")
    write_input(atm, "c1", root=str(tmp_path))
    n = 200000
    r = run(driver, tmp_path, "c1", str(n), "-o", "run1", "-k", "detector:phi=60", "-k", "gpu:seed=5")
    assert r.returncode == 0, r.stdout + r.stderr
    out = tmp_path / "output" / "run1"
    for f in ("input/artes.in", "input/atmosphere.fits", "plot.dat", "error.log", "output/stokes.fits", "output/error.fits",
              "output/photometry.dat", "output/normalization.dat", "output/cell_depth.dat"):
        assert (out / f).exists(), f
    assert "detector:phi=60" in open(out / "input" / "artes.in").read()      # -k appended to the copied artes.in
    stokes = fitsio.read_hdus(str(out / "output" / "stokes.fits"))[0][1]
    err = fitsio.read_hdus(str(out / "output" / "error.fits"))[0][1]
    assert stokes.shape == (4, 25, 25) and err.shape == (5, 25, 25)
    p = host.Params(det_phi=math.radians(60.0))
    t = host.Transport(atm, p, mode=abi.MODE_FAST)
    det, phot, _ = t.radiative_transfer(n, seed=5)
    x_fov = 2.0 * math.atan(t.x_max / p.distance_planet) * 3600.0 * 180.0 / math.pi * 1000.0
    img = det[0] * 1e-6 / (x_fov / 25) ** 2
    np.testing.assert_allclose(stokes, img, rtol=1e-9, atol=1e-12 * np.abs(img).max())
    # (single-deposit pixels: the variance is a rounding-level difference of two equal numbers)
    np.testing.assert_allclose(err, host.stokes_error(det), rtol=1e-6, atol=1e-6 * np.abs(err).max())
    tab = _read_table(out / "output" / "photometry.dat")
    np.testing.assert_allclose(tab[0, 0], atm.wavelengths[0], rtol=1e-12)
    np.testing.assert_allclose(tab[0, 1:9], 1e-6 * phot[:8], rtol=1e-9, atol=1e-30)
    # ... and against the ORACLE's restatement of photon_package (:2509-2539), the reduction / photometry tail (:957-1004)
    # and the error planes (:3481-3519), fed with the raw GPU sums of the same launch
    import oracle_lib
    raw = t.gpu.run(t.launch_struct(n, seed=5))["det"]
    e_ref = oracle_lib.package_energy(p, atm.rfront, atm.wavelengths[0] * 1e-6, n)
    det_ref, phot_ref = oracle_lib.finish_detector(raw, e_ref)
    np.testing.assert_allclose(stokes, det_ref[0] * 1e-6 / (x_fov / 25) ** 2, rtol=1e-9, atol=1e-12 * np.abs(img).max())
    np.testing.assert_allclose(err, oracle_lib.stokes_error(det_ref), rtol=1e-6, atol=1e-6 * np.abs(err).max())
    np.testing.assert_allclose(tab[0, 1:9], 1e-6 * phot_ref[:8], rtol=1e-9, atol=1e-30)
    t.close()


@pytest.mark.gpu
def test_driver_phase_curve_and_spectrum(driver, tmp_path, atmospheres):
    write_input(atmospheres("c2_hg_deck"), "c2", root=str(tmp_path))
    r = run(driver, tmp_path, "c2", "20000", "-o", "ph", "-k", "detector:type=phase")
    assert r.returncode == 0, r.stdout + r.stderr
    ph = _read_table(tmp_path / "output" / "ph" / "output" / "phase.dat")
    assert ph.shape == (73, 9) and ph[0, 0] == 0.0 and ph[-1, 0] == 180.0 and abs(ph[1, 0] - 2.5) < 1e-9     # :215-245
    assert (ph[:, 1] > 0).all() and ph[0, 1] > 5 * ph[-1, 1]          # bright at full phase, faint near new phase
    assert len(_read_table(tmp_path / "output" / "ph" / "output" / "normalization.dat")) == 1                # only at phase 0 (:3631)
    # default: ONE walk per packet observed from all azimuths (artes_gpu_run_multi); gpu:phase_walks=independent repeats the walk
    # per azimuth like the reference (one batched launch).  Same curve within the photon noise; the single-walk curve equals the
    # library call made directly (angles below 170 deg: photon ids [0, packages)).
    assert "one walk per packet" in r.stdout
    r2 = run(driver, tmp_path, "c2", "20000", "-o", "ph2", "-k", "detector:type=phase", "-k", "gpu:phase_walks=independent")
    assert r2.returncode == 0 and "one batched launch" in r2.stdout
    ph2 = _read_table(tmp_path / "output" / "ph2" / "output" / "phase.dat")
    assert ph2.shape == (73, 9)
    np.testing.assert_allclose(ph[:56, 1], ph2[:56, 1], rtol=0.12)
    atm2 = atmospheres("c2_hg_deck")
    t = host.Transport(atm2, host.Params(nx=1, ny=1, phase_curve=True), mode=abi.MODE_FAST)
    t.set_wavelength(0)
    phis = [1.e-5 * math.pi / 180.0, 2.5 * math.pi / 180.0]
    while len(phis) < 72:
        phis.append(phis[-1] + 2.5 * math.pi / 180.0)
    for k in (10, 40, 67):       # a detector's image depends on the walk (ids [0, packages), seed 1) and its own direction only
        m = t.gpu.run_multi([t.launch_struct(20000, seed=1, det_phi=phis[k])])
        e_k = host.package_energy(t.p, atm2.rfront, atm2.wavelengths[0] * 1e-6, 20000)
        np.testing.assert_allclose(ph[k, 1], 1e-6 * e_k * m["det"][0][0, 0, 0, 0], rtol=1e-7)
        assert abs(ph[k, 0] - 2.5 * k) < 1e-9
    t.close()
    atm3 = A.c3_molecular(nr=30, nl=4)
    write_input(atm3, "c3", root=str(tmp_path))
    r = run(driver, tmp_path, "c3", "20000", "-o", "sp", "-k", "detector:type=spectrum")
    assert r.returncode == 0, r.stdout + r.stderr
    sp = _read_table(tmp_path / "output" / "sp" / "output" / "spectrum.dat")
    assert sp.shape == (4, 5)
    np.testing.assert_allclose(sp[:, 0], atm3.wavelengths, rtol=1e-12)
    assert (sp[:, 1] > 0).all()
    assert _read_table(tmp_path / "output" / "sp" / "output" / "optical_depth.dat").shape == (4, 4)


@pytest.mark.gpu
def test_driver_broadband_image_is_the_batched_wavelength_sum(driver, tmp_path):
    """imaging_broad (:167-204) through the driver: all wavelengths in one batched launch, the image is the sum over the
    wavelengths scaled with the LAST wavelength's package energy -- compared with the same sum made launch by launch."""
    atm3 = A.c3_molecular(nr=30, nl=4)
    write_input(atm3, "c3b", root=str(tmp_path))
    # the detector:type flags are only ever SET (:4457-4466) and `spectrum` is tested first (:132), so the template's
    # spectrum line has to go from artes.in itself; a -k override would leave both flags on and run the spectrum
    ain = tmp_path / "input" / "c3b" / "artes.in"
    ain.write_text(ain.read_text().replace("detector:type=spectrum", "detector:type=imaging_broad"))
    assert "detector:type=imaging_broad" in ain.read_text()
    n = 40000
    r = run(driver, tmp_path, "c3b", str(n), "-o", "bb", "-k", "detector:phi=60", "-k", "gpu:seed=9")
    assert r.returncode == 0, r.stdout + r.stderr
    assert "one batched launch" in r.stdout
    out = tmp_path / "output" / "bb" / "output"
    stokes = fitsio.read_hdus(str(out / "stokes.fits"))[0][1]
    assert _read_table(out / "optical_depth.dat").shape == (4, 4)
    npix = stokes.shape[-1]                                   # the spectrum template of c3 has a 1 x 1 detector
    p = host.Params(nx=npix, ny=npix, det_phi=math.radians(60.0))
    t = host.Transport(atm3, p, mode=abi.MODE_FAST)
    acc = None
    for l in range(4):
        t.set_wavelength(l)
        res = t.gpu.run(t.launch_struct(n, seed=9, photon_id_base=l * n))
        acc = res["det"].copy() if acc is None else acc + res["det"]
    energy = host.package_energy(p, atm3.rfront, atm3.wavelengths[3] * 1e-6, n, t.emis_total)
    det = host.detector_from_sums(acc, energy)
    x_fov = 2.0 * math.atan(t.x_max / p.distance_planet) * 3600.0 * 180.0 / math.pi * 1000.0
    img = det[0] * 1e-6 / (x_fov / stokes.shape[-1]) ** 2
    assert stokes.shape == img.shape
    np.testing.assert_allclose(stokes, img, rtol=1e-8, atol=1e-8 * np.abs(img).max())
    t.close()


@pytest.mark.gpu
def test_driver_spectrum_with_more_wavelengths_than_one_batch(driver, tmp_path):
    """300 wavelengths: the driver splits the wavelength loop into batches of ARTES_MAX_BATCH = 256 launches."""
    atm = A.c3_molecular(nr=12, nl=300)
    write_input(atm, "c3w", root=str(tmp_path))
    r = run(driver, tmp_path, "c3w", "3000", "-o", "sp", "-k", "gpu:seed=3")
    assert r.returncode == 0, r.stdout + r.stderr
    assert "batched launches of 256 wavelengths" in r.stdout
    sp = _read_table(tmp_path / "output" / "sp" / "output" / "spectrum.dat")
    assert sp.shape == (300, 5)
    np.testing.assert_allclose(sp[:, 0], atm.wavelengths, rtol=1e-12)
    assert (sp[:, 1] > 0).all()
    # wavelength 299 through the second batch equals the single launch with its photon-id range
    t = host.Transport(atm, host.Params(nx=1, ny=1), mode=abi.MODE_FAST)
    t.set_wavelength(299)
    det, phot, _ = t.radiative_transfer(3000, seed=3, photon_id_base=299 * 3000)
    np.testing.assert_allclose(sp[299, 1:5], 1e-6 * det[0, :, 0, 0], rtol=1e-8, atol=1e-30)
    t.close()


@pytest.mark.gpu
def test_driver_flow_on_and_flow_off_paths_walk_the_same_streams(driver, tmp_path):
    """With a flow counter on, the wavelength loop runs launch by launch (artes_gpu_run) instead of as one batched launch.
    Every call walks its own photon-id range (call k: k*packages + [0, packages)) on both paths, so the spectra are equal
    and the launches are statistically independent (the reference carries its generator state from call to call)."""
    atm = A.c3_molecular(nr=30, nl=4)
    write_input(atm, "c3f", root=str(tmp_path))
    n = 30000
    r0 = run(driver, tmp_path, "c3f", str(n), "-o", "batched", "-k", "gpu:seed=11")
    r1 = run(driver, tmp_path, "c3f", str(n), "-o", "looped", "-k", "gpu:seed=11", "-k", "output:flow_latitudinal=on")
    assert r0.returncode == 0 and r1.returncode == 0, r0.stderr + r1.stderr
    assert "one batched launch" in r0.stdout and "one batched launch" not in r1.stdout
    a = _read_table(tmp_path / "output" / "batched" / "output" / "spectrum.dat")
    b = _read_table(tmp_path / "output" / "looped" / "output" / "spectrum.dat")
    np.testing.assert_allclose(a, b, rtol=1e-8, atol=1e-30)
    # ... and launch l is the single launch with photon_id_base = l*packages, not a replay of ids 0..packages
    t = host.Transport(atm, host.Params(nx=1, ny=1), mode=abi.MODE_FAST)
    t.set_wavelength(2)
    det, _, _ = t.radiative_transfer(n, seed=11, photon_id_base=2 * n)
    np.testing.assert_allclose(b[2, 1:5], 1e-6 * det[0, :, 0, 0], rtol=1e-8, atol=1e-30)
    det0, _, _ = t.radiative_transfer(n, seed=11, photon_id_base=0)
    assert abs(1e-6 * det0[0, 0, 0, 0] - b[2, 1]) > 1e-6 * abs(b[2, 1])
    assert (tmp_path / "output" / "looped" / "output" / "flow_latitudinal.fits").exists()
    t.close()
