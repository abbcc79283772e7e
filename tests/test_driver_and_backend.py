"""Batch driver (examples/decoding/quantum_surface.py mirror) and backend shim (mdopt/backend/array.py mirror)."""
import os
import pickle
import re

import numpy as np
import pytest

import mps_oracle as O
from conftest import load_golden
from mdopt_b200.drivers import quantum_surface as drv

# the file-name pattern the reference's analysis code scans a results directory with (data_handling.py:115)
PATTERN = (r"^latticesize{L}_bonddim{chi}_errorrate[0-9\.]+(?:e[+-]?[0-9]+)?_errormodel{model}+_bias_prob[0-9\.]+"
           r"(?:e[+-]?[0-9]+)?_numexperiments[0-9]+_tolerance[0-9\.]+(?:e[+-]?[0-9]+)?_cut[0-9\.]+(?:e[+-]?[0-9]+)?"
           r"_seed\d+\.pkl$")


def test_seeded_errors_are_the_reference_generator():
    assert drv.generate_errors(3, 0.05, 96, "Bitflip", 123) == O.seeded_errors(13, 0.05, 96, 123, "Bitflip")
    z = drv.generate_errors(3, 0.3, 8, "Phaseflip", 5)
    assert all(set(e) <= {"I", "Z"} for e in z)


def test_pickle_is_what_the_reference_analysis_consumes(tmp_path):
    """data_handling.process_failure_statistics (data_handling.py:56-175): file-name regex, data['failures'],
    data['error_rate'], nan-aware mean and s.e.m."""
    from scipy.stats import sem

    data = {"logicals_distributions": [np.array([1.0, 0, 0, 0]), np.nan], "failures": [0, 1, float("nan"), 0],
            "lattice_size": 3, "chi_max": 64, "error_rate": 0.05, "bias_prob": 0.1, "error_model": "Bitflip",
            "seed": 123, "tolerance": 1e-08, "cut": 1e-17}
    path = drv.save_experiment_data(data, 3, 64, 0.05, "Bitflip", 0.1, 4, 123, 1e-08, 1e-17, str(tmp_path))
    assert re.match(PATTERN.format(L=3, chi=64, model="Bitflip"), os.path.basename(path))
    with open(path, "rb") as fh:
        back = pickle.load(fh)
    assert set(back) >= {"logicals_distributions", "failures", "lattice_size", "chi_max", "error_rate", "bias_prob",
                         "error_model", "seed"}
    fails = np.array(back["failures"])
    assert round(back["error_rate"], 5) == 0.05
    assert np.nanmean(fails) == pytest.approx(1 / 3)
    assert np.isfinite(sem(fails, nan_policy="omit"))


def test_backend_shim_surface_without_gpu():
    from mdopt_b200.backend import array as xp

    assert xp.GPU is True and xp.is_cuda_backend() and xp.backend_name() == "mdopt_b200"
    for name in ("to_device", "to_host", "stream", "synchronize", "einsum", "svd", "asfortran", "linalg"):
        assert hasattr(xp, name)
    assert xp.zeros is not None            # forwarded to the device array library
    with pytest.raises(AttributeError, match="has no attribute 'no_such_function'"):
        xp.no_such_function
    np.testing.assert_array_equal(xp.to_host([1, 2]), [1, 2])


@pytest.mark.gpu
def test_backend_shim_on_device():
    from mdopt_b200.backend import array as xp

    rng = np.random.default_rng(0)
    a, b = rng.standard_normal((5, 7)), rng.standard_normal((7, 3))
    with xp.stream():
        c = xp.einsum("ij,jk->ik", a, b)
    xp.synchronize()
    np.testing.assert_allclose(xp.to_host(c), a @ b, rtol=1e-13)
    u, s, vh = xp.linalg.svd(xp.to_device(a))
    np.testing.assert_allclose(s, np.linalg.svd(a, compute_uv=False), rtol=1e-10)
    np.testing.assert_allclose((u * s) @ vh, a, atol=1e-12)
    f = xp.asfortran(a)
    assert f.stride() == (1, 5) and np.array_equal(xp.to_host(f), a)


@pytest.mark.gpu
def test_driver_end_to_end_matches_the_reference_fixture(tmp_path):
    g = load_golden("d3_chi64_bitflip")   # reference decodes of seeded_errors(13, 0.05, 96, 123)
    drv.main(["--lattice_size", "3", "--bond_dim", "64", "--error_rate", "0.05", "--bias_prob", "0.1",
              "--num_experiments", "96", "--error_model", "Bitflip", "--seed", "123", "--num_processes", "1",
              "--silent", "True", "--tolerance", "1e-12", "--cut", "1e-17", "--output_dir", str(tmp_path)])
    files = os.listdir(tmp_path)
    assert len(files) == 1 and re.match(PATTERN.format(L=3, chi=64, model="Bitflip"), files[0])
    with open(os.path.join(tmp_path, files[0]), "rb") as fh:
        data = pickle.load(fh)
    assert data["failures"] == [1 - r["success"] for r in g["results"]]
    for row, r in zip(data["logicals_distributions"], g["results"]):
        np.testing.assert_allclose(row, r["dist"], rtol=1e-10, atol=1e-14)


@pytest.mark.gpu
def test_driver_passes_the_error_model_as_bias_type():
    """Phaseflip errors are decoded with bias_type='Phaseflip', i.e. the depolarising-bias branch, as in the reference
    (quantum_surface.py:157-169, decoding.py:1267-1283)."""
    errors = drv.generate_errors(3, 0.08, 12, "Phaseflip", 9)
    data = drv.run_experiment(3, 64, 0.08, 0.1, 12, "Phaseflip", 9, errors, True, 1, 1e-12, 1e-17)
    ocode = O.surface_code(3)
    for e, row, fail in zip(errors, data["logicals_distributions"], data["failures"]):
        ref, ok = O.decode_css(ocode, e, chi_max=64, cut=1e-17, bias_type="Phaseflip", bias_prob=0.1,
                               contraction_strategy="Optimised", tolerance=1e-12)
        assert fail == 1 - ok
        if not O.decode_truncates(ocode, e, chi_max=64, cut=1e-17, bias_type="Phaseflip", bias_prob=0.1,
                                  contraction_strategy="Optimised"):
            np.testing.assert_allclose(row, np.asarray(ref, float), rtol=1e-9, atol=1e-12)
