"""CPU test of the drop-in host classes of the angle-resolved single-layer mechanisms (viennaemc_b200/host/include/
ScatterMechanisms: 2-D charged impurities, interface roughness, remote surface-optical phonons, screened intravalley optical
phonons, Froehlich, piezoelectric): their scattering rates -- the numbers the rate tables are built from -- and the device
sampler descriptors they hand to the C ABI equal the oracle's bit for bit.  The oracle itself is pinned against the reference's
classes by tests/test_oracle_sl.py (recorder cases mos2_kaasbjerg*, mos2_pilotto_screened)."""
import os
import subprocess

import numpy as np

from oracle import pyoracle as po

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def oracle_model():
    m = po.Model(48, 0.48, 300.0, 1.0, 1.0)
    m.set_electron2d(4)
    m.add_valley(po.VALLEY_PARABOLIC_ISO_SL, 0.48, 1)
    m.add_valley(po.VALLEY_NONPARABOLIC_ISO_SL, 0.47, 6, 0.94)
    qs, rho = 2.5e8, 3.1e-6
    expected = []
    for v in range(2):
        m.add_charged_impurity_sl(v, 0, 1e16, 4.0, qs, 4.0e-9, 1.0e-9, 2.0)
        expected.append((10, (1.0e-9, 4.0e-9, qs)))
        m.add_charged_impurity_sl(v, 0, 2e15, 1.0, 0.0)
        expected.append((10, (0.0, 0.0, 0.0)))
        m.add_surface_roughness_sl(v, 0, 3e8, 3.0e-10, 1.5e-9, qs)
        expected.append((11, (0.0, 1.5e-9 * 1.5e-9, qs)))
        m.add_remote_so_sl(False, v, 0, 0.0484, 0.06, 5.0e-10, qs)
        expected.append((12, (0.0484, 5.0e-10, qs)))
        m.add_remote_so_sl(True, v, 0, 0.0484, 0.06, 5.0e-10, 0.0)
        expected.append((12, (-0.0484, 5.0e-10, 0.0)))
        m.add_screened_optical_sl(False, v, 0, 5.8e10, rho, 0.048, qs)
        expected.append((13, (0.048, 0.0, qs)))
        m.add_screened_optical_sl(True, v, 0, 5.8e10, rho, 0.048, qs)
        expected.append((13, (-0.048, 0.0, qs)))
        m.add_froehlich_sl(False, v, 0, 0.048, 0.4e-10, 5.41e-10, qs)
        expected.append((8, (0.048, 5.41e-10, qs)))
        m.add_froehlich_sl(True, v, 0, 0.048, 0.4e-10, 5.41e-10, 0.0)
        expected.append((8, (-0.048, 5.41e-10, 0.0)))
        m.add_piezo_sl(v, 0, 3.0e-11, 5.41e-10, rho, 4.2e3, qs)
        expected.append((9, (0.0, 5.41e-10, qs)))
    return m, expected


def test_host_rates_and_sampler_descriptors_equal_the_oracle_bit_for_bit(tmp_path):
    exe = tmp_path / "host_sl_mechanisms"
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-I", os.path.join(ROOT, "viennaemc_b200", "host", "include"),
                           "-I", os.path.join(ROOT, "include"), "-o", str(exe), os.path.join(ROOT, "tests", "host_sl_mechanisms.cpp")])
    lines = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines()
    m, expected = oracle_model()
    ref = m.raw_rates()
    assert len(lines) == len(expected) * 49 and ref.shape == (len(expected), 48)
    names = []
    for g, (sampler, param) in enumerate(expected):
        head = lines[49 * g].split()
        names.append(head[1])
        assert int(head[3]) == g // 10 and int(head[5]) == sampler and int(head[7]) == g // 10, head
        assert tuple(float(x) for x in head[9:12]) == param, head
        ours = np.array([float(x) for x in lines[49 * g + 1: 49 * g + 49]])
        assert np.array_equal(ours, ref[g]), (head[1], np.max(np.abs(ours / np.where(ref[g] != 0, ref[g], 1) - 1)))
        assert np.all(ours >= 0) and ours.max() > 0
    assert names[:10] == ["ChargedImpurity2D", "ChargedImpurity2D", "SurfaceRoughness", "RemoteSOAb", "RemoteSOEm", "ScreenedIntraOpticalAb",
                          "ScreenedIntraOpticalEm", "froehlichAbsorptionSL", "froehlichEmissionSL", "PiezoelectricSLTA"]
    # emission has a threshold at the phonon energy (levels of 10 meV: the first four are below 48.4 / 48 meV)
    for g in (4, 6, 8, 14, 16, 18):
        assert np.all(ref[g][:4] == 0) and np.all(ref[g][5:] > 0)
