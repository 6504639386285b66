"""Result files: what our drivers write is what the reference writes, as the reference's own plotting
helpers read it.

tests/golden/result_files/reference: written by the UNMODIFIED reference examples (oracle/make_ref_result_files.py);
tests/golden/result_files/ours:      written by our GPU-backed drivers (tools/make_result_file_fixtures.py, GPU box).
Both sets are cut after 300 lines.  The layout checks run everywhere; the pass through the reference's readers
(helper/emcPlottingFiles/emcPlottingFiles/readResultFile.py) runs where the reference tree is mounted."""
import importlib.util
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "golden", "result_files", "reference")
OUR_DIR = os.path.join(HERE, "golden", "result_files", "ours")
READER = "/root/reference/helper/emcPlottingFiles/emcPlottingFiles/readResultFile.py"

# (our file, the reference's file of the same kind)
PAIRS = [
    ("resistorV50as1000ElectronsFinal.txt", "resistorV50as1000ElectronsFinal.txt", "particle"),
    ("resistorV50as1000ElectronsEq.txt", "resistorV50as1000ElectronsEq.txt", "particle"),
    ("bulkSimulationElectrons300.txt", "bulkSimulationElectronsEq.txt", "particle"),
    ("resistorV50as1000PotentialAvg.txt", "resistorV50as1000PotentialAvg.txt", "grid"),
    ("resistorV50as1000PotentialEq.txt", "resistorV50as1000PotentialEq.txt", "grid"),
    ("resistorV50as1000ElectronsConcAvg.txt", "resistorV50as1000ElectronsConcAvg.txt", "grid"),
    ("resistorV50as1000ElectronsConcEq.txt", "resistorV50as1000ElectronsConcEq.txt", "grid"),
    ("resistorV50as1000EFieldXEq.txt", "resistorV50as1000EFieldXEq.txt", "grid"),
    ("resistorV50as1000EFieldYEq.txt", "resistorV50as1000EFieldYEq.txt", "grid"),
    ("resistorV50as1000ElectronsCurrent.txt", "resistorV50as1000ElectronsCurrent.txt", "current"),
    ("bulkSimulationAvgEnergy.txt", "bulkSimulationAvgEnergy.txt", "avg"),
    ("bulkSimulationAvgDriftVelocity.txt", "bulkSimulationAvgDriftVelocity.txt", "avg"),
    ("bulkSimulationvalleyOccupation.txt", "bulkSimulationvalleyOccupation.txt", "avg"),
    ("Acoustic00ScatterMechanism.txt", "Acoustic00ScatterMechanism.txt", "rate"),
    ("Coulomb00ScatterMechanism.txt", "Coulomb00ScatterMechanism.txt", "rate"),
    ("ZeroInterValleyEmissionF00ScatterMechanism.txt", "ZeroInterValleyEmissionF00ScatterMechanism.txt", "rate"),
    ("FirstInterValleyAbsorptionG00ScatterMechanism.txt", "FirstInterValleyAbsorptionG00ScatterMechanism.txt", "rate"),
]

pytestmark = pytest.mark.skipif(not os.path.isdir(OUR_DIR), reason="fixtures of our drivers not generated yet")


def _rows(path):
    with open(path) as f:
        return [line.rstrip("\n").split(" ") for line in f if line.strip()]


@pytest.mark.parametrize("ours,theirs,kind", PAIRS, ids=[p[0][:-4] for p in PAIRS])
def test_same_layout_as_the_reference_files(ours, theirs, kind):
    a, b = _rows(os.path.join(OUR_DIR, ours)), _rows(os.path.join(REF_DIR, theirs))
    # header line (extent of the box / grid) and columns per line
    if kind in ("particle", "grid"):
        assert len(a[0]) == len(b[0])
        if kind == "grid":
            assert a[0] == b[0]  # same grid extent
            assert len(a) == len(b)
        assert {len(r) for r in a[1:]} == {len(r) for r in b[1:]}
    else:
        assert {len(r) for r in a} == {len(r) for r in b}
    # every field parses as a number; the integer columns are written as integers (particle index, sub-valley, valley;
    # netto particle counts of the current file)
    if kind == "particle":  # idx, pos (2 or 3), [k (3), energy, sub-valley, valley[, tau]]
        dim = len(b[0])
        int_cols = [0] + ([dim + 5, dim + 6] if len(b[1]) > dim + 1 else [])
    else:
        int_cols = [1, 2] if kind == "current" else []
    for ra, rb in zip(a[1:40], b[1:40]):
        for x, y in zip(ra, rb):
            float(x), float(y)
        for c in int_cols:
            assert ra[c].lstrip("-").isdigit() and rb[c].lstrip("-").isdigit(), (c, ra[c], rb[c])


def test_rate_files_agree_with_the_reference_to_their_six_digits():
    for ours, theirs, kind in PAIRS:
        if kind != "rate":
            continue
        a, b = np.loadtxt(os.path.join(OUR_DIR, ours)), np.loadtxt(os.path.join(REF_DIR, theirs))
        assert a.shape == b.shape and np.allclose(a, b, rtol=2e-6, atol=0)


def test_equilibrium_potential_agrees_with_the_reference():
    a = np.loadtxt(os.path.join(OUR_DIR, "resistorV50as1000PotentialEq.txt"), skiprows=1)
    b = np.loadtxt(os.path.join(REF_DIR, "resistorV50as1000PotentialEq.txt"), skiprows=1)
    assert a.shape == b.shape == (21, 101) and np.max(np.abs(a - b)) < 2e-4  # solver accuracy 1e-4 V


@pytest.mark.skipif(not os.path.exists(READER), reason="the reference tree (its plotting helpers) is not mounted")
def test_the_reference_readers_read_our_files(capsys):
    spec = importlib.util.spec_from_file_location("readResultFile", READER)
    rr = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(rr)
    for ours, theirs, kind in PAIRS:
        pa, pb = os.path.join(OUR_DIR, ours), os.path.join(REF_DIR, theirs)
        if kind == "particle":
            def read(path):
                try:
                    return rr.readParticleFile(path, return_maxPos=True)
                except SystemExit:  # the helper gives up on 3-D files without a tau column -- the reference's own included
                    return None
            ra_, rb_ = read(pa), read(pb)
            assert (ra_ is None) == (rb_ is None), (ours, "the reference's reader treats the two files differently")
            if ra_ is None:
                continue
            (da, ma), (db, mb) = ra_, rb_
            assert list(da.columns) == list(db.columns) and ma.shape == mb.shape and np.allclose(ma, mb)
            assert list(da.dtypes) == list(db.dtypes)
            assert (da["idx"].to_numpy() == np.arange(len(da))).all()
        elif kind == "grid":
            ga, gb = rr.readGridFile(pa), rr.readGridFile(pb)
            assert ga.shape == gb.shape and np.isfinite(ga).all()
        elif kind == "current":
            ca, cb = rr.readCurrentFile(pa), rr.readCurrentFile(pb)
            assert list(ca.columns) == list(cb.columns) == ["time", "netParContact0", "netParContact1", "currentContact0",
                                                            "currentContact1"]
            assert np.allclose(ca["time"], cb["time"])
        elif kind == "avg":
            aa, ab = rr.readBulkSimulationAvgFile(pa), rr.readBulkSimulationAvgFile(pb)
            assert list(aa.columns) == list(ab.columns) and np.allclose(aa.iloc[:, 0], ab.iloc[:, 0])
        else:
            ra, rb = rr.readScatterMechanismFile(pa), rr.readScatterMechanismFile(pb)
            assert list(ra.columns) == list(rb.columns) == ["energy", "rate"]
