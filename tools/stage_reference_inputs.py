"""Pin the host layer + oracle + CUDA path on the reference's OWN shipped input sets (the inputs that pin this path).

Run where /root/reference and oracle/_ref exist (python tools/stage_reference_inputs.py):
  * runs the unmodified reference (oracle/_ref/ref_dump) on each set and commits only its OUTPUTS as
    tests/golden/reference_inputs/<set>.npz (FP64 states at selected steps, every hydrograph value, watershed record);
  * stages the input files, untouched, under tests/_ref_inputs/<set>/ -- git-ignored (the reference's files are not part of this
    repository's history) but not gpurun-ignored, so that the GPU box, which has no /root/reference, can run the same inputs.
tests/test_reference_inputs.py reads the inputs from tests/_ref_inputs/ or /root/reference, whichever exists.
"""
import glob
import os
import shutil
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H  # noqa: E402

REF_INPUTS = "/root/reference/benchmarking/_InputFiles"
# set -> number of steps pinned (config 1 of BASELINE.json is the 10-year daily GR4J run)
SETS = {"Salmon_GR4J": 3653, "Salmon_HMETS": 3653, "Salmon_HBV": 3653, "Salmon_MOHYSE": 3653, "Irondequoit": 366,
        "York_nongridded_m_daily_i_daily": 1461, "LOTW": 2192, "Nith": 730,
        # York_nc holds several set-ups side by side ("dir__basename"): hourly model steps and / or hourly station inputs (ASCII station files; the gridded
        # variants of the same folder need NetCDF)
        "York_nc__York_nongridded_m_subdaily_i_subdaily": 35064, "York_nc__York_nongridded_m_subdaily_i_daily": 35064,
        "York_nc__York_nongridded_m_daily_i_subdaily": 1461, "York_nc__York": 35064}
# York (7 subbasins, ROUTE_HYDROLOGIC on surveyed :ChannelProfile sections, LAI interception, depression abstraction, a user-specified
# :BasinInflowHydrograph into subbasin 6) is read as shipped.
# LOTW (Lake of the Woods / Rainy River: 374 HRUs in 73 subbasins, 32 lake-type reservoirs behind weirs with linked lake HRUs, three
# :ReservoirExtraction series, two :BasinInflowHydrograph series, ROUTE_HYDROLOGIC, 45-gauge INTERP_FROM_FILE weights; the reference's own
# 151,912 HRU-steps/s benchmarking case) is read as shipped; its npz also carries the reservoir stages.
# Nith (32 HRUs / 4 subbasins, three climate stations with INTERP_NEAREST_NEIGHBOR in UTM coordinates, ROUTE_HYDROLOGIC on surveyed sections
# read through a :RedirectToFile inside the .rvp, :LakeStorage) is read as shipped.
DROP_LINES = {}


def main():
    assert H.have_reference(), "oracle/_ref/ref_dump missing"
    out_dir = os.path.join(H.GOLDEN, "reference_inputs")
    stage = os.path.join(ROOT, "tests", "_ref_inputs")
    os.makedirs(out_dir, exist_ok=True)
    only = sys.argv[1:]
    for name, n in SETS.items():
        folder, _, stem = name.partition("__")
        src = os.path.join(REF_INPUTS, folder)
        dst = os.path.join(stage, folder)
        if only and name not in only:
            continue
        shutil.rmtree(dst, ignore_errors=True)
        shutil.copytree(src, dst, ignore=shutil.ignore_patterns("*.nc"))
        base = os.path.join(dst, stem) if stem else glob.glob(os.path.join(dst, "*.rvi"))[0][:-4]
        if name in DROP_LINES:
            lines = open(base + ".rvt").read().split("\n")
            kept = [l for l in lines if DROP_LINES[name] not in l]
            assert len(kept) == len(lines) - 1
            open(base + ".rvt", "w").write("\n".join(kept))
        with tempfile.TemporaryDirectory() as td:
            ref = H.reference_run(base, os.path.join(td, "out"), n)
        rec = sorted(set([1, 2, 3, 10, 100, 365, 366, 1000, 2000, 3000, n]) & set(range(1, n + 1)))
        extra = {}
        if np.any(ref["res"][:, :, 3] != 0.0):   # reservoirs: stage of every basin that has one, at the recorded steps
            extra["res_stage"] = ref["res"][rec][:, :, 0]
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), state_steps=np.array(rec), state=ref["state"][rec],
                            qout=ref["basin"][1:n + 1, :, 0], wshed=ref["wshed"][1:n + 1],
                            sv_type=np.array([t for t, _ in ref["sv"]]), sv_layer=np.array([l for _, l in ref["sv"]]), **extra)
        print("pinned %-14s nHRU=%d NS=%d NB=%d steps=%d" % (name, ref["nHRU"], ref["NS"], ref["NB"], n))


if __name__ == "__main__":
    main()
