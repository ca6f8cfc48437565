"""CPU tests of the boundary: the C-ABI library loads and exports every symbol include/rvn_gpu.h declares, fails loudly
without a GPU (no CPU fallback), and the host layer's parsers reject what they do not implement."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import helpers as H
from ravenhydroframework_b200 import api, synth

ROOT = H.ROOT


def test_engine_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "rvn_gpu.h")).read()
    names = sorted(set(re.findall(r"\b(rvn_gpu_[a-z_0-9]+)\s*\(", hdr)))
    assert len(names) >= 20
    lib = api.engine_lib()
    for n in names:
        assert hasattr(lib, n), "librvn_gpu.so does not export %s" % n
    assert lib.rvn_gpu_abi_version() == 11


def test_loaded_libraries_were_built_from_the_present_sources():
    """build() stamps a digest of sources + flags into each library and rebuilds on mismatch: what is loaded is what is in the tree."""
    from ravenhydroframework_b200 import build as b
    want = b.source_digests()
    eng, host = api.engine_lib(), api.host_lib()
    eng.rvn_gpu_build_digest.restype = C.c_char_p
    host.rvh_build_digest.restype = C.c_char_p
    assert eng.rvn_gpu_build_digest().decode() == want["librvn_gpu.so"]
    assert host.rvh_build_digest().decode() == want["libraven_b200.so"]


def test_engine_is_sm100a_with_specialised_and_interpreter_sweeps():
    import shutil
    import subprocess
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    so = os.path.join(ROOT, "ravenhydroframework_b200", "librvn_gpu.so")
    out = subprocess.run(["cuobjdump", "-sass", so], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True).stdout
    assert "sm_100a" in out
    assert "hru_sweep_static" in out and "Prog_HMETS" in out, "specialised sweep of the HMETS program missing"
    assert "hru_sweep_interp" in out and "route_level_kernel" in out and "balance_kernel" in out
    assert "DFMA" in out and "DADD" in out      # FP64 arithmetic
    assert "LDG.E.64.EF" in out or "LDG.E.EF.64" in out, "state streams should use evict-first loads (__ldcs)"


def test_static_program_header_is_current():
    """csrc/rvn_static_programs.cuh must equal what the engine emits for the current templates (tools/gen_static_programs.py)."""
    import ctypes as C
    import tempfile
    eng = api.engine_lib()
    hdr = open(os.path.join(ROOT, "ravenhydroframework_b200", "csrc", "rvn_static_programs.cuh")).read()
    for name, writer in (("HMETS", synth.write_hmets), ("HBVEC", synth.write_hbv), ("GR4J", synth.write_gr4j), ("BLENDED", synth.write_blended)):
        with tempfile.TemporaryDirectory() as td:
            base = writer(td, n_sb=1, hru_per_sb=1, n_days=5, seed=1)
            m = api.RavenModel(base)
            m.flatten(1)
            buf = C.create_string_buffer(1 << 16)
            assert eng.rvn_gpu_emit_static_program(m.desc, name.encode(), buf, len(buf)) == 0
            assert buf.value.decode() in hdr, name


def test_no_cpu_fallback(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    base = synth.write_hmets(str(tmp_path), n_sb=1, hru_per_sb=1, n_days=10)
    m = api.RavenModel(base)
    with pytest.raises(api.RavenError, match="no CUDA device"):
        api.GpuSimulation(m)


def test_product_path_does_not_touch_the_oracle():
    pkg = os.path.join(ROOT, "ravenhydroframework_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cpp", ".h", ".cu", ".cuh")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "liboracle" not in txt and "raven_oracle" not in txt and "rvo_run" not in txt, f


def test_unsupported_commands_fail_loudly(tmp_path):
    base = synth.write_hmets(str(tmp_path), n_sb=1, hru_per_sb=1, n_days=10)
    rvi = open(base + ".rvi").read()
    open(base + ".rvi", "w").write(rvi.replace("INF_HMETS", "INF_UPSCALED_GREEN_AMPT"))
    with pytest.raises(api.RavenError, match="not implemented"):
        api.RavenModel(base)
    open(base + ".rvi", "w").write(rvi + ":ReservoirDemandAllocation X\n")
    with pytest.raises(api.RavenError, match="not supported"):
        api.RavenModel(base)
    open(base + ".rvi", "w").write(rvi.replace(":Method                ORDERED_SERIES", ":Method RUNGE_KUTTA_4"))
    with pytest.raises(api.RavenError, match="ORDERED_SERIES"):
        api.RavenModel(base)
    open(base + ".rvi", "w").write(rvi.replace(":Method                ORDERED_SERIES", ":Method ITER_HEUN"))   # the reference reads two numbers unchecked
    with pytest.raises(api.RavenError, match="ITER_HEUN needs"):
        api.RavenModel(base)


def test_atmospheric_estimators_are_fatal_only_when_a_consumed_pet_needs_them(tmp_path):
    """:CloudCoverMethod / :SWRadiationMethod options outside the built set: the reference derives the whole atmospheric chain in every
    run whether or not anything reads it; here it is derived only for a PET method that consumes it, and only then is an unbuilt
    estimator an error.  PET_VAPDEFICIT without PET_VAP_COEFF fails like the reference's AutoCalculateLandUseProps."""
    import re
    base = synth.write_gr4j(str(tmp_path), n_sb=1, hru_per_sb=2, n_days=10)
    rvi = open(base + ".rvi").read()
    open(base + ".rvi", "w").write(":SWRadiationMethod SW_RAD_UBCWM\n:CloudCoverMethod CLOUDCOV_UBCWM\n" + rvi)
    api.RavenModel(base).flatten(1)      # GR4J template: PET_DATA-free Hargreaves 1985 / Oudin do not read the atmospheric chain
    turc = re.sub(r":Evaporation\s+\S+", ":Evaporation PET_TURC_1961", rvi)
    open(base + ".rvi", "w").write(":SWRadiationMethod SW_RAD_UBCWM\n" + turc)
    with pytest.raises(api.RavenError, match="SW_RAD_UBCWM is not implemented"):
        api.RavenModel(base).flatten(1)
    open(base + ".rvi", "w").write(turc)
    assert api.RavenModel(base).flatten(1) is not None or True
    open(base + ".rvi", "w").write(re.sub(r":Evaporation\s+\S+", ":Evaporation PET_VAPDEFICIT", rvi))
    with pytest.raises(api.RavenError, match="PET_VAP_COEFF"):
        api.RavenModel(base).flatten(1)


def test_routing_order_heap_tree(tmp_path):
    base = synth.write_hmets(str(tmp_path), n_sb=15, hru_per_sb=2, n_days=10)
    m = api.RavenModel(base)
    m.flatten(1)
    o = np.zeros(15, dtype=np.int32)
    m.lib.rvh_desc_order.argtypes = [C.c_void_p, C.c_void_p]
    m.lib.rvh_desc_order(m.desc, o.ctypes.data)
    # deepest level first (subbasins 8..15 -> indices 7..14), ascending inside a level, outlet last
    assert list(o) == list(range(7, 15)) + list(range(3, 7)) + [1, 2] + [0]


def test_timeseries_reader_matches_reference_digit_arithmetic(tmp_path):
    """Forcing values are parsed with the reference's own decimal accumulation (src/CommonFunctions.cpp:1267-1342)."""
    base = synth.write_hmets(str(tmp_path), n_sb=1, hru_per_sb=1, n_days=5)
    txt = open(base + ".rvt").read().split("\n")
    i = [j for j, l in enumerate(txt) if ":Units" in l][0] + 1
    txt[i] = "    0.1 0.7 -17.991528 -4.024469 0.3"
    open(base + ".rvt", "w").write("\n".join(txt))
    m = api.RavenModel(base)
    m.flatten(1)
    m.lib.rvh_desc_gauge_series.restype = C.c_void_p
    m.lib.rvh_desc_gauge_series.argtypes = [C.c_void_p]
    s = np.ctypeslib.as_array((C.c_double * (9 * 5)).from_address(m.lib.rvh_desc_gauge_series(m.desc))).reshape(9, 5)

    def ref_parse(t):
        sign = -1.0 if t.startswith("-") else 1.0
        t = t.lstrip("+-")
        ip, fp = (t.split(".") + [""])[:2]
        v = 0.0
        for ch in ip:
            v = v * 10.0 + (ord(ch) - 48)
        p10 = 10.0
        for ch in fp:
            v += (ord(ch) - 48) / p10
            p10 *= 10.0
        return sign * v
    assert s[0, 0] == ref_parse("0.1") + ref_parse("0.7")            # PRECIP = RAINFALL + SNOWFALL
    assert s[4, 0] == ref_parse("-17.991528") and s[5, 0] == ref_parse("-4.024469")
