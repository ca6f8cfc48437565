"""Development / profiling driver: a few steps of one named shape so that `ncu -k regex:hru_sweep` can capture its sweep kernel.
usage: prof_shapes.py {hmets|hbv|gr4j|blended|interp} [members] [n_sb] [steps]"""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ravenhydroframework_b200 import api, synth
kind = sys.argv[1]
E = int(sys.argv[2]) if len(sys.argv) > 2 else 64
n_sb = int(sys.argv[3]) if len(sys.argv) > 3 else 500
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 16
td = tempfile.mkdtemp()
force = False
if kind == "hmets":
    base = synth.write_hmets(td, n_sb=n_sb, hru_per_sb=20, n_days=steps, seed=1, single_class=True, season_offset=105)
elif kind == "hbv":
    base = synth.write_hbv(td, n_sb=n_sb, hru_per_sb=20, n_days=steps, seed=2, routing="ROUTE_MUSKINGUM", season_offset=100)
elif kind == "gr4j":
    base = synth.write_gr4j(td, n_sb=n_sb, hru_per_sb=20, n_days=steps, seed=3, season_offset=100)
elif kind.startswith("blended"):
    hourly = "hourly" in kind
    base = synth.write_blended(td, n_sb=n_sb, hru_per_sb=20, n_days=((steps + 23) // 24 if hourly else steps), seed=4, routing="ROUTE_MUSKINGUM", catchment_route="TRIANGULAR_UH", season_offset=124)
    if hourly:
        rvi = open(base + ".rvi").read()
        open(base + ".rvi", "w").write(rvi.replace(":TimeStep              1.0", ":TimeStep              01:00:00"))
else:
    base = synth.write_hmets(td, n_sb=n_sb, hru_per_sb=20, n_days=steps, seed=1, single_class=True, season_offset=105)
    force = True
model = api.RavenModel(base)
model.flatten(E)
if "grid" in kind:
    synth.apply_cell_grid(model, base, 200, 200)
    model.flatten(E)
sim = api.GpuSimulation(model, n_members=E, force_interpreter=force)
sim.set_kernel_timing(True)
sim.run(steps)
st = sim.last_run_stats()
print(kind, sim.kernel_variant(), "E", E, "nHRU", model.nHRU, "steps", steps, "total ms/step %.4f sweep ms/step %.4f launches %d" % (st["total_ms"] / steps, st["sweep_ms"] / steps, st["launches"]))
