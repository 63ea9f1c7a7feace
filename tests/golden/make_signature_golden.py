"""Record the parameter names (in order) of the reference's kernel-backend launchers this repository replaces, so that
tests/test_signature_parity_cpu.py can hold curobo_b200/backends/* to them without the reference tree present.
Needs /root/reference (authoring container only):  python tests/golden/make_signature_golden.py"""
import inspect
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _reference_under_shim as R  # noqa: E402

R.prepare()
TARGETS = {
    "kinematics": ["launch_kinematics_forward", "launch_kinematics_forward_spheres", "launch_kinematics_backward"],
    "geometry": ["self_collision_distance"],
    "trajectory": ["launch_bspline_interpolation_forward_kernel", "launch_bspline_interpolation_single_dt_kernel",
                   "launch_bspline_interpolation_backward_kernel"],
    "optimization": ["launch_lbfgs_step", "launch_line_search"],
    "dynamics": ["launch_rnea_forward", "launch_rnea_backward"],
    "pba": ["launch_pba3d"],
}
out = {}
for mod, names in TARGETS.items():
    m = R.ref(f"curobo._src.curobolib.backends.cuda_core_backend.{mod}")
    for n in names:
        fn = getattr(m, n)
        sig = inspect.signature(fn)
        out[f"{mod}.{n}"] = {"file": os.path.relpath(inspect.getsourcefile(fn), "/root/reference"),
                            "line": inspect.getsourcelines(fn)[1],
                            "params": [p.name for p in sig.parameters.values()],
                            "defaults": {p.name: repr(p.default) for p in sig.parameters.values()
                                         if p.default is not inspect.Parameter.empty}}
path = os.path.join(HERE, "reference_backend_signatures.json")
json.dump(out, open(path, "w"), indent=1, sort_keys=True)
print(f"wrote {path}: {len(out)} launchers")
