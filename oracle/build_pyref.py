"""TEST INFRASTRUCTURE -- recipe for oracle/_ref/pyref: the reference's own Python call sites of the hot path, compiled to
byte code where they lie under /root/reference.

The reference's autograd Functions (curobo/_src/curobolib/cuda_ops/{kinematics,geometry,trajectory,optimization,dynamics}.py)
are the CALLERS of the kernel-backend modules this repository replaces.  tests/test_gpu_reference_callsites.py runs THEIR
forward / backward on top of curobo_b200.backends (INTEGRATION.md section 1 overlay) on the GPU box -- where /root/reference
does not exist.  Like oracle/_ref/libcurobo_ref.so (the reference's CUDA kernels compiled from their sources), this recipe
compiles the needed reference modules from the sources where they lie into a build product under oracle/_ref/ (git-ignored,
travels to the GPU box): sourceless .pyc files, one per module of the import closure of the call sites.  No reference source
is copied into the repository.

    python oracle/build_pyref.py            # needs /root/reference; writes oracle/_ref/pyref/ + MANIFEST.json
"""
import importlib
import json
import os
import py_compile
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = "/root/reference"
OUT = os.path.join(ROOT, "oracle", "_ref", "pyref")
CALL_SITES = ["curobo._src.curobolib.cuda_ops.kinematics", "curobo._src.curobolib.cuda_ops.geometry",
              "curobo._src.curobolib.cuda_ops.trajectory", "curobo._src.curobolib.cuda_ops.optimization",
              "curobo._src.curobolib.cuda_ops.dynamics", "curobo._src.curobolib.backends",
              # the optimizer that drives the Rollout protocol (its step direction and line search call the backends too)
              "curobo._src.optim.gradient.lbfgs"]


def build(verbose: bool = False) -> str:
    if not os.path.isdir(REFERENCE):
        raise RuntimeError("/root/reference is not present")
    sys.path[:0] = [os.path.join(ROOT, "tests", "golden"), ROOT]
    import _reference_under_shim as shim
    shim.prepare()
    for m in CALL_SITES:
        importlib.import_module(m)
    mods = {k: v.__file__ for k, v in sys.modules.items()
            if (k == "curobo" or k.startswith("curobo.")) and getattr(v, "__file__", None)
            and os.path.abspath(v.__file__).startswith(REFERENCE + os.sep)}
    if os.path.isdir(OUT):
        shutil.rmtree(OUT)
    manifest = {}
    for name, src in sorted(mods.items()):
        rel = os.path.relpath(src, REFERENCE)
        dst = os.path.join(OUT, rel + "c")                     # x.py -> x.pyc next to where the source would be
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        py_compile.compile(src, cfile=dst, dfile=os.path.join("<reference>", rel), doraise=True)
        manifest[name] = rel
        if verbose:
            print("compiled", rel)
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
        json.dump({"python": sys.version.split()[0], "modules": manifest}, f, indent=1)
    return OUT


if __name__ == "__main__":
    print(build(verbose=True))
