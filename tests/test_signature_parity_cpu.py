"""Drop-in check of the boundary: every launcher in curobo_b200/backends/* takes the reference launcher's parameters, by the
same names and in the same order (so the reference's call sites -- positional or keyword -- bind unchanged); extra parameters
are allowed only after them and only with defaults.  The reference side comes from tests/golden/reference_backend_signatures.json
(recorded from curobo/_src/curobolib/backends/cuda_core_backend/*.py by tests/golden/make_signature_golden.py)."""
import importlib
import inspect
import json
import os

import pytest

REF = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_backend_signatures.json")))


@pytest.mark.parametrize("key", sorted(REF))
def test_launcher_signature_is_a_drop_in(key):
    mod, name = key.split(".")
    ours = getattr(importlib.import_module(f"curobo_b200.backends.{mod}"), name)
    got = list(inspect.signature(ours).parameters.values())
    want = REF[key]["params"]
    names = [p.name for p in got]
    assert names[:len(want)] == want, f"{key} ({REF[key]['file']}:{REF[key]['line']}): {names} vs {want}"
    for p in got[len(want):]:
        assert p.default is not inspect.Parameter.empty, f"{key}: extension parameter {p.name} must be optional"
    for pname in REF[key]["defaults"]:            # what the reference lets callers omit, we let them omit too
        p = inspect.signature(ours).parameters[pname]
        assert p.default is not inspect.Parameter.empty, f"{key}: {pname} is optional in the reference"
