"""The drop-in classes against the public surface of the reference classes they stand in for
(tests/golden/reference_api.json: method and attribute NAMES taken from the reference's sources by
tests/golden/make_reference_api.py).  Every method a script can call from Python scope -- plain methods and
@ti.kernel methods -- must exist here; @ti.func methods are device functions (Taichi refuses to call them outside
a kernel) and are not part of the host surface.  Every attribute the reference's methods assign on `self` must be
there too, except the scratch listed below with the reason."""
import ast
import importlib
import inspect
import json
import os

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
API = json.load(open(os.path.join(HERE, "golden", "reference_api.json")))

# reference attributes the drop-ins do not carry, and why
NOT_CARRIED = {
    "ParticleSystemV4": {
        # the second copy of every array that resort() scatters into and copies back from (partice_systemv4.py:226-249);
        # nothing reads them outside resort().  The library ping-pongs two record arrays instead.
        "x_buffer", "v_buffer", "density_buffer", "pressure_buffer", "material_buffer", "color_buffer", "mass_buffer",
        "volume_buffer", "m_buffer", "grid_ids_buffer"},
    "ParticleSystem": {
        # Taichi SNode handles of the field layout (partice_system.py:41-57)
        "particle_node", "particles_node",
        # dense cell lists [cell, 100], read only by search_neighbors (:111-113); the library keeps a cell-sorted
        # index array instead and exposes what they are for: particle_neighbors / particle_neighbors_num, bit-exact
        "grid_particles"},
}
NOT_CARRIED["ParticleSystemV2"] = NOT_CARRIED["ParticleSystem"]


def _ours(key):
    rel, cls = key.split(":")
    return getattr(importlib.import_module(rel[:-3].replace("/", ".")), cls)


def _names(klass):
    """names the class (and its bases) defines: class attributes / properties / methods, attributes assigned on
    self, and the field names handed to setattr loops (string constants)"""
    names = set()
    for k in klass.__mro__:
        if k is object:
            continue
        names |= set(vars(k))
        for n in ast.walk(ast.parse(inspect.getsource(k))):
            if isinstance(n, ast.Attribute) and isinstance(n.value, ast.Name) and n.value.id == "self" \
                    and isinstance(n.ctx, ast.Store):
                names.add(n.attr)
            if isinstance(n, ast.Constant) and isinstance(n.value, str) and n.value.isidentifier():
                names.add(n.value)
    return names


@pytest.mark.parametrize("key", sorted(API))
def test_every_host_callable_method_of_the_reference_class_exists(key):
    ref, ours = API[key], _ours(key)
    host = [m for m in ref["methods"] if ref["kinds"][m] != "func"]
    assert len(host) >= 4
    missing = [m for m in host if not callable(getattr(ours, m, None))]
    assert not missing, missing
    # same positional parameter names for the methods scripts call with arguments
    for m in ("add_cube", "add_particles", "compute_cube_particles_num", "copy_to_numpy", "load_rigid_body"):
        if m in host:
            params = list(inspect.signature(getattr(ours, m)).parameters)
            assert params[0] == "self" and len(params) >= 2, (m, params)


@pytest.mark.parametrize("key", sorted(API))
def test_every_attribute_of_the_reference_class_exists_or_is_accounted_for(key):
    ref, ours = API[key], _ours(key)
    have = _names(ours)
    missing = set(ref["attributes"]) - have
    assert missing == NOT_CARRIED.get(ours.__name__, set()) & set(ref["attributes"]), sorted(missing)


def test_the_one_dead_method_of_the_reference_raises_like_the_reference():
    """ParticleSystemV4.search_neighbors reads four names the class never defines (found mechanically by
    make_reference_api.py): the reference raises when it is called; so does the drop-in, without touching a GPU."""
    dead = {k.split(":")[1]: v["undefined_reads"] for k, v in API.items() if v["undefined_reads"]}
    assert dead == {"ParticleSystemV4": {"search_neighbors": ["grid_particles", "particle_neighbors",
                                                              "particle_neighbors_num", "support_radius"]}}
    ours = _ours("core/partice_system/partice_systemv4.py:ParticleSystemV4")
    with pytest.raises(AttributeError):
        ours.search_neighbors(object.__new__(ours))


def test_solver_classes_keep_the_reference_inheritance():
    for key, ref in API.items():
        for base in ref["bases"]:
            assert base in [k.__name__ for k in _ours(key).__mro__[1:]], (key, base)
