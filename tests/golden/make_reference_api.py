#!/usr/bin/env python
"""Lists the public surface of the reference classes this repository stands in for: per class, the names of its
methods and of the attributes its methods assign on `self` (taken from the reference's sources with `ast`; no
code is copied, only names).  TEST INFRASTRUCTURE: runs in the build container, where /root/reference exists,
and writes tests/golden/reference_api.json for tests/test_cpu_api_surface.py.

    python tests/golden/make_reference_api.py [--ref /root/reference]
"""
import argparse
import ast
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# reference file -> classes on the hot path; partice_systemv3 is unrunnable, boundary_akinci empty
FILES = {
    "core/partice_system/partice_systemv4.py": ["ParticleSystemV4"],
    "core/partice_system/partice_system.py": ["ParticleSystem"],
    "core/partice_system/partice_systemv2.py": ["ParticleSystemV2"],
    "core/sph/sph_basev2.py": ["SPHBaseV2"],
    "core/sph/wcsphv2.py": ["WCSPHV2"],
    "core/sph/sph_base.py": ["SPHBase"],
    "core/sph/wcsph.py": ["WCSPH"],
}


def surface(cls):
    methods, attrs, kinds, reads = [], set(), {}, {}
    for node in cls.body:
        if isinstance(node, ast.FunctionDef):
            methods.append(node.name)
            decos = [ast.unparse(d) for d in node.decorator_list]
            # ti.kernel and plain methods are called from Python scope; a ti.func only from inside a kernel
            kinds[node.name] = "func" if "ti.func" in decos else "kernel" if "ti.kernel" in decos else "python"
            reads[node.name] = sorted({leaf.attr for leaf in ast.walk(node)
                                       if isinstance(leaf, ast.Attribute) and isinstance(leaf.value, ast.Name)
                                       and leaf.value.id == "self" and isinstance(leaf.ctx, ast.Load)})
            for sub in ast.walk(node):
                targets = []
                if isinstance(sub, ast.Assign):
                    targets = sub.targets
                elif isinstance(sub, (ast.AugAssign, ast.AnnAssign)):
                    targets = [sub.target]
                for t in targets:
                    for leaf in ast.walk(t):
                        if isinstance(leaf, ast.Attribute) and isinstance(leaf.value, ast.Name) and leaf.value.id == "self" \
                                and isinstance(leaf.ctx, ast.Store):
                            attrs.add(leaf.attr)
    bases = [b.id if isinstance(b, ast.Name) else ast.unparse(b) for b in cls.bases]
    return {"bases": bases, "methods": methods, "kinds": kinds, "attributes": sorted(attrs), "reads": reads}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    args = ap.parse_args()
    out = {}
    for rel, classes in FILES.items():
        tree = ast.parse(open(os.path.join(args.ref, rel), encoding="utf-8").read())
        for node in tree.body:
            if isinstance(node, ast.ClassDef) and node.name in classes:
                out[f"{rel}:{node.name}"] = surface(node)
    # a method that reads a `self.` name which neither the class nor its bases ever define cannot run in the reference
    by_name = {k.split(":")[1]: v for k, v in out.items()}
    for v in out.values():
        known, todo = set(), [v]
        while todo:
            c = todo.pop()
            known |= set(c["attributes"]) | set(c["methods"])
            todo += [by_name[b] for b in c["bases"] if b in by_name]
        v["undefined_reads"] = {m: sorted(set(r) - known) for m, r in v["reads"].items() if set(r) - known}
        del v["reads"]
    path = os.path.join(HERE, "reference_api.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(path, {k: (len(v["methods"]), len(v["attributes"])) for k, v in out.items()})


if __name__ == "__main__":
    main()
