"""Extracts the Python-visible surface of the reference's pybind11 module (src/psdr.cpp:41-295: classes, bases, methods, properties)
into tests/golden/ref_python_surface.json, which tests/test_host_module.py checks `import psdr_cuda` (this repo's host module) against.

    python tests/golden/make_ref_surface.py      (needs /root/reference)"""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.environ.get("PSDR_REFERENCE", "/root/reference")


def extract(src):
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)                                   # the commented-out Interaction / Sampler bindings
    src = re.sub(r"#ifdef PSDR_MESH_ENABLE_1D_VERTEX_OFFSET.*?#endif", "", src, flags=re.S)   # not defined in macros.h
    src = "\n".join(l for l in src.splitlines() if not l.strip().startswith("//"))
    out = {}
    for m in re.finditer(r"py::class_<([^>]*)>\(m, \"(\w+)\"\)(.*?);\n", src, flags=re.S):
        cpp, name, body = m.group(1), m.group(2), m.group(3)
        bases = [b.strip() for b in cpp.split(",")[1:]]
        out[name] = {
            "cpp": cpp.split(",")[0].strip(), "cpp_bases": bases,
            "methods": sorted(set(re.findall(r"\.def\(\"(\w+)\"", body))),
            "properties": sorted(set(re.findall(r"\.def_read(?:write|only)\(\"(\w+)\"", body))),
            "constructors": len(re.findall(r"\.def\(py::init<", body)),
        }
    return out


def main():
    with open(os.path.join(REF, "src", "psdr.cpp")) as fh:
        surface = extract(fh.read())
    path = os.path.join(ROOT, "tests", "golden", "ref_python_surface.json")
    with open(path, "w") as fh:
        json.dump(surface, fh, indent=1, sort_keys=True)
    print("wrote", path, len(surface), "classes")


if __name__ == "__main__":
    main()
