"""Extracts the public API of the reference's hot-path modules (names, argument lists, defaults) with `ast`
- nothing is imported or executed - and writes tests/golden/reference_api.json.  Run in the build container,
where /root/reference exists; the JSON is committed because the GPU box has no reference checkout."""
import ast
import json
import os

REF = "/root/reference"
FILES = {"modules/encoder.py": "encoder", "facenet_gpu.py": "facenet_gpu", "modules/hnsw_manager.py": "hnsw_manager",
         "processing/preprocess.py": "preprocess"}


def sig(fn: ast.FunctionDef):
    a = fn.args
    names = [x.arg for x in a.args]
    defaults = [ast.unparse(d) for d in a.defaults]
    return {"args": names, "defaults": defaults}


out = {}
for path, key in FILES.items():
    tree = ast.parse(open(os.path.join(REF, path)).read())
    mod = {"functions": {}, "classes": {}}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef):
            mod["functions"][node.name] = sig(node)
        elif isinstance(node, ast.ClassDef):
            mod["classes"][node.name] = {m.name: sig(m) for m in node.body if isinstance(m, ast.FunctionDef)}
    out[key] = mod
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_api.json"), "w") as f:
    json.dump(out, f, indent=1, sort_keys=True)
print({k: (list(v["functions"]), {c: len(m) for c, m in v["classes"].items()}) for k, v in out.items()})
