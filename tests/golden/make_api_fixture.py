#!/usr/bin/env python
"""Generate tests/golden/sam_api_used.json: every `sam.<name>` the reference's training scripts under
example/samgraph use, and whether the reference's own package (samgraph/common/__init__.py +
samgraph/torch/adapter.py) defines that name.  Run in the build container (needs /root/reference).

  python tests/golden/make_api_fixture.py
"""
import ast
import glob
import json
import os
import re

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def defined_names(path):
    tree = ast.parse(open(path).read())
    names = set()
    for node in ast.walk(tree):
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)):
            names.add(node.name)
        elif isinstance(node, ast.Assign):
            for t in node.targets:
                for n in ast.walk(t):
                    if isinstance(n, ast.Name):
                        names.add(n.id)
                    elif isinstance(n, ast.Attribute):
                        names.add(n.attr)
    return names


def main():
    used = {}
    for f in sorted(glob.glob(os.path.join(REF, "example/samgraph/**/*.py"), recursive=True)):
        rel = os.path.relpath(f, REF)
        for m in re.finditer(r"\bsam\.([A-Za-z_][A-Za-z0-9_]*)", open(f).read()):
            used.setdefault(m.group(1), set()).add(rel)
    defined = defined_names(os.path.join(REF, "samgraph/common/__init__.py")) | \
        defined_names(os.path.join(REF, "samgraph/torch/adapter.py"))
    out = {name: {"defined_in_reference": name in defined, "scripts": sorted(files)} for name, files in sorted(used.items())}
    with open(os.path.join(HERE, "sam_api_used.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("%d names, %d not defined by the reference itself: %s"
          % (len(out), sum(not v["defined_in_reference"] for v in out.values()),
             [k for k, v in out.items() if not v["defined_in_reference"]]))


if __name__ == "__main__":
    main()
