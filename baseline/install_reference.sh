#!/usr/bin/env bash
# Installs the UNMODIFIED reference (read-only at /root/reference) into baseline/_ref for `bench.py --impl reference`.
# The reference ships no setup.py/pyproject, so `pip install /root/reference` fails ("not installable"); we install
# from a /tmp copy to which ONLY a packaging stub (setup.py) is added.  No reference source file is changed.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
SRC=/tmp/ref_src_$$
rm -rf "$SRC" "$HERE/_ref"
cp -r /root/reference "$SRC"
chmod -R u+w "$SRC"
cat > "$SRC/setup.py" <<'PY'
import os
from setuptools import setup
pkgs, data = [], {}
for root, dirs, files in os.walk("."):
    dirs[:] = [d for d in dirs if d not in (".git", ".github", "doc", "build", "dist") and not d.endswith(".egg-info")]
    rel = os.path.relpath(root, ".")
    if rel == "." or "-" in rel:
        continue
    if any(f.endswith(".py") for f in files) or any(f.endswith((".yaml", ".md")) for f in files):
        pkg = rel.replace(os.sep, ".")
        pkgs.append(pkg)
        data[pkg] = ["*.yaml", "*.yml", "*.md", "*.txt", "*.json"]
setup(name="msrflute-reference", version="1.0.0", packages=pkgs, package_data=data, py_modules=["e2e_trainer"])
PY
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target "$HERE/_ref" "$SRC" 2>&1 | tail -3
rm -rf "$SRC"
ls "$HERE/_ref" | head -20
