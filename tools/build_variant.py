#!/usr/bin/env python
"""Build a tuning variant of libtpb200.so into build_variants/lib_<name>.so:
    tools/build_variant.py <name> [-DFLAG=... ...]     (flags go to the 3-D Float32 unit `3ff`)
Only the unit whose flags changed is recompiled (object cache in build/obj)."""
import importlib.util, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("_b", os.path.join(ROOT, "trixiparticles.jl_b200", "build.py"))
b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
name, flags = sys.argv[1], sys.argv[2:]
units = os.environ.get("TPB_VARIANT_UNITS", "3ff").split(",")
os.makedirs(os.path.join(ROOT, "build_variants"), exist_ok=True)
out = os.path.join(ROOT, "build_variants", f"lib_{name}.so")
print(b.build(out=out, unit_flags={u: flags for u in units} if flags else None))
