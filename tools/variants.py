"""Tuning sweep helper: build the native library under several -D settings (launch-bound /
occupancy targets) into tce-rvos_b200/_native_variants/<name>/ and print the kbench command lines.
Developer tool only.  Usage: python tools/variants.py name:DEF1=V,DEF2=V ..."""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
build = importlib.import_module("tce-rvos_b200.build")

for spec in sys.argv[1:]:
    name, _, defs = spec.partition(":")
    d = build.build_variant(name, [x for x in defs.split(",") if x])
    print(f"MSDA_NATIVE_DIR={d} python tools/kbench.py")
