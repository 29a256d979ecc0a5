#!/bin/bash
# Runs tools/prof_prefilter.py with every library variant under build/variants (tuning sweeps).
cp interpolations.jl_b200/libb200itp.so /tmp/lib_default.so
for f in build/variants/*.so; do
  echo "== $f"
  cp "$f" interpolations.jl_b200/libb200itp.so
  timeout 120 python tools/prof_prefilter.py 2>&1 | tail -2
done
cp /tmp/lib_default.so interpolations.jl_b200/libb200itp.so
