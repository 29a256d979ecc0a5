#!/bin/bash
# Runs tools/perf_probe.py (prefilter lines of C2 / C4) with every library variant under build/variants.
cp interpolations.jl_b200/libb200itp.so /tmp/lib_default.so
for f in build/variants/*.so; do
  echo "== $f"
  cp "$f" interpolations.jl_b200/libb200itp.so
  timeout 300 python tools/perf_probe.py 2>&1 | grep -E "^C2 4096|^C4 8192|512\^3 f64" | sed 's/"axis0.*"all"/"all"/'
done
cp /tmp/lib_default.so interpolations.jl_b200/libb200itp.so
