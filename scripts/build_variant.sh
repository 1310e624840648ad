#!/bin/bash
# Build an A/B variant of the library next to the product one (no GPU needed; nvcc cross-compiles):
#   scripts/build_variant.sh <name> "<extra nvcc flags, e.g. -DSBX_ARRIVE_SINK=0>"   ->  csrc/ab/libsbx_<name>.so
# scripts/ab_variants.sh then times base and variants on the same box.
set -e
NAME=$1; FLAGS=$2
D=$(cd "$(dirname "$0")/../soap_bubble_simulation_b200/csrc" && pwd)
T=$(mktemp -d)
cp "$D"/*.cu "$D"/*.cuh "$D"/*.h "$D"/Makefile "$T"/
mkdir -p "$T/../../include" 2>/dev/null || true
sed -i "s|\.\./\.\./include/sbx.h|$D/../../include/sbx.h|" "$T"/Makefile "$T"/sbx_internal.cuh
make -C "$T" -j8 -s EXTRA="$FLAGS"
mkdir -p "$D/ab"
cp "$T/libsbx.so" "$D/ab/libsbx_$NAME.so"
rm -rf "$T"
echo "built $D/ab/libsbx_$NAME.so"
