#!/bin/sh
# Compiles gen_glm_golden.cpp against the reference's vendored glm IN PLACE and rewrites
# tests/golden/glm_golden.json. Outputs (the binary) go to oracle/_ref/ only.
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
REF=${REFERENCE_ROOT:-/root/reference}
mkdir -p "$HERE/../_ref"
/usr/bin/g++ -std=c++17 -O0 -ffp-contract=off -I"$REF/Dependencies" -o "$HERE/../_ref/gen_glm_golden" "$HERE/gen_glm_golden.cpp"
"$HERE/../_ref/gen_glm_golden" "$HERE/../../tests/golden/glm_golden.json"
echo "wrote tests/golden/glm_golden.json"
