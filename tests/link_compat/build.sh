#!/bin/sh
# Link-compatibility check ("existing callers link unchanged"): compile the reference's OWN tutorial
# tutorials/minimal/minimal.cpp, untouched and against the reference's OWN headers, and link it against
# libembree4_b200.so instead of libembree4.so.  Only the resulting binary is kept (tests/link_compat/_bin, git-ignored,
# travels to the GPU box); no reference source enters the repo.  Runs only where /root/reference exists.
set -e
cd "$(dirname "$0")"
REF=${EMBREE_REFERENCE:-/root/reference}
GEN=../../oracle/_ref/gen_rel/include/embree4
[ -f "$REF/tutorials/minimal/minimal.cpp" ] && [ -d "$GEN" ] || { echo "reference not present: keeping prebuilt binary"; exit 0; }
mkdir -p _bin
ln -sf ../../../embree_b200/csrc/libembree4_b200.so _bin/libembree4.so
g++ -O1 -std=c++11 -I"$REF/include" -I"$GEN" -o _bin/embree_minimal "$REF/tutorials/minimal/minimal.cpp" \
    -L_bin -lembree4 -Wl,-rpath,'$ORIGIN/../../../embree_b200/csrc'
echo built tests/link_compat/_bin/embree_minimal
