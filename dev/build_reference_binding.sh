#!/bin/bash
# dev/build_reference_binding.sh — build the REFERENCE's own callers against libmoocore_b200.so, as INTEGRATION.md
# describes, so that the reference's own tests and doctests run against the swapped path.
#
#   1. copies /root/reference/{python,c} to a scratch directory (the reference tree is read-only),
#   2. applies the INTEGRATION.md change to python/src/moocore/_ffi_build.py:32-51,131-141 — "hvapprox.c" leaves the
#      `sources` list, the link line gains -lmoocore_b200 — and builds the CFFI module with the system gcc,
#   3. relinks the command line tool c/main-hvapprox.c (c/Makefile:195) without hvapprox.o / rng.o / mt19937.o,
#   4. stages the results under baseline/_ref/binding/ (git-ignored, travels to the GPU box):
#        moocore/                      the reference's Python package + _libmoocore.abi3.so bound to the library
#        moocore-*.dist-info/          so that importlib.metadata.version("moocore") works without an install
#        tests/, conftest.py           the reference's own python/tests and doctest conftest, unmodified
#        hvapprox_b200                 the relinked CLI
#        ref/                          the same two artefacts built from the UNMODIFIED reference (for A/B runs)
#
# Nothing here is product code and nothing is committed: it is the reference's code built two ways.
set -euo pipefail
REPO=$(cd "$(dirname "$0")/.." && pwd)
REF=${REF:-/root/reference}
OUT=$REPO/baseline/_ref/binding
LIBDIR=$REPO/moocore_b200
SCRATCH=$(mktemp -d)
trap 'rm -rf "$SCRATCH"' EXIT
export CC=/usr/bin/gcc CXX=/usr/bin/g++

[ -f "$LIBDIR/libmoocore_b200.so" ] || make -C "$REPO/moocore_b200/csrc"
VERSION=$(sed -n 's/^version *= *"\(.*\)"/\1/p' "$REF/python/pyproject.toml" | head -1)

build_module() {   # $1 = scratch subdir, $2 = "b200" | "ref"
    local W=$SCRATCH/$1
    mkdir -p "$W"
    cp -r "$REF/python" "$W/python"
    cp -r "$REF/c" "$W/c"
    chmod -R u+w "$W"
    rm -rf "$W/python/src/moocore/libmoocore"
    ln -s ../../../c "$W/python/src/moocore/libmoocore"
    if [ "$2" = b200 ]; then
        python3 - "$W/python/src/moocore/_ffi_build.py" "$LIBDIR" <<'EOF'
import sys
path, libdir = sys.argv[1], sys.argv[2]
s = open(path).read()
assert s.count('    "hvapprox.c",\n') == 1
s = s.replace('    "hvapprox.c",\n', '')                                              # _ffi_build.py:39
old = "    extra_link_args=ldflags,\n"
assert s.count(old) == 1
new = ('    extra_link_args=ldflags + ["-L%s", "-lmoocore_b200", "-Wl,-rpath,%s", '
       '"-Wl,-rpath,$ORIGIN/../../../../moocore_b200"],\n' % (libdir, libdir))           # _ffi_build.py:140
open(path, "w").write(s.replace(old, new))
EOF
    fi
    (cd "$W/python" && python3 src/moocore/_ffi_build.py > "$W/build.log" 2>&1) || { tail -30 "$W/build.log"; exit 1; }
}

stage_module() {   # $1 = scratch subdir, $2 = destination
    local W=$SCRATCH/$1 D=$2
    rm -rf "$D/moocore"
    mkdir -p "$D/moocore"
    cp "$W"/python/src/moocore/*.py "$W"/python/src/moocore/libmoocore.h "$D/moocore/"
    cp -r "$W/python/src/moocore/data" "$D/moocore/data"
    cp "$W"/python/moocore/_libmoocore*.so "$D/moocore/"
    mkdir -p "$D/moocore-$VERSION.dist-info"
    printf 'Metadata-Version: 2.1\nName: moocore\nVersion: %s\n' "$VERSION" > "$D/moocore-$VERSION.dist-info/METADATA"
}

build_cli() {      # $1 = output path, $2... = extra sources / link flags
    local out=$1; shift
    $CC -O2 -DNDEBUG -DDEBUG=0 -D_GNU_SOURCE '-DVERSION="0.20"' '-DMARCH="x86-64-v2"' -march=x86-64-v2 -w -I"$REF/c" -o "$out" \
        "$REF/c/main-hvapprox.c" "$REF/c/timer.c" "$REF/c/cmdline.c" "$REF/c/io.c" "$@" -lm
}

build_module b200 b200
build_module ref ref
rm -rf "$OUT"
mkdir -p "$OUT/ref"
stage_module b200 "$OUT"
stage_module ref "$OUT/ref"
cp -r "$REF/python/tests" "$OUT/tests"
cp "$REF/python/conftest.py" "$OUT/conftest.py"
# pytest options of python/pyproject.toml:208-211 (doctest NUMBER flag, importlib import mode)
printf '[pytest]\ndoctest_optionflags = NUMBER\naddopts = --import-mode=importlib -p no:cacheprovider\n' > "$OUT/pytest.ini"
cp "$OUT/pytest.ini" "$OUT/ref/pytest.ini"
cp -r "$REF/python/tests" "$OUT/ref/tests"
cp "$REF/python/conftest.py" "$OUT/ref/conftest.py"

# c/Makefile:195 — hvapprox: main-hvapprox.o timer.o [hvapprox.o rng.o mt19937/mt19937.o] + cmdline.o io.o
build_cli "$OUT/hvapprox_b200" -L"$LIBDIR" -lmoocore_b200 -Wl,-rpath,"$LIBDIR" '-Wl,-rpath,$ORIGIN/../../../moocore_b200'
build_cli "$OUT/ref/hvapprox_ref" "$REF/c/hvapprox.c" "$REF/c/rng.c" "$REF/c/mt19937/mt19937.c"
printf '5 5\n4 6\n2 7\n7 4\n' > "$OUT/kat2d.dat"

# evidence that the swapped module really binds the library
echo "--- undefined hv_approx symbols of the swapped CFFI module (resolved by libmoocore_b200.so):"
nm -D --undefined-only "$OUT"/moocore/_libmoocore*.so | grep hv_approx || { echo "hv_approx_* not undefined: swap failed"; exit 1; }
echo "--- the unmodified module defines them itself:"
nm -D "$OUT"/ref/moocore/_libmoocore*.so | grep -c ' T hv_approx' || true
ldd "$OUT"/moocore/_libmoocore*.so | grep moocore_b200
ldd "$OUT/hvapprox_b200" | grep moocore_b200
echo "staged in $OUT"
