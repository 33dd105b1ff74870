#!/usr/bin/env bash
# oracle/build_ref.sh -- TEST INFRASTRUCTURE.
#
# Compiles the UNMODIFIED reference (yandex/pire) from the sources where they
# lie under $PIRE_REF (default /root/reference) into oracle/_ref/ :
#     oracle/_ref/libpire_ref.so   reference library + oracle/ref_capi.cpp (C face)
#     oracle/_ref/pire_ut          the reference's own tests/pire_ut.cpp + easy_ut.cpp
#     oracle/_ref/pire_bench       the reference's own tools/bench/bench.cpp
# No reference source is copied into the repository; oracle/_ref/ is git-ignored
# (but travels to the GPU box with gpurun).  The reference's build system
# (autotools + bison + flex) is not run: bison is absent from this image, so the
# generated parser is replaced by oracle/yre_parse_rd.cpp (our own recursive-
# descent driver) plus the grammar file's two helper functions, which are
# extracted from re_parser.y at build time into the git-ignored gen/ directory.
set -euo pipefail

HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${PIRE_REF:-/root/reference}"
OUT="$HERE/_ref"
GEN="$OUT/gen"
OBJ="$OUT/obj"
CXX="${CXX:-g++}"
# -include limits: pire/extra/count.cpp uses std::numeric_limits without including <limits> (fine with the
# compilers of its day, an error with gcc 13); the reference source is left untouched.
CXXFLAGS="-std=c++11 -O2 -w -fPIC -DPIRE_NO_CONFIG -include limits"

if [ ! -d "$REF/pire" ]; then
    echo "build_ref.sh: reference tree $REF not present; keeping prebuilt oracle/_ref" >&2
    [ -f "$OUT/libpire_ref.so" ] && exit 0
    exit 3
fi

mkdir -p "$GEN" "$OBJ"

# --- generated pieces (git-ignored) -----------------------------------------
# Token ids: only re_lexer.cpp:163-178 and the parser see them; any distinct
# values above the byte range work.
cat > "$GEN/re_parser.h" <<'EOF'
#ifndef PIRE_ORACLE_RE_PARSER_H
#define PIRE_ORACLE_RE_PARSER_H
enum { YRE_LETTERS = 258, YRE_COUNT = 259, YRE_DOT = 260, YRE_AND = 261, YRE_NOT = 262 };
#endif
EOF
# AppendRange + ConvertToFSM, verbatim from the grammar file's epilogue.
s=$(grep -n '^void AppendRange(const Encoding& encoding' "$REF/pire/re_parser.y" | tail -1 | cut -d: -f1)
e=$(grep -n '^} // namespace' "$REF/pire/re_parser.y" | head -1 | cut -d: -f1)
sed -n "${s},$((e-1))p" "$REF/pire/re_parser.y" > "$GEN/re_parser_helpers.inc"

INC="-I$REF -I$REF/pire -I$GEN -I$HERE"

LIBSRC="approx_matching classes easy encoding fsm half_final_fsm re_lexer read_unicode scanner_io scanners/null stub/utf8 extra/count extra/capture extra/glyphs"
OBJS=""
for f in $LIBSRC; do
    o="$OBJ/$(echo "$f" | tr / _).o"
    if [ ! -f "$o" ] || [ "$REF/pire/$f.cpp" -nt "$o" ]; then
        $CXX $CXXFLAGS $INC -c "$REF/pire/$f.cpp" -o "$o" &
    fi
    OBJS="$OBJS $o"
done
$CXX $CXXFLAGS $INC -c "$HERE/yre_parse_rd.cpp" -o "$OBJ/re_parser.o" &
$CXX $CXXFLAGS $INC -c "$HERE/ref_capi.cpp" -o "$OBJ/ref_capi.o" &
wait
OBJS="$OBJS $OBJ/re_parser.o"

$CXX -shared -o "$OUT/libpire_ref.so" $OBJS "$OBJ/ref_capi.o" -lpthread

# --- the reference's own unit tests as the gate ------------------------------
$CXX $CXXFLAGS $INC -I"$REF/tests" \
    "$REF/tests/stub/cppunit.cpp" "$REF/tests/pire_ut.cpp" "$REF/tests/easy_ut.cpp" "$REF/tests/count_ut.cpp" \
    $OBJS -o "$OUT/pire_ut" &
# --- the reference's own benchmark driver -------------------------------------
$CXX $CXXFLAGS $INC -I"$REF/tools" -I"$REF/tools/common" \
    "$REF/tools/bench/bench.cpp" $OBJS -o "$OUT/pire_bench" &
wait

# the reference's own benchmark text (tools/bench/test_file, 20 KB of prose) beside the binary, for
# continuity runs of tools/bench on the GPU box's host (run-bench doubles it to >= 300 MB)
cp "$REF/tools/bench/test_file" "$OUT/test_file"

if [ "${1:-}" != "--no-check" ]; then
    "$OUT/pire_ut" > "$OUT/pire_ut.log" 2>&1 || { tail -20 "$OUT/pire_ut.log"; echo "reference unit tests FAILED" >&2; exit 1; }
    tail -1 "$OUT/pire_ut.log"
fi
echo "oracle/_ref built from $REF"
