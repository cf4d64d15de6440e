#!/bin/bash
# Parity suite against a build with every MAS_CHECK index / protocol assertion compiled in
# (-DMAS_DEBUG_CHECKS): a violation prints the expression and traps, which fails the test.
# compute-sanitizer is closed on this pool (profiles/r2_sanitizer_closed.txt).
set -e
python profiles/build_debug.py
MAS_LIB_PATH=art_tts_b200/lib/libmas_dbg.so python -m pytest tests -m gpu -x -q
