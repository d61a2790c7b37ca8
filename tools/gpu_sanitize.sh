#!/bin/bash
# compute-sanitizer memcheck on a small residual + SAO run (one tool per gpurun call)
OUT=gpurun_out; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/san_plain.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/san_plain.log; exit 1; }
compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/sanitizer_memcheck.log 2>&1
echo "memcheck exit $?"; grep -E "ERROR SUMMARY|smoke ok|Invalid|out of bounds" $OUT/sanitizer_memcheck.log | head
