#!/bin/bash
# The host check (tests/hostcheck: the CUDA sources of everything but the fast kernel compiled for the host) under
# AddressSanitizer + UndefinedBehaviorSanitizer: engine threads / staging / error propagation, reference-order kernel, photon -> MCPE
# converter, step generator, reference-order table maker.  About five minutes on eight cores; not part of the default test run.
# usage: tools/hostcheck_asan.sh [log file]
set -e
cd "$(dirname "$0")/.."
LOG=${1:-profiles/hostcheck_asan_r02.txt}
LIB=$(PYTHONPATH=. python -c "from tests import hostcheck; print(hostcheck.build(sanitize=True))")
ASAN=$(gcc -print-file-name=libasan.so)
STDCPP=$(gcc -print-file-name=libstdc++.so)   # (preloaded as well: the interceptor of __cxa_throw needs it in a python process)
SELECT="(not two_converters_disjoint and not fast_kernel_converter_history and not device_rng_draw_assignment and not (steps_at_infinity and 0]) and not persistent_kernel and not bunches_generated_and_propagated and not converter_feeds_the_engine)"
{
  echo "# $(date -u +%Y-%m-%dT%H:%MZ)  g++ $(g++ -dumpversion), -fsanitize=address,undefined -O1 -g; library: ${LIB#$PWD/}"
  echo "# LD_PRELOAD=libasan.so libstdc++.so  ASAN_OPTIONS=detect_leaks=0:halt_on_error=1  UBSAN_OPTIONS=print_stacktrace=1"
  echo "# pytest tests/test_gpu_reference_kernel.py tests/test_gpu_engine.py tests/test_gpu_mcpe.py tests/test_gpu_stepgen.py tests/test_gpu_tabulator.py"
  echo "#        tests/test_zz_gpu_steps_at_infinity.py tests/hostcheck/check_bit_identity.py -m gpu -k \"$SELECT\""
} > "$LOG"
set +e
CLSIM_HOSTCHECK=1 CLSIMCU_LIB="$LIB" CLSIMCU_SAFEPRIMES_CACHE="$PWD/clsim_b200/data/safeprimes_base32.bin" LD_PRELOAD="$ASAN $STDCPP" ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=0 \
  python -m pytest tests/test_gpu_reference_kernel.py tests/test_gpu_engine.py tests/test_gpu_mcpe.py tests/test_gpu_stepgen.py tests/test_gpu_tabulator.py \
  tests/test_zz_gpu_steps_at_infinity.py tests/hostcheck/check_bit_identity.py -q -m gpu -p no:cacheprovider -k "$SELECT" > /tmp/hostcheck_asan.out 2>&1
RC=$?
{
  tail -3 /tmp/hostcheck_asan.out
  echo "pytest rc=$RC"
  echo "sanitizer reports (lines with 'runtime error' or 'AddressSanitizer'): $(grep -c 'runtime error\|AddressSanitizer' /tmp/hostcheck_asan.out)"
  grep 'runtime error\|ERROR: AddressSanitizer' /tmp/hostcheck_asan.out | sort | uniq -c | head -20
} >> "$LOG"
cat "$LOG"
exit $RC
