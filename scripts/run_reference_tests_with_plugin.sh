#!/bin/bash
# Runs the reference's OWN test suites with the B200 plug-in installed in interpolator_map (build container only:
# needs /root/reference).  Every test that does not ask for solver='b200' must behave as with the stock classes.
#   scripts/run_reference_tests_with_plugin.sh [extra pytest args]
ROOT=$(cd "$(dirname "$0")/.." && pwd)
export PYTHONPATH=$ROOT/tests/golden/_stubs:/root/reference/packages/loop_common/src:/root/reference/packages/loop_interpolation/src:/root/reference:$ROOT:$ROOT/scripts
export PYTHONDONTWRITEBYTECODE=1
cd /tmp
echo "== loop_interpolation tests, plug-in installed"
python -m pytest /root/reference/packages/loop_interpolation/tests -p no:cacheprovider -p b200_pytest_plugin -q --rootdir=/tmp -c /dev/null "$@" 2>&1 | tail -2
echo "== loop_interpolation tests, stock"
python -m pytest /root/reference/packages/loop_interpolation/tests -p no:cacheprovider -q --rootdir=/tmp -c /dev/null "$@" 2>&1 | tail -2
echo "== LoopStructural tests/unit/interpolator + tests/integration (GeologicalModel -> InterpolatorFactory -> interpolator_map), plug-in installed"
python -m pytest /root/reference/tests/unit/interpolator /root/reference/tests/integration -p no:cacheprovider -p b200_pytest_plugin -q --rootdir=/root/reference/tests --confcutdir=/root/reference -c /dev/null "$@" 2>&1 | tail -2
echo "== the same, stock"
python -m pytest /root/reference/tests/unit/interpolator /root/reference/tests/integration -p no:cacheprovider -q --rootdir=/root/reference/tests --confcutdir=/root/reference -c /dev/null "$@" 2>&1 | tail -2
