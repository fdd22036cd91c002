#!/usr/bin/env bash
# Run the reference's OWN device-plugin conformance binary (docs/template_plugin/tests/functional -> templateFuncTests, built
# UNMODIFIED by oracle/build_ref.sh: behaviour suites, hetero / single-layer instances) against libb200Plugin.so.
#
# How: the binary asks the Core for the library "templatePlugin" under the device name "TEMPLATE"
# (inference-engine/tests/functional/plugin/shared/include/behavior/core_integration.hpp:65-71,
#  inference-engine/tests/ie_test_utils/functional_test_utils/src/plugin_cache.cpp:52-57).  A scratch library directory holds the
# reference's libraries plus a copy of libb200Plugin.so under that file name, so every test of the binary lands in the B200 plugin.
# Nothing of the reference is modified and no test source lives in this repository.
#
#   tools/run_conformance.sh [gtest filter] [output prefix]
# Writes <prefix>.xml (gtest report), <prefix>.log and <prefix>_summary.txt (per-suite pass / fail counts + the failing tests).
set -uo pipefail
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
REF="$ROOT/oracle/_ref"
FILTER="${1:-*}"
PREFIX="${2:-$ROOT/gpurun_out/conformance}"
case "$PREFIX" in /*) ;; *) PREFIX="$PWD/$PREFIX" ;; esac
BIN="$REF/bin/templateFuncTests"
[ -x "$BIN" ] || { echo "run_conformance: $BIN missing (oracle/build_ref.sh builds it)" >&2; exit 3; }
[ -f "$ROOT/plugin/lib/libb200Plugin.so" ] || { echo "run_conformance: plugin not built" >&2; exit 3; }

ALIAS="$(mktemp -d /tmp/b200_conformance.XXXXXX)"
trap 'rm -rf "$ALIAS"' EXIT
cp "$REF"/lib/*.so "$ALIAS/"
cp -f "$ROOT/plugin/lib/libb200Plugin.so" "$ALIAS/libtemplatePlugin.so"
# the registry the reference ships has no TEMPLATE entry either: the tests register it themselves
cat > "$ALIAS/plugins.xml" <<XML
<ie>
    <plugins>
        <plugin name="CPU" location="libMKLDNNPlugin.so"></plugin>
        <plugin name="MULTI" location="libMultiDevicePlugin.so"></plugin>
        <plugin name="HETERO" location="libHeteroPlugin.so"></plugin>
    </plugins>
</ie>
XML
mkdir -p "$(dirname "$PREFIX")"
# fp32 execution of fp32 networks: the behaviour suites compare values for equality (the reference's CPU instances switch its implicit
# bf16 mode off for the same reason: inference-engine/tests/functional/plugin/cpu/shared_tests_instances/core_config.cpp)
export B200_INFERENCE_PRECISION="${B200_INFERENCE_PRECISION:-FP32}"
cd "$ALIAS"   # the caching tests write their cache directories and IR files into the working directory
LD_LIBRARY_PATH="$ALIAS:$ROOT/dldt_b200/lib:${LD_LIBRARY_PATH:-}" timeout "${CONFORMANCE_TIMEOUT:-900}" \
  "$BIN" --gtest_filter="$FILTER" --gtest_output="xml:$PREFIX.xml" ${CONFORMANCE_ARGS:-} > "$PREFIX.log" 2>&1
RC=$?
cd "$ROOT"
python "$ROOT/tools/conformance_summary.py" "$PREFIX.xml" "$PREFIX.log" > "${PREFIX}_summary.txt" 2>&1 || true
tail -n 40 "${PREFIX}_summary.txt"
echo "templateFuncTests exit code $RC"
exit 0
