#!/usr/bin/env bash
# Build the UNMODIFIED reference (OpenVINO 2021.4-era Inference Engine + MKLDNNPlugin CPU oracle,
# MULTI plugin, IR reader, benchmark_app) from the sources where they lie under /root/reference,
# and install ONLY build outputs (shared libraries + benchmark_app) into oracle/_ref/.
#
# This is test infrastructure: oracle/_ref is (1) the parity oracle ("CPU" device) for the B200
# plugin, (2) the CPU baseline timed by `bench.py --impl reference`, (3) the libinference_engine.so
# the B200 plugin is a drop-in for.  No reference SOURCE is copied into the repository: the
# writable copy the reference's build insists on (it writes inference-engine/temp and bin/ into
# its own source dir, cmake/dependencies.cmake:8) lives under /tmp and is deleted afterwards
# unless KEEP_REF_BUILD=1.
#
# Recipe = SURVEY.md §8c / BASELINE.md §4 (three workarounds: force-include <cstdint> for gcc 13,
# /usr/bin/g++ instead of the static-libstdc++ /opt/gcc wrapper, drop the OpenCV "dependency" of
# benchmark_app which it only uses for -load_config/-dump_config).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${REFERENCE_DIR:-/root/reference}"
WORK="${REF_BUILD_DIR:-/tmp/b200_refbuild}"
OUT="$HERE/_ref"
JOBS="${JOBS:-$(nproc)}"

if [ ! -d "$REF/inference-engine" ]; then
  echo "build_ref.sh: $REF not present (GPU box?) - using prebuilt oracle/_ref if any" >&2
  [ -f "$OUT/lib/libinference_engine.so" ] && exit 0 || exit 3
fi
if [ -f "$OUT/lib/libMKLDNNPlugin.so" ] && [ -f "$OUT/bin/benchmark_app" ] && [ -f "$OUT/bin/ie_run" ] && [ "${FORCE:-0}" != 1 ]; then
  echo "build_ref.sh: oracle/_ref already built (FORCE=1 to rebuild)"; exit 0
fi

mkdir -p "$WORK"
if [ ! -f "$WORK/src/CMakeLists.txt" ]; then
  echo "[build_ref] copying reference tree to $WORK/src (build writes into its source dir)"
  mkdir -p "$WORK/src"
  (cd "$REF" && tar cf - --exclude=./.git . ) | (cd "$WORK/src" && tar xf -)
  chmod -R u+w "$WORK/src"
  # workaround 3: benchmark_app only uses OpenCV for -load_config/-dump_config
  sed -i 's/OPENCV_DEPENDENCIES core)/)/' "$WORK/src/inference-engine/samples/benchmark_app/CMakeLists.txt"
fi
# THREADING=OMP over the system libgomp (the resolver honours $OMP instead of downloading)
mkdir -p "$WORK/fakeomp/lib"
ln -sf /lib/x86_64-linux-gnu/libgomp.so.1 "$WORK/fakeomp/lib/libiomp5.so"

mkdir -p "$WORK/build"
cd "$WORK/build"
if [ ! -f build.ninja ]; then
  OMP="$WORK/fakeomp" cmake -G Ninja "$WORK/src" \
    -DCMAKE_C_COMPILER=/usr/bin/gcc -DCMAKE_CXX_COMPILER=/usr/bin/g++ \
    -DCMAKE_BUILD_TYPE=Release -DCMAKE_POLICY_VERSION_MINIMUM=3.5 \
    -DTHREADING=OMP -DENABLE_MKL_DNN=ON -DENABLE_CLDNN=OFF -DENABLE_GNA=OFF -DENABLE_VPU=OFF \
    -DENABLE_MYRIAD=OFF -DENABLE_OPENCV=OFF -DNGRAPH_ONNX_IMPORT_ENABLE=OFF \
    -DNGRAPH_PDPD_FRONTEND_ENABLE=OFF -DENABLE_TESTS=OFF -DENABLE_SAMPLES=ON -DENABLE_PYTHON=OFF \
    -DENABLE_DOCS=OFF -DENABLE_SPEECH_DEMO=OFF -DENABLE_CPPLINT=OFF -DENABLE_CLANG_FORMAT=OFF \
    -DTREAT_WARNING_AS_ERROR=OFF \
    -DCMAKE_CXX_FLAGS="-include cstdint -include limits -include cstddef -include cstring -include string -w" \
    > "$WORK/configure.log" 2>&1
fi
nice -n 5 ninja -j"$JOBS" MKLDNNPlugin inference_engine inference_engine_ir_reader \
    MultiDevicePlugin templatePlugin benchmark_app > "$WORK/build.log" 2>&1
# The reference's own conformance binary (docs/template_plugin/tests/functional: the behaviour suites, the hetero / single-layer
# instances and the op_reference tests a device plugin is expected to pass), UNMODIFIED.  tools/run_conformance.sh runs it against
# libb200Plugin.so by registering that library under the name the binary asks for.  ~430 extra translation units (funcSharedTests,
# sharedTestClasses, ngraphFunctions, gtest), ~15 min on 8 cores; CONFORMANCE=0 skips it.
if [ "${CONFORMANCE:-1}" = 1 ]; then
  OMP="$WORK/fakeomp" cmake . -DENABLE_TESTS=ON -DENABLE_FUNCTIONAL_TESTS=ON -DENABLE_BEH_TESTS=ON -DENABLE_DATA=OFF \
      -DNGRAPH_UNIT_TEST_ENABLE=OFF -DENABLE_GAPI_TESTS=OFF > "$WORK/configure_tests.log" 2>&1
  nice -n 5 ninja -j"$JOBS" templateFuncTests > "$WORK/build_tests.log" 2>&1
fi

BIN="$WORK/src/bin/intel64/Release"
mkdir -p "$OUT/lib" "$OUT/bin"
for f in libinference_engine.so libinference_engine_transformations.so libinference_engine_legacy.so \
         libinference_engine_lp_transformations.so libinference_engine_preproc.so \
         libinference_engine_ir_reader.so libinference_engine_ir_v7_reader.so libngraph.so \
         libMKLDNNPlugin.so libMultiDevicePlugin.so libtemplatePlugin.so libngraph_backend.so \
         libinterpreter_backend.so libformat_reader.so libfrontend_manager.so libHeteroPlugin.so libmock_engine.so; do
  [ -f "$BIN/lib/$f" ] && cp -f "$BIN/lib/$f" "$OUT/lib/"
done
cp -f "$BIN/benchmark_app" "$OUT/bin/"
[ -f "$BIN/templateFuncTests" ] && cp -f "$BIN/templateFuncTests" "$OUT/bin/"
strip --strip-unneeded "$OUT"/lib/*.so "$OUT/bin/benchmark_app" || true
# plugins.xml: the reference's own registry + one line for the drop-in (location is relative to
# libinference_engine.so, ie_core.cpp:421-433)
cat > "$OUT/lib/plugins.xml" <<XML
<ie>
    <plugins>
        <plugin name="CPU" location="libMKLDNNPlugin.so"></plugin>
        <plugin name="MULTI" location="libMultiDevicePlugin.so"></plugin>
        <plugin name="HETERO" location="libHeteroPlugin.so"></plugin>
        <plugin name="TEMPLATE" location="libtemplatePlugin.so"></plugin>
        <plugin name="B200" location="../../../plugin/lib/libb200Plugin.so"></plugin>
    </plugins>
</ie>
XML
# ie_run: the small driver (tools/ie_run.cpp, OUR source) that feeds raw inputs through the reference's public API;
# compiled against the reference headers where they lie, linked to the libraries built above
/usr/bin/g++ -O2 -std=c++14 -include cstdint -w -I"$REF/inference-engine/include" -I"$REF/ngraph/core/include" -I"$REF/inference-engine/src/plugin_api" \
    "$HERE/../tools/ie_run.cpp" -o "$OUT/bin/ie_run" -L"$OUT/lib" -linference_engine -lngraph -Wl,-rpath,'$ORIGIN/../lib'
echo "$WORK/src" > "$OUT/BUILT_FROM.txt"
date -u +%FT%TZ >> "$OUT/BUILT_FROM.txt"
[ "${KEEP_REF_BUILD:-1}" = 1 ] || rm -rf "$WORK"
echo "[build_ref] done -> $OUT"
