#!/bin/bash
# Cross-check of the CPU oracle against the REAL reference: build lolzballs/vkhel
# as its own CI does (.github/workflows/lavapipe.yml:17-28: meson + glslang over
# a Vulkan ICD, Mesa lavapipe or the NVIDIA driver), run its own tests, then run
# seeded random vectors through its public API (tools/vulkan_dump.c) and diff
# every result against oracle/ (tools/vulkan_parity_check.py).
#
#   tools/vulkan_parity.sh [--check-tools] [REFERENCE_DIR]
#
# Needs: meson, ninja, glslangValidator (src/kernels/shaders/meson.build:12-18),
# the Vulkan headers + loader (pkg-config vulkan), a Vulkan ICD, a C and a C++
# compiler, and the VulkanMemoryAllocator 3.0.1 sources the reference's wrap file
# downloads (subprojects/vulkan-memory-allocator.wrap) -- either network access
# for meson or $VKHEL_VMA_TARBALL pointing at VulkanMemoryAllocator-v3.0.1.tar.gz.
# Exit status: 0 identical, 1 mismatch or build failure, 77 prerequisites missing
# (the names are printed on one line starting with "missing:").
set -u
check_only=0
if [ "${1:-}" = "--check-tools" ]; then check_only=1; shift; fi
REF=${1:-/root/reference}
ROOT=$(cd "$(dirname "$0")/.." && pwd)

missing=()
for tool in meson ninja glslangValidator pkg-config cc c++; do
  command -v $tool > /dev/null 2>&1 || missing+=("$tool")
done
if command -v pkg-config > /dev/null 2>&1 && ! pkg-config --exists vulkan; then
  missing+=("vulkan-headers+loader(pkg-config vulkan)")
fi
icd=${VK_ICD_FILENAMES:-}
if [ -z "$icd" ]; then
  icd=$(ls /usr/share/vulkan/icd.d/*.json /etc/vulkan/icd.d/*.json 2> /dev/null | head -1)
fi
[ -n "$icd" ] || missing+=("vulkan-icd(json under /usr/share/vulkan/icd.d or \$VK_ICD_FILENAMES)")
[ -f "$REF/meson.build" ] || missing+=("reference-sources($REF)")
if [ ${#missing[@]} -gt 0 ]; then
  echo "missing: ${missing[*]}"
  exit 77
fi
echo "tools present; Vulkan ICD: $icd"
[ $check_only = 1 ] && exit 0

work=$(mktemp -d /tmp/vkhel_vulkan_parity.XXXXXX)
trap 'rm -rf "$work"' EXIT
cp -r "$REF" "$work/ref"          # the reference tree is read-only; meson writes into subprojects/
if [ -n "${VKHEL_VMA_TARBALL:-}" ]; then
  mkdir -p "$work/ref/subprojects/packagecache"
  cp "$VKHEL_VMA_TARBALL" "$work/ref/subprojects/packagecache/VulkanMemoryAllocator-v3.0.1.tar.gz"
fi
export VK_ICD_FILENAMES="$icd"
( cd "$work/ref" && meson setup build && meson compile -C build ) || { echo "reference build failed"; exit 1; }
( cd "$work/ref" && meson test -C build -v ) || { echo "the reference's own tests failed on this ICD"; exit 1; }
cc -O2 -I"$work/ref/include/vkhel" "$ROOT/tools/vulkan_dump.c" -L"$work/ref/build" -lvkhel \
   -Wl,-rpath,"$work/ref/build" -o "$work/vulkan_dump" || exit 1
"$work/vulkan_dump" > "$work/dump.txt" || { echo "dump program failed"; exit 1; }
make -C "$ROOT" -s oracle
python "$ROOT/tools/vulkan_parity_check.py" "$work/dump.txt"
